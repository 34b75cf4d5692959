// rr_route.cu -- the Muskingum wavefront solve for sm_100a.
//
// Replaces the three numba loops of river_route/routers/_numba_kernels.py
// (muskingum_route :9-46, rapid_route :49-84, unit_route :88-171).
//
// Formulation (per reach i, upstream set U(i) in ascending index order, one routing substep):
//     q'[i] = c3*q[i] + c4_dt*ql[t,i] + sum_u c2[i]*q[u] + sum_u c1[i]*q'[u]
// which is the reference's push / forward-substitution pair written as a pull; the two sums are
// accumulated one after the other in ascending upstream order exactly as the reference's
// column-ascending scatters deliver them.
//
// Parallel decomposition
//   * A work item is (block b of 32 consecutive reaches, tile j of `tile_rows` output rows) and is
//     executed by ONE WARP, lane = reach.  Consecutive reaches keep the lateral / discharge rows
//     coalesced in the reference's own (T, n) layout -- no permutation of the user's arrays.
//   * Inside an item the warp is a systolic array: lane l runs `skew[l]` steps behind, and the plan
//     guarantees that every in-block upstream lane is exactly ONE step ahead, so its newest and
//     previous discharge are fetched with warp shuffles -- no shared memory, no in-block flags.
//   * Between blocks, a reach whose downstream lives in another block exports its raw (unclamped)
//     substep series for the tile into its private ring of rows in the exchange buffer.  Row layout
//     (doubles): [14] = q_full carry-in (UNIT), [15] = carry-in, [16 + s] = value after substep s, so the
//     series starts on a 128-byte line and the four values of steps 4g..4g+3 are one aligned sector.  The producer publishes
//     "tile j done" with a release store on done[b]; consumers poll it and acquire.  A reach therefore
//     advances to tile j+1 as soon as its upstream blocks and its own tile j are done: the wavefront
//     pipelines time through deep networks with no grid-wide barrier and one launch per call.
//   * Persistent warps take tickets from a global counter; the ticket order (rr_plan.cpp) is a
//     linear extension of all dependencies, so the lowest unfinished ticket can always finish.
//   * Blocks without in-block edges and with at most RR_MAX_FAST_DEG upstreams per reach (all blocks
//     of a level-sorted network) run a register-blocked fast path: four time steps per iteration,
//     32-byte exchange loads/stores, next iteration's operands in flight while this one computes.
#include <cuda_runtime.h>

#include "rr_route.cuh"

#ifndef RR_SPIN_NS0
#define RR_SPIN_NS0 64
#endif
#ifndef RR_SPIN_NSMAX
#define RR_SPIN_NSMAX 4096
#endif
#ifndef RR_MIN_CTAS
#define RR_MIN_CTAS 2
#endif
#define FULL_MASK 0xffffffffu
#define SLOT_NONE ((int32_t)0x80000000)
#define RAW_QF 14     // row entry holding the q_full carry-in (UNIT)
#define RAW_CARRY 15  // row entry holding the value before the tile's first substep
#define RAW_S0 16     // row entry of substep 0: the series starts on a 128-byte line

#ifdef RR_PROFILE
#define PROF_DECL long long prof_t = clock64(); unsigned long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define PROF_MARK(k) { const long long n_ = clock64(); prof_acc[k] += (unsigned long long)(n_ - prof_t); prof_t = n_; }
#define PROF_FLUSH if ((threadIdx.x & 31) == 0 && P.prof) { for (int k_ = 0; k_ < 8; ++k_) atomicAdd(P.prof + k_, prof_acc[k_]); }
#else
#define PROF_DECL
#define PROF_MARK(k)
#define PROF_FLUSH
#endif

namespace {

__device__ __forceinline__ int32_t ld_relaxed(const int32_t *p) {
    int32_t v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int32_t ld_acquire(const int32_t *p) {
    int32_t v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int32_t *p, int32_t v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Spin with relaxed loads (an acquire load invalidates the SM's L1 on every poll), then acquire once.
__device__ __forceinline__ void wait_ge(const int32_t *flag, int32_t want) {
    if (ld_acquire(flag) >= want) return;          // the common case: one round trip
    unsigned ns = RR_SPIN_NS0;
    while (ld_relaxed(flag) < want) {
        __nanosleep(ns);
        if (ns < RR_SPIN_NSMAX) ns <<= 1;
    }
    (void)ld_acquire(flag);
}
// latency mode (narrow levels of the block DAG, where the launch is dependency bound rather than bandwidth bound)
__device__ __forceinline__ void prefetch_l2_now(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// L2 prefetch hints are compiled out by default: measured on B200 the kernel is limited by L2 transaction
// throughput (~2.4 L2 bytes per DRAM byte), and every prefetched line crosses L2 twice (-7% with them on).
#ifdef RR_L2_PREFETCH
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#else
__device__ __forceinline__ void prefetch_l2(const void *) {}
#endif
// streaming read of data that is never written during the launch
__device__ __forceinline__ double ld_stream(const double *p) { return __ldg(p); }

struct d4 { double a, b, c, d; };
// one aligned 32-byte sector of an exchange row (written by another SM earlier in this launch:
// plain coherent loads, ordered after the acquire fence)
__device__ __forceinline__ d4 ld_sector(const double *p) {
    // one 256-bit load (LDG.E.256 on sm_100a); L2 fetches the whole 128-byte line, i.e. this and the next
    // three sectors of the series, in one DRAM burst instead of four half-used 64-byte ones
    d4 v;
    asm volatile("ld.global.L2::128B.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.d) : "l"(p) : "memory");
    return v;
}
// read-only data (never written during the launch): non-coherent path
__device__ __forceinline__ d4 ld_sector_ro(const double *p) {
    d4 v;
    asm volatile("ld.global.nc.L2::128B.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.d) : "l"(p));
    return v;
}
// one 256-bit store: the whole sector is written at once, so L2 never has to fill it from DRAM first
__device__ __forceinline__ void st_sector(double *p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// Lateral inflow of working reach u at row `row` of tile j (UnitMuskingum reads its upstreams' laterals).
__device__ __forceinline__ const double *lat_ptr(const rr_route_params &P, int m, int j, int t0, int64_t u, int row) {
    if (P.tile_major == 3)   // [tile][block][group of 4 rows][lane][4]: a warp's 256-bit loads of its own rows are one contiguous KB
        return P.lateral[m] + ((((size_t)j * P.n_blocks + (size_t)(u >> 5)) * (size_t)(P.tile_pitch >> 2) + (size_t)(row >> 2)) * RR_BLOCK +
                               (size_t)(u & 31)) * 4 + (row & 3);
    if (P.tile_major == 2)
        return P.lateral[m] + (((size_t)j * P.n_blocks + (size_t)(u >> 5)) * RR_BLOCK + (size_t)(u & 31)) * (size_t)P.tile_pitch + row;
    if (P.tile_major == 1)
        return P.lateral[m] + (((size_t)j * P.n_blocks + (size_t)(u >> 5)) * (size_t)P.tile_rows + row) * RR_BLOCK + (u & 31);
    return P.lateral[m] + (size_t)(t0 + row) * P.ldl + u;
}
// Direct exchange (P.direct; level-sorted plans, one substep per row, reach-major discharge tiles): the working
// discharge array holds the RAW series and doubles as the exchange buffer -- a consumer reads its upstream reach's
// rows straight from that reach's discharge tile, the clamp moves into permute_to_user, and the separate exchange
// write (8 of ~40 B per reach-timestep), the ring buffer and the ring-reuse waits disappear.
// Row 0 of reach u's series in tile jj:
__device__ __forceinline__ const double *direct_tile(const rr_route_params &P, int m, int jj, int64_t u) {
    return P.out[m] + (((size_t)jj * P.n_blocks + (size_t)(u >> 5)) * RR_BLOCK + (size_t)(u & 31)) * (size_t)P.tile_pitch;
}
// value of reach u before the first substep of tile j: the shared initial state, or the last row of its previous tile
__device__ __forceinline__ double direct_carry(const rr_route_params &P, int m, int j, int64_t u) {
    if (j == 0) return P.q_init[(size_t)m * P.q_init_stride + u];
    return direct_tile(P, m, j - 1, u)[P.tile_rows - 1];
}
// offset of row r of one reach's lateral series from its row 0 (layout 3: groups of 4 rows are 32 lanes x 4 doubles apart)
__device__ __forceinline__ size_t lat_row_off(int layout, int64_t stride, int r) {
    return layout == 3 ? (((size_t)(r >> 2)) << 7) + (size_t)(r & 3) : (size_t)r * (size_t)stride;
}

struct item_ctx {
    int m, b, j, lane;
    int64_t i;
    bool valid, use_init;
    bool narrow;                 // the block's level has few blocks: optimise the item for latency, not bandwidth
#ifdef RR_PROFILE
    mutable unsigned long long prof_wait, prof_setup;
#endif
    int t0, rows, TT;
    double *raw_m;
    double c1, c2, c3, c4, q;   // q = state before the tile (UNIT: q_ch)
    int e0, deg, ex;
    rr_blk_meta M;
    int2 ro_me;                  // {first row, ring depth} of this lane's exported series (valid when ex >= 0)
    int2 ro_up[RR_MAX_FAST_DEG]; // the same pair for the first upstream slots (fast-path blocks)
    int dep_lo, dep_hi;          // this block's range in dep_idx (from the ticket table)
    int32_t up_u[RR_MAX_FAST_DEG]; // direct exchange: working index of the first upstream reaches
    const double *lat0;          // this lane's lateral value of the tile's first row
    int64_t lstride;             // distance between consecutive rows of the lateral tile
    double *out0;                // this lane's discharge value of the tile's first row
    int64_t ostride;
};

// ------------------------------------------------------------------------------------------------
// Fast path: K == 1, no in-block edges, every reach has at most NS upstreams (all external).
// ------------------------------------------------------------------------------------------------
template <int MODE, int NS, bool VEC>
__device__ __forceinline__ void fast_item(const rr_route_params &P, const item_ctx &c) {
#ifdef RR_PROFILE
    const long long s0_ = clock64();
#endif
    const double c1 = c.c1, c2 = c.c2, c3 = c.c3, c4 = c.c4;
    const int e0 = c.e0, deg = c.deg, ex = c.ex;
    double q = c.q;
    constexpr bool HAS_LAT = (MODE == RR_MODE_RAPID);
    constexpr int NA = NS > 0 ? NS : 1;
    const int j = c.j;
    const double *up[NA];
    bool has[NA];
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        has[k] = k < deg;
        up[k] = c.raw_m;
        if (has[k]) {
            const int2 ro = c.ro_up[k];
            up[k] = P.direct ? direct_tile(P, c.m, j, c.up_u[k]) - RAW_S0
                             : c.raw_m + ((size_t)ro.x + (size_t)(j % ro.y)) * P.raw_pitch;
        }
    }
    double *myraw = nullptr;
    if (ex >= 0 && !P.direct) {
        const int2 ro = c.ro_me;
        myraw = c.raw_m + ((size_t)ro.x + (size_t)(j % ro.y)) * P.raw_pitch;
        myraw[RAW_CARRY] = q;
    }
    const bool raw_out = P.direct != 0;   // the discharge tile keeps the raw series; permute_to_user clamps
    const double *lat = c.lat0;
    double *outp = c.out0;
    const int64_t ldl = c.lstride, ldo = c.ostride;
    const int TT = c.TT;

    // Software pipeline, two groups (8 time steps) deep: while group g is computed, the loads of groups g+1 and
    // g+2 are in flight.  Under load one memory round trip takes ~2500 cycles, a group ~600 cycles of arithmetic.
    auto lat_group = [&](int s0) -> d4 {
        d4 v{0, 0, 0, 0};
        if (!(HAS_LAT && c.valid) || s0 >= TT) return v;
        if (VEC) return ld_sector_ro(lat + lat_row_off(P.tile_major, 1, s0));   // 4 rows of this reach: one aligned sector
        const double *lp = lat + (size_t)s0 * ldl;
        v.a = ld_stream(lp);
        if (s0 + 1 < TT) v.b = ld_stream(lp + ldl);
        if (s0 + 2 < TT) v.c = ld_stream(lp + 2 * ldl);
        if (s0 + 3 < TT) v.d = ld_stream(lp + 3 * ldl);
        return v;
    };
    double old[NA];   // upstream value before the group's first substep
    d4 nxt[NA], fut[NA];
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        old[k] = 0.0;
        nxt[k] = fut[k] = d4{0, 0, 0, 0};
        if (has[k]) {
            old[k] = P.direct ? direct_carry(P, c.m, j, c.up_u[k]) : up[k][RAW_CARRY];
            nxt[k] = ld_sector(up[k] + RAW_S0);
            if (4 < TT) fut[k] = ld_sector(up[k] + RAW_S0 + 4);
        }
    }
    d4 lcur = lat_group(0), lnxt = lat_group(4);
#ifdef RR_PROFILE
    __syncwarp();
    c.prof_setup = (unsigned long long)(clock64() - s0_);
#endif
    for (int s = 0; s < TT; s += 4) {
        // ---- issue the loads of group g+2 ----
        d4 far[NA];
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            far[k] = d4{0, 0, 0, 0};
            if (has[k] && s + 8 < TT) far[k] = ld_sector(up[k] + RAW_S0 + s + 8);
        }
        const d4 lfar = lat_group(s + 8);
        // ---- four substeps; old = value before the substep, new = value after it ----
        // step s: old = old[k], new = nxt.a;  step s+1: old = nxt.a, new = nxt.b;  ...
        double r0, r1, r2, r3;
        {
            double r = c3 * q;                                       // _numba_kernels.py:27-28 / :68-69
            if (HAS_LAT) r = fma(c4, lcur.a, r);
#pragma unroll
            for (int k = 0; k < NS; ++k) r = fma(c2, old[k], r);     // :29-33 / :70-74, ascending upstream
#pragma unroll
            for (int k = 0; k < NS; ++k) r = fma(c1, nxt[k].a, r);   // :36-39 / :75-78 (lhs_off = -c1)
            r0 = r;
            r = c3 * r0;
            if (HAS_LAT) r = fma(c4, lcur.b, r);
#pragma unroll
            for (int k = 0; k < NS; ++k) r = fma(c2, nxt[k].a, r);
#pragma unroll
            for (int k = 0; k < NS; ++k) r = fma(c1, nxt[k].b, r);
            r1 = r;
            r = c3 * r1;
            if (HAS_LAT) r = fma(c4, lcur.c, r);
#pragma unroll
            for (int k = 0; k < NS; ++k) r = fma(c2, nxt[k].b, r);
#pragma unroll
            for (int k = 0; k < NS; ++k) r = fma(c1, nxt[k].c, r);
            r2 = r;
            r = c3 * r2;
            if (HAS_LAT) r = fma(c4, lcur.d, r);
#pragma unroll
            for (int k = 0; k < NS; ++k) r = fma(c2, nxt[k].c, r);
#pragma unroll
            for (int k = 0; k < NS; ++k) r = fma(c1, nxt[k].d, r);
            r3 = r;
        }
        if (c.valid && ldo == 1 && raw_out) {
            st_sector(outp + s, r0, r1, r2, r3);
        } else if (c.valid && ldo == 1) {   // this reach's rows are contiguous in the working discharge array
            st_sector(outp + s, r0 > 0.0 ? r0 : 0.0, r1 > 0.0 ? r1 : 0.0, r2 > 0.0 ? r2 : 0.0, r3 > 0.0 ? r3 : 0.0);
        } else if (c.valid) {
            // K == 1: the interval mean is the value itself; clamp as :44-46 / :82-84
            double *o = outp + (size_t)s * ldo;
            o[0] = r0 > 0.0 ? r0 : 0.0;
            if (s + 1 < TT) o[ldo] = r1 > 0.0 ? r1 : 0.0;
            if (s + 2 < TT) o[2 * ldo] = r2 > 0.0 ? r2 : 0.0;
            if (s + 3 < TT) o[3 * ldo] = r3 > 0.0 ? r3 : 0.0;
        }
        if (myraw) st_sector(myraw + RAW_S0 + s, r0, r1, r2, r3);
        q = (s + 3 < TT) ? r3 : ((s + 2 < TT) ? r2 : ((s + 1 < TT) ? r1 : r0));
#pragma unroll
        for (int k = 0; k < NS; ++k) { old[k] = nxt[k].d; nxt[k] = fut[k]; fut[k] = far[k]; }
        lcur = lnxt;
        lnxt = lfar;
    }
    if (c.valid) P.q_state[c.m][c.i] = q;
}

// ------------------------------------------------------------------------------------------------
// Fast path for K > 1 routing substeps per row (dt_routing < dt_runoff): same register blocking over
// substeps; the lateral value changes and the interval mean is written every K substeps
// (_numba_kernels.py:60-84 / :19-46).
// ------------------------------------------------------------------------------------------------
template <int MODE, int NS>
__device__ __forceinline__ void fast_item_k(const rr_route_params &P, const item_ctx &c) {
    constexpr bool HAS_LAT = (MODE == RR_MODE_RAPID);
    constexpr int NA = NS > 0 ? NS : 1;
    const double c1 = c.c1, c2 = c.c2, c3 = c.c3, c4 = c.c4;
    const int e0 = c.e0, deg = c.deg, ex = c.ex, j = c.j, K = P.K, TT = c.TT;
    const double inv_k = 1.0 / (double)K;
    double q = c.q;
    const double *up[NA];
    bool has[NA];
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        has[k] = k < deg;
        up[k] = c.raw_m;
        if (has[k]) {
            const int2 ro = c.ro_up[k];
            up[k] = c.raw_m + ((size_t)ro.x + (size_t)(j % ro.y)) * P.raw_pitch;
        }
    }
    double *myraw = nullptr;
    if (ex >= 0) {
        const int2 ro = c.ro_me;
        myraw = c.raw_m + ((size_t)ro.x + (size_t)(j % ro.y)) * P.raw_pitch;
        myraw[RAW_CARRY] = q;
    }
    const double *lat = c.lat0;
    double *outp = c.out0;
    const int64_t ldl = c.lstride, ldo = c.ostride;
    double old[NA];
    d4 nxt[NA];
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        old[k] = 0.0;
        nxt[k] = d4{0, 0, 0, 0};
        if (has[k]) { old[k] = up[k][RAW_CARRY]; nxt[k] = ld_sector(up[k] + RAW_S0); }
    }
    double ql = 0.0, ql_next = 0.0, acc = 0.0;
    int sub = 0, row = 0;
    if (HAS_LAT && c.valid) {
        ql = ld_stream(lat);
        if (1 < c.rows) ql_next = ld_stream(lat + ldl);
    }
    for (int s = 0; s < TT; s += 4) {
        d4 fut[NA];
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            fut[k] = d4{0, 0, 0, 0};
            if (has[k] && s + 4 < TT) fut[k] = ld_sector(up[k] + RAW_S0 + s + 4);
        }
        double r4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            double r = c3 * q;                                        // :27-28 / :68-69
            if (HAS_LAT) r = fma(c4, ql, r);
#pragma unroll
            for (int k = 0; k < NS; ++k) {
                const double vo = u == 0 ? old[k] : (u == 1 ? nxt[k].a : (u == 2 ? nxt[k].b : nxt[k].c));
                r = fma(c2, vo, r);                                   // :29-33 / :70-74
            }
#pragma unroll
            for (int k = 0; k < NS; ++k) {
                const double vn = u == 0 ? nxt[k].a : (u == 1 ? nxt[k].b : (u == 2 ? nxt[k].c : nxt[k].d));
                r = fma(c1, vn, r);                                   // :36-39 / :75-78
            }
            r4[u] = r;
            if (s + u < TT) {
                q = r;
                acc += r;                                             // :41-42 / :79-80
                if (++sub == K) {
                    const double v = acc * inv_k;                     // :44-46 / :82-84
                    if (c.valid) outp[(size_t)row * ldo] = v > 0.0 ? v : 0.0;
                    acc = 0.0;
                    sub = 0;
                    ++row;
                    ql = ql_next;
                    if (HAS_LAT && c.valid && row + 1 < c.rows) ql_next = ld_stream(lat + (size_t)(row + 1) * ldl);
                }
            }
        }
        if (myraw) st_sector(myraw + RAW_S0 + s, r4[0], r4[1], r4[2], r4[3]);
#pragma unroll
        for (int k = 0; k < NS; ++k) { old[k] = nxt[k].d; nxt[k] = fut[k]; }
    }
    if (c.valid) P.q_state[c.m][c.i] = q;
}

// ------------------------------------------------------------------------------------------------
// UnitMuskingum fast path: K == 1, no in-block edges, at most NS upstreams per reach.
// Headwater lanes copy their lateral inflow to the output (unclamped, _numba_kernels.py:122-123); inner
// lanes read, per upstream, its lateral series (from the lateral working array) and -- for inner upstreams --
// its exported q_ch series; q_full_old of an upstream is q_ch_old + its lateral of the previous step (:165).
// ------------------------------------------------------------------------------------------------
template <int NS, bool VEC>
__device__ __forceinline__ void unit_fast_item(const rr_route_params &P, const item_ctx &c) {
    constexpr int NA = NS > 0 ? NS : 1;
    const double c1 = c.c1, c2 = c.c2, c3 = c.c3;
    const int e0 = c.e0, deg = c.deg, ex = c.ex, j = c.j, m = c.m, TT = c.TT;
    const bool inner = deg > 0;
    const double *lat = c.lat0;
    double *outp = c.out0;
    const int64_t ldl = c.lstride, ldo = c.ostride;
    double q = c.q;                                                      // q_ch
    double qf = c.use_init ? q : (c.valid ? P.q_full[m][c.i] : 0.0);     // q_full (UnitMuskingum.py:78-79)
    const double *ex_row[NA];      // exported q_ch series of an inner upstream (nullptr: headwater upstream)
    const double *lu_row[NA];      // lateral series of the upstream
    int64_t lu_stride = 1;
    bool has[NA], hw[NA];
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        has[k] = k < deg;
        hw[k] = false;
        ex_row[k] = nullptr;
        lu_row[k] = c.raw_m;
        if (has[k]) {
            const int32_t sk = __ldg(P.slot_src + e0 + k);
            hw[k] = (sk & RR_SLOT_HW_BIT) != 0;
            const int64_t u = __ldg(P.up_idx + e0 + k);
            lu_row[k] = lat_ptr(P, m, j, c.t0, u, 0);
            if (!hw[k]) {
                const int2 ro = c.ro_up[k];
                ex_row[k] = c.raw_m + ((size_t)ro.x + (size_t)(j % ro.y)) * P.raw_pitch;
            }
        }
    }
    if (P.tile_major == 1) lu_stride = RR_BLOCK; else if (P.tile_major == 0) lu_stride = P.ldl;
    double *myraw = nullptr;
    if (ex >= 0 && inner) {
        const int2 ro = c.ro_me;
        myraw = c.raw_m + ((size_t)ro.x + (size_t)(j % ro.y)) * P.raw_pitch;
        myraw[RAW_CARRY] = q;
        myraw[RAW_QF] = qf;
    }
    auto group = [&](const double *p, int64_t stride, int s0, bool on, bool vec) -> d4 {
        d4 v{0, 0, 0, 0};
        if (!on || s0 >= TT) return v;
        if (vec) return ld_sector_ro(p + lat_row_off(P.tile_major, 1, s0));
        const double *lp = p + (size_t)s0 * stride;
        v.a = ld_stream(lp);
        if (s0 + 1 < TT) v.b = ld_stream(lp + stride);
        if (s0 + 2 < TT) v.c = ld_stream(lp + 2 * stride);
        if (s0 + 3 < TT) v.d = ld_stream(lp + 3 * stride);
        return v;
    };
    const bool uvec = P.tile_major >= 2;
    // state carried between groups: per inner upstream its q_full before the group's first substep
    double qfo[NA];
    d4 exn[NA], lun[NA];
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        qfo[k] = 0.0;
        exn[k] = d4{0, 0, 0, 0};
        if (has[k] && !hw[k]) { qfo[k] = ex_row[k][RAW_QF]; exn[k] = ld_sector(ex_row[k] + RAW_S0); }
        lun[k] = group(lu_row[k], lu_stride, 0, has[k], uvec);
    }
    d4 lcur = group(lat, ldl, 0, c.valid, VEC);
    double l_last = 0.0;
    for (int s = 0; s < TT; s += 4) {
        d4 exf[NA], luf[NA];
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            exf[k] = d4{0, 0, 0, 0};
            if (has[k] && !hw[k] && s + 4 < TT) exf[k] = ld_sector(ex_row[k] + RAW_S0 + s + 4);
            luf[k] = group(lu_row[k], lu_stride, s + 4, has[k], uvec);
        }
        const d4 lnxt = group(lat, ldl, s + 4, c.valid, VEC);
        const double lv[4] = {lcur.a, lcur.b, lcur.c, lcur.d};
        double r4[4], o4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            double a_in = 0.0, a_hw = 0.0;                               // :126-139, ascending upstream per class
#pragma unroll
            for (int k = 0; k < NS; ++k) {
                const double l = u == 0 ? lun[k].a : (u == 1 ? lun[k].b : (u == 2 ? lun[k].c : lun[k].d));
                if (has[k]) { if (hw[k]) a_hw += l; else a_in += l; }
            }
            double r = c1 * (a_in + a_hw) + c2 * a_hw;                   // :142-143, :151
            r = r + c3 * q;
#pragma unroll
            for (int k = 0; k < NS; ++k)
                if (has[k] && !hw[k]) r = fma(c2, qfo[k], r);            // :152-156  c2 * q_full_old[upstream]
#pragma unroll
            for (int k = 0; k < NS; ++k) {
                const double qn = u == 0 ? exn[k].a : (u == 1 ? exn[k].b : (u == 2 ? exn[k].c : exn[k].d));
                if (has[k] && !hw[k]) r = fma(c1, qn, r);                // :159-162  lhs_off = -c1
            }
            // upstream q_full after this substep = its q_ch + its lateral (:165-166), used by the next substep
#pragma unroll
            for (int k = 0; k < NS; ++k) {
                const double qn = u == 0 ? exn[k].a : (u == 1 ? exn[k].b : (u == 2 ? exn[k].c : exn[k].d));
                const double l = u == 0 ? lun[k].a : (u == 1 ? lun[k].b : (u == 2 ? lun[k].c : lun[k].d));
                qfo[k] = qn + l;
            }
            r4[u] = r;
            if (s + u < TT) {
                l_last = lv[u];
                if (inner) { q = r; qf = r + lv[u]; o4[u] = qf > 0.0 ? qf : 0.0; }   // :165-171 (K == 1: mean == value)
                else o4[u] = lv[u];                                                    // :122-123
            } else o4[u] = 0.0;
        }
        if (c.valid && ldo == 1) st_sector(outp + s, o4[0], o4[1], o4[2], o4[3]);
        else if (c.valid) {
            double *o = outp + (size_t)s * ldo;
            o[0] = o4[0];
            if (s + 1 < TT) o[ldo] = o4[1];
            if (s + 2 < TT) o[2 * ldo] = o4[2];
            if (s + 3 < TT) o[3 * ldo] = o4[3];
        }
        if (myraw) st_sector(myraw + RAW_S0 + s, r4[0], r4[1], r4[2], r4[3]);
#pragma unroll
        for (int k = 0; k < NS; ++k) { exn[k] = exf[k]; lun[k] = luf[k]; }
        lcur = lnxt;
    }
    if (c.valid) {
        const bool last = (j == P.n_tiles - 1);
        if (last && P.last_call) P.q_state[m][c.i] = inner ? qf : l_last;    // UnitMuskingum.py:94-98
        else { P.q_state[m][c.i] = q; P.q_full[m][c.i] = qf; }
    }
}

// ------------------------------------------------------------------------------------------------
// General path: systolic item with shuffles -- any skew, any in-degree, any number of substeps,
// UnitMuskingum.  Ends with the state write-back (the caller publishes done[b]).
// ------------------------------------------------------------------------------------------------
// GEN_SLOTS upstream slots are kept in registers (2: every reach of a binary network); further upstreams go
// through the slow loops.  Four register slots made this function spill, and its spill reloads miss L1 whenever
// another warp of the SM executes an acquire (L1 invalidate): ~3000 cycles per step on the critical path.
#define GEN_SLOTS 2
template <int MODE>
__device__ __noinline__ void general_item(const rr_route_params &P, const item_ctx &c) {
    constexpr bool HAS_LAT = (MODE != RR_MODE_MUSKINGUM);
    constexpr bool UNIT = (MODE == RR_MODE_UNIT);
    const int lane = c.lane, m = c.m, b = c.b, j = c.j, K = P.K;
    const int64_t i = c.i;
    const bool valid = c.valid, use_init = c.use_init;
    const int64_t ic = valid ? i : P.n - 1;
    const double c1 = c.c1, c2 = c.c2, c3 = c.c3, c4 = c.c4;
    const int e0 = c.e0, deg = c.deg, ex = c.ex;
    const rr_blk_meta M = c.M;
    const int t0 = c.t0, TT = c.TT;
    double *raw_m = c.raw_m;
    double qcur = c.q;
    const double inv_k = 1.0 / (double)K;  // _numba_kernels.py:19, :60, :104
    (void)b;
    const int d = __ldg(P.skew + ic);
    const int nfast = M.max_deg < GEN_SLOTS ? M.max_deg : GEN_SLOTS;
    auto raw_row = [&](int e) -> double * {   // e = entry of the upstream-CSR (an external edge)
        const int2 ro = __ldg(reinterpret_cast<const int2 *>(P.edge_ro) + e);
        return raw_m + ((size_t)ro.x + (size_t)(j % ro.y)) * P.raw_pitch;
    };
    int32_t src[GEN_SLOTS];
    const double *rp[GEN_SLOTS];   // exported series of an external upstream
    int ilane[GEN_SLOTS];          // lane of an in-block upstream (own lane if none)
    int64_t ug[GEN_SLOTS];         // UNIT: global index of the upstream reach
#pragma unroll
    for (int k = 0; k < GEN_SLOTS; ++k) {
        src[k] = (k < deg) ? __ldg(P.slot_src + e0 + k) : SLOT_NONE;
        rp[k] = nullptr;
        ilane[k] = lane;
        ug[k] = 0;
        if (src[k] != SLOT_NONE) {
            if (src[k] >= 0) {
                rp[k] = P.direct ? direct_tile(P, m, j, __ldg(P.up_idx + e0 + k)) - RAW_S0 : raw_row(e0 + k);
                // items on this path sit on the critical path of deep networks (chains inside a block): get the
                // whole upstream series moving towards L2 now, the loop then only pays L1 / L2 hits
                for (int e = 0; e < TT; e += 16) prefetch_l2_now(rp[k] + RAW_S0 + e);
            } else ilane[k] = (-src[k] - 1) & 31;
            if (UNIT) ug[k] = __ldg(P.up_idx + e0 + k);
        }
    }
    const double *lat = c.lat0;
    double *outp = c.out0;

    double qprev = qcur;
    double qf_cur = 0.0, qf_prev = 0.0;  // UNIT: q_full and its previous value
    if (UNIT && valid) {
        qf_cur = use_init ? qcur : P.q_full[m][i];
        qf_prev = qf_cur;
    }
    double *myraw = nullptr;
    if (ex >= 0 && !P.direct) {
        const int2 ro = c.ro_me;
        myraw = raw_m + ((size_t)ro.x + (size_t)(j % ro.y)) * P.raw_pitch;
        myraw[RAW_CARRY] = qcur;                 // carry-in for consumers
        if (UNIT) myraw[RAW_QF] = qf_cur;        // q_full carry-in
    }

    // one-step lookahead registers for the external series and the lateral row
    double eo[GEN_SLOTS], en[GEN_SLOTS];
    double lu[GEN_SLOTS], lu_old[GEN_SLOTS];   // UNIT: upstream lateral (this / previous row)
#pragma unroll
    for (int k = 0; k < GEN_SLOTS; ++k) {
        eo[k] = en[k] = lu[k] = lu_old[k] = 0.0;
        if (rp[k] && !(UNIT && (src[k] & RR_SLOT_HW_BIT))) {
            eo[k] = UNIT ? rp[k][RAW_QF] : (P.direct ? direct_carry(P, m, j, __ldg(P.up_idx + e0 + k)) : rp[k][RAW_CARRY]);
            en[k] = rp[k][RAW_S0];
        }
    }
    double ql = 0.0, ql_nx = 0.0;
    if (HAS_LAT && valid) ql_nx = ld_stream(lat);
    double acc = 0.0, base = 0.0;
    int sub = 0, row = 0;

    const int nsteps = TT + M.max_skew;
    for (int sig = 0; sig < nsteps; ++sig) {
        const int s = sig - d;
        const bool act = valid && s >= 0 && s < TT;
        const bool row_start = act && sub == 0;

        // ---- gather upstream values (shuffles are executed by every lane) ----
        double vo[GEN_SLOTS], vn[GEN_SLOTS];
#pragma unroll
        for (int k = 0; k < GEN_SLOTS; ++k) {
            vo[k] = eo[k];
            vn[k] = en[k];
            if (k < nfast && (M.int_mask >> k) & 1) {
                const double so = __shfl_sync(FULL_MASK, UNIT ? qf_prev : qprev, ilane[k]);
                const double sn = __shfl_sync(FULL_MASK, qcur, ilane[k]);
                if (src[k] < 0 && src[k] != SLOT_NONE) { vo[k] = so; vn[k] = sn; }
            }
        }

        double r = 0.0;
        if (UNIT) {
            if (row_start) {
                // _numba_kernels.py:116-143: lateral gathers, A_inner@ql and A_hw@ql in ascending
                // column order, c1*(a_inner + a_hw); rhs base = c1_A_ql + c2*a_hw (:151)
                ql = ql_nx;
                double a_in = 0.0, a_hw = 0.0;
#pragma unroll
                for (int k = 0; k < GEN_SLOTS; ++k) {
                    if (k < deg) {
                        lu_old[k] = lu[k];
                        lu[k] = ld_stream(lat_ptr(P, m, j, t0, ug[k], row));
                        const bool hw = src[k] >= 0 ? (src[k] & RR_SLOT_HW_BIT) != 0 : (((-src[k] - 1) >> 6) & 1) != 0;
                        if (hw) a_hw += lu[k]; else a_in += lu[k];
                    }
                }
                for (int k = GEN_SLOTS; k < deg; ++k) {
                    const int32_t sk = __ldg(P.slot_src + e0 + k);
                    const double l = ld_stream(lat_ptr(P, m, j, t0, __ldg(P.up_idx + e0 + k), row));
                    const bool hw = sk >= 0 ? (sk & RR_SLOT_HW_BIT) != 0 : (((-sk - 1) >> 6) & 1) != 0;
                    if (hw) a_hw += l; else a_in += l;
                }
                base = c1 * (a_in + a_hw) + c2 * a_hw;
            }
            r = base + c3 * qcur;
        } else {
            if (row_start) ql = ql_nx;
            r = c3 * qcur;                      // :27-28 / :68-69
            if (HAS_LAT) r = fma(c4, ql, r);
        }

        // ---- pass A: c2 * (previous-substep discharge of each upstream), ascending ----
#pragma unroll
        for (int k = 0; k < GEN_SLOTS; ++k) {
            if (k < deg) {
                if (UNIT) {
                    const bool hw = src[k] >= 0 ? (src[k] & RR_SLOT_HW_BIT) != 0 : (((-src[k] - 1) >> 6) & 1) != 0;
                    if (!hw) {
                        // external inner upstream: q_full_old = q_ch_old + lateral of the row that substep
                        // belonged to; at s == 0 the exported carry already is q_full
                        double qfo = vo[k];
                        if (src[k] >= 0 && s > 0) qfo = vo[k] + (sub == 0 ? lu_old[k] : lu[k]);
                        r = fma(c2, qfo, r);
                    }
                } else {
                    r = fma(c2, vo[k], r);
                }
            }
        }
        if (M.max_deg > GEN_SLOTS) {
            for (int k = GEN_SLOTS; k < M.max_deg; ++k) {
                const int32_t sk = (k < deg) ? __ldg(P.slot_src + e0 + k) : SLOT_NONE;
                const int il = (sk < 0 && sk != SLOT_NONE) ? ((-sk - 1) & 31) : lane;
                double v = __shfl_sync(FULL_MASK, UNIT ? qf_prev : qprev, il);
                bool use = act && k < deg;
                if (use && sk >= 0 && P.direct) {
                    const int64_t uu = __ldg(P.up_idx + e0 + k);
                    v = s == 0 ? direct_carry(P, m, j, uu) : direct_tile(P, m, j, uu)[s - 1];
                } else if (use && sk >= 0) {
                    const double *q = raw_row(e0 + k);
                    if (UNIT) {
                        if (sk & RR_SLOT_HW_BIT) use = false;
                        else if (s == 0) v = q[RAW_QF];
                        else {
                            const int prow = (sub == 0) ? row - 1 : row;
                            v = q[RAW_CARRY + s] + ld_stream(lat_ptr(P, m, j, t0, __ldg(P.up_idx + e0 + k), prow));
                        }
                    } else v = q[RAW_CARRY + s];
                } else if (use && UNIT && (((-sk - 1) >> 6) & 1)) use = false;
                if (use) r = fma(c2, v, r);
            }
        }
        // ---- pass B: c1 * (this-substep discharge of each upstream), ascending ----
#pragma unroll
        for (int k = 0; k < GEN_SLOTS; ++k) {
            if (k < deg) {
                if (UNIT) {
                    const bool hw = src[k] >= 0 ? (src[k] & RR_SLOT_HW_BIT) != 0 : (((-src[k] - 1) >> 6) & 1) != 0;
                    if (!hw) r = fma(c1, vn[k], r);
                } else {
                    r = fma(c1, vn[k], r);      // rhs -= lhs_off * q_new with lhs_off = -c1 (:36-39)
                }
            }
        }
        if (M.max_deg > GEN_SLOTS) {
            for (int k = GEN_SLOTS; k < M.max_deg; ++k) {
                const int32_t sk = (k < deg) ? __ldg(P.slot_src + e0 + k) : SLOT_NONE;
                const int il = (sk < 0 && sk != SLOT_NONE) ? ((-sk - 1) & 31) : lane;
                double v = __shfl_sync(FULL_MASK, qcur, il);
                bool use = act && k < deg;
                if (use && sk >= 0) {
                    if (UNIT && (sk & RR_SLOT_HW_BIT)) use = false;
                    else if (P.direct) v = direct_tile(P, m, j, __ldg(P.up_idx + e0 + k))[s];
                    else v = raw_row(e0 + k)[RAW_S0 + s];
                } else if (use && UNIT && (((-sk - 1) >> 6) & 1)) use = false;
                if (use) r = fma(c1, v, r);
            }
        }

        // ---- commit ----
        if (act) {
            const bool inner = !UNIT || deg > 0;
            if (inner) {
                qprev = qcur;
                qcur = r;
                if (UNIT) {
                    qf_prev = qf_cur;
                    qf_cur = r + ql;            // :165-166
                    acc += qf_cur;
                } else acc += r;                // :41-42 / :79-80
                if (myraw) myraw[RAW_S0 + s] = r;
            }
            if (++sub == K) {
                double v;
                if (UNIT && !inner) v = ql;     // headwater: lateral inflow, unclamped (:122-123)
                else { v = acc * inv_k; if (!P.direct) v = v > 0.0 ? v : 0.0; }   // :44-46 / :82-84 / :169-171 (direct: clamped by permute_to_user)
                outp[(size_t)row * c.ostride] = v;
                acc = 0.0;
                sub = 0;
                ++row;
            }
            // lookahead for the next step of this lane
            if (s + 1 < TT) {
#pragma unroll
                for (int k = 0; k < GEN_SLOTS; ++k) {
                    if (rp[k] && !(UNIT && (src[k] & RR_SLOT_HW_BIT))) { eo[k] = en[k]; en[k] = rp[k][RAW_S0 + s + 1]; }
                }
                if (HAS_LAT && sub == 0) ql_nx = ld_stream(lat + lat_row_off(P.tile_major, c.lstride, row));
            }
        }
    }

    // ---------------- publish ----------------
    if (valid) {
        const bool last = (j == P.n_tiles - 1);
        if (UNIT) {
            if (last && P.last_call) P.q_state[m][i] = deg > 0 ? qf_cur : ql;   // UnitMuskingum.py:94-98
            else { P.q_state[m][i] = qcur; P.q_full[m][i] = qf_cur; }
        } else P.q_state[m][i] = qcur;
    }
}

// ticket -> (member, block, tile); false when the tickets are exhausted
// Tickets are drawn in batches of P.ticket_batch consecutive tickets per warp: one global atomic per batch keeps
// the single-address atomic rate (an L2 atomic unit serialises them) off the critical path of large launches.
// A warp works through its batch in order, so the lowest unfinished ticket is still always being processed.
__device__ __forceinline__ bool next_item(const rr_route_params &P, int lane, int &m, int &b, int &j, int &dep_lo, int &dep_hi,
                                          unsigned long long &tk_next, unsigned long long &tk_end) {
    if (tk_next >= tk_end) {
        unsigned long long t0 = 0;
        if (lane == 0) t0 = atomicAdd(P.ticket, (unsigned long long)P.ticket_batch);
        tk_next = __shfl_sync(FULL_MASK, t0, 0);
        tk_end = tk_next + (unsigned long long)P.ticket_batch;
    }
    const unsigned long long tk = tk_next++;
    if (tk >= (unsigned long long)P.n_items * (unsigned)P.n_members) return false;
    m = 0;
    int64_t ticket = (int64_t)tk;
    if (P.n_members > 1) { m = (int)(tk % (unsigned)P.n_members); ticket = (int64_t)(tk / (unsigned)P.n_members); }
    const int4 it = __ldg(P.items + ticket);   // {block, tile, first and one-past-last entry of the block in dep_idx}
    b = it.x;
    j = it.y;
    dep_lo = it.z;
    dep_hi = it.w;
    return true;
}

// per-lane constants of an item, lateral prefetch, dependency waits, initial state
template <int MODE>
__device__ __forceinline__ void open_item(const rr_route_params &P, item_ctx &c, int lane, int m, int b, int j, int dep_lo,
                                          int dep_hi) {
    constexpr bool HAS_LAT = (MODE != RR_MODE_MUSKINGUM);
    constexpr bool UNIT = (MODE == RR_MODE_UNIT);
    const int64_t i = (int64_t)b * RR_BLOCK + lane;
    const bool valid = i < P.n;
    const int64_t ic = valid ? i : P.n - 1;
    c.m = m; c.b = b; c.j = j; c.lane = lane; c.i = i; c.valid = valid;
    c.c1 = __ldg(P.c1 + ic); c.c2 = __ldg(P.c2 + ic); c.c3 = __ldg(P.c3 + ic);
    c.c4 = HAS_LAT && !UNIT ? __ldg(P.c4 + ic) : 0.0;
    c.e0 = __ldg(P.up_ptr + ic);
    c.deg = valid ? __ldg(P.up_ptr + ic + 1) - c.e0 : 0;
    c.ex = valid ? __ldg(P.export_id + ic) : -1;
    c.M = P.meta[b];
    c.narrow = false;
    c.t0 = j * P.tile_rows;
    c.rows = min(P.tile_rows, P.T - c.t0);
    c.TT = c.rows * P.K;
    c.raw_m = P.raw + (size_t)m * (size_t)P.raw_rows * P.raw_pitch;
    // Working-array layouts (renumbered plans; the caller's arrays are permuted into them on the device):
    //   0 row-major (T, ld)   1 [tile][block][row][lane]   2 [tile][block][lane][row] (tile_pitch doubles per reach)
    //   3 [tile][block][row / 4][lane][row % 4] (lateral only)
    auto place = [&](int layout, const double *base, int64_t ld, const double *&p0, int64_t &stride) {
        if (layout == 3) { p0 = base + (((size_t)j * P.n_blocks + b) * (size_t)(P.tile_pitch >> 2) * RR_BLOCK + lane) * 4; stride = 1; }
        else if (layout == 2) { p0 = base + (((size_t)j * P.n_blocks + b) * RR_BLOCK + lane) * (size_t)P.tile_pitch; stride = 1; }
        else if (layout == 1) { p0 = base + ((size_t)j * P.n_blocks + b) * (size_t)P.tile_rows * RR_BLOCK + lane; stride = RR_BLOCK; }
        else { p0 = base + (size_t)c.t0 * ld + i; stride = ld; }
    };
    c.lat0 = nullptr;
    c.lstride = 1;
    if (HAS_LAT) place(P.tile_major, P.lateral[m], P.ldl, c.lat0, c.lstride);
    const double *o0 = nullptr;
    place(P.out_layout, P.out[m], P.ldo, o0, c.ostride);
    c.out0 = const_cast<double *>(o0);
    if (HAS_LAT) {
        // pull the first 16 rows of this item's lateral tile towards L2 while the dependency wait runs; the
        // fast path keeps prefetching 16 rows ahead of its time loop (a whole tile per warp up front would
        // not survive in L2 until it is used: the chip streams ~L2-size bytes during one item)
        const bool fast = (c.M.int_mask & 0x40) && P.K == 1 && MODE != RR_MODE_UNIT;
        const int pf_rows = fast ? min(c.rows, 16) : c.rows;
        if (P.tile_major >= 2) {
            prefetch_l2(c.lat0);                                   // this lane's first line (16 rows)
            if (!fast) for (int r = 16; r < c.rows; r += 16) prefetch_l2(c.lat0 + lat_row_off(P.tile_major, 1, r));
        } else if (P.tile_major) {
            const char *t = (const char *)(c.lat0 - lane);
            for (int l = lane; l < pf_rows * 2; l += 32) prefetch_l2(t + (size_t)l * 128);
        } else {
            const double *lt = c.lat0 - lane;
            const int64_t left = P.n - (int64_t)b * RR_BLOCK;
            const int row_bytes = (int)(left < RR_BLOCK ? left : RR_BLOCK) * 8;
            for (int r = lane; r < pf_rows; r += 32) {
                const char *a = (const char *)(lt + (size_t)r * c.lstride);
                prefetch_l2(a);
                if (row_bytes > 128) prefetch_l2(a + 128);
                prefetch_l2(a + row_bytes - 1);
            }
        }
    }
    // ---- dependencies ----
#ifdef RR_PROFILE
    const long long w0_ = clock64();
#endif
    // Everything the waits and the item need from the plan is requested BEFORE the first wait (the waits are
    // ordered asm statements: loads written after them cannot start earlier): the ids of the upstream blocks, this
    // lane's downstream block (ring reuse) and the ring rows of its own and its upstreams' exported series.
    c.dep_lo = dep_lo; c.dep_hi = dep_hi;
    int32_t dep_blk = -1;
    if (dep_lo + lane < dep_hi) dep_blk = __ldg(P.dep_idx + dep_lo + lane);
    int32_t down_blk = -1;
    c.ro_me = make_int2(0, 1);
    if (c.ex >= 0) {
        c.ro_me = __ldg(reinterpret_cast<const int2 *>(P.exp_ro) + c.ex);
        down_blk = __ldg(P.down + i) / RR_BLOCK;
    }
#pragma unroll
    for (int k = 0; k < RR_MAX_FAST_DEG; ++k) {
        c.ro_up[k] = make_int2(0, 1);
        c.up_u[k] = 0;
        if ((c.M.int_mask & 0x40) && k < c.deg) {
            if (P.direct) c.up_u[k] = __ldg(P.up_idx + c.e0 + k);
            else c.ro_up[k] = __ldg(reinterpret_cast<const int2 *>(P.edge_ro) + c.e0 + k);
        }
    }
    int32_t *done = P.done + (size_t)m * P.n_blocks;
    if (lane == 0 && j > 0) wait_ge(done + b, j);                       // own previous tile (acquire)
    if (dep_blk >= 0) wait_ge(done + dep_blk, j + 1);                   // upstream blocks, this tile
    for (int e = dep_lo + 32 + lane; e < dep_hi; e += 32) wait_ge(done + __ldg(P.dep_idx + e), j + 1);
    if (c.ex >= 0 && !P.direct && j >= c.ro_me.y) wait_ge(done + down_blk, j - c.ro_me.y + 1);   // exchange-ring reuse
    __syncwarp();
#ifdef RR_PROFILE
    c.prof_wait = (unsigned long long)(clock64() - w0_);
    c.prof_setup = 0;
#endif
    // first tile of a reference call: every member starts from the shared initial state and
    // (UNIT) q_ch = q_full = state (UnitMuskingum.py:78-79); later tiles / chunks continue
    // from the member's own running state.
    c.use_init = (j == 0) && P.first_call;
    c.q = 0.0;
    if (valid) c.q = c.use_init ? P.q_init[(size_t)m * P.q_init_stride + i] : P.q_state[m][i];
}

// ------------------------------------------------------------------------------------------------
// TMA-staged fast path (tile-major working arrays, K == 1, no in-block edges).
// One bulk async copy (cp.async.bulk, the 1-D TMA engine) brings the item's whole lateral tile
// [rows][32] -- contiguous in the tile-major layout -- into shared memory, and one more per upstream
// reach brings that reach's exchange row; all complete on one mbarrier per warp.  The time loop then
// runs out of shared memory, writes the clamped discharge over the lateral tile in place, and a bulk
// copy stores the tile back.  DRAM sees 16 KB / 640 B contiguous bursts instead of 32-byte sectors.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// returns false when the block needs more upstream row slots than the warp's shared-memory region has
template <int MODE>
__device__ __forceinline__ bool tma_item(const rr_route_params &P, const item_ctx &c, unsigned char *region,
                                         uint32_t &phase, bool &store_pending) {
    constexpr bool HAS_LAT = (MODE == RR_MODE_RAPID);
    const int lane = c.lane, j = c.j, TT = c.TT;
    const int cnt = c.deg;                         // every upstream is in another block
    int pre = cnt;                                 // inclusive warp scan of the slot counts
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL_MASK, pre, o);
        if (lane >= o) pre += v;
    }
    const int total = __shfl_sync(FULL_MASK, pre, 31);
    if (total > P.row_slots) return false;
    pre -= cnt;

    double *tile = reinterpret_cast<double *>(region);
    const int row_stride = P.raw_pitch + 2;        // doubles; +16 bytes keeps the lanes' rows on different banks
    double *rows = tile + (size_t)P.tile_rows * RR_BLOCK;
    const uint32_t bar = smem_u32(region + P.smem_region - 16);
    const uint32_t tile_bytes = (uint32_t)c.rows * RR_BLOCK * 8;

    if (store_pending) {                           // the previous item's tile store must have read the tile
        if (lane == 0) bulk_wait_read();
        store_pending = false;
    }
    __syncwarp();
    fence_async_smem();                            // generic-proxy accesses of the last item before new async writes
    const uint32_t my_bytes = (uint32_t)cnt * P.raw_pitch * 8 + ((HAS_LAT && lane == 0) ? tile_bytes : 0u);
    mbar_arrive_expect_tx(bar, my_bytes);
    if (HAS_LAT && lane == 0) bulk_load(smem_u32(tile), c.lat0, tile_bytes, bar);
    const double *rowp[RR_MAX_FAST_DEG];
#pragma unroll
    for (int k = 0; k < RR_MAX_FAST_DEG; ++k) {
        rowp[k] = rows;
        if (k < cnt) {
            const int2 ro = c.ro_up[k];
            const double *src = c.raw_m + ((size_t)ro.x + (size_t)(j % ro.y)) * P.raw_pitch;
            double *dst = rows + (size_t)(pre + k) * row_stride;
            bulk_load(smem_u32(dst), src, (uint32_t)P.raw_pitch * 8, bar);
            rowp[k] = dst;
        }
    }
    double q = c.q;
    double *myraw = nullptr;
    if (c.ex >= 0) {
        const int2 ro = c.ro_me;
        myraw = c.raw_m + ((size_t)ro.x + (size_t)(j % ro.y)) * P.raw_pitch;
        myraw[RAW_CARRY] = q;
    }
    const double c1 = c.c1, c2 = c.c2, c3 = c.c3, c4 = c.c4;
    mbar_wait(bar, phase);
    phase ^= 1u;

    double *col = tile + lane;
    for (int s = 0; s < TT; s += 4) {
        double r4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = RAW_CARRY + s + u;       // row entry before substep s+u; e + 1 is the entry after it
            double r = c3 * q;                                        // _numba_kernels.py:27-28 / :68-69
            if (HAS_LAT) r = fma(c4, col[(s + u) * RR_BLOCK], r);
#pragma unroll
            for (int k = 0; k < RR_MAX_FAST_DEG; ++k)
                if (k < cnt) r = fma(c2, rowp[k][e], r);              // :29-33 / :70-74, ascending upstream
#pragma unroll
            for (int k = 0; k < RR_MAX_FAST_DEG; ++k)
                if (k < cnt) r = fma(c1, rowp[k][e + 1], r);          // :36-39 / :75-78 (lhs_off = -c1)
            r4[u] = r;
            if (s + u < TT) {
                col[(s + u) * RR_BLOCK] = r > 0.0 ? r : 0.0;          // K == 1: mean == value; clamp :44-46 / :82-84
                q = r;
            }
        }
        if (myraw) st_sector(myraw + RAW_S0 + s, r4[0], r4[1], r4[2], r4[3]);
    }
    if (c.valid) P.q_state[c.m][c.i] = q;
    fence_async_smem();                            // make the tile visible to the async proxy
    __syncwarp();
    if (lane == 0) bulk_store(c.out0, smem_u32(tile), tile_bytes);
    store_pending = true;
    return true;
}

template <int MODE>
__device__ __forceinline__ void register_fast_item(const rr_route_params &P, const item_ctx &c) {
    if (P.tile_major >= 2) {
        switch (c.M.max_deg) {
            case 0: fast_item<MODE, 0, true>(P, c); break;
            case 1: fast_item<MODE, 1, true>(P, c); break;
            case 2: fast_item<MODE, 2, true>(P, c); break;
            case 3: fast_item<MODE, 3, true>(P, c); break;
            default: fast_item<MODE, 4, true>(P, c); break;
        }
    } else {
        switch (c.M.max_deg) {
            case 0: fast_item<MODE, 0, false>(P, c); break;
            case 1: fast_item<MODE, 1, false>(P, c); break;
            case 2: fast_item<MODE, 2, false>(P, c); break;
            case 3: fast_item<MODE, 3, false>(P, c); break;
            default: fast_item<MODE, 4, false>(P, c); break;
        }
    }
}

}  // namespace

template <int MODE>
__global__ void __launch_bounds__(256, RR_MIN_CTAS) rr_wavefront_kernel(const __grid_constant__ rr_route_params P) {
    constexpr bool UNIT = (MODE == RR_MODE_UNIT);
    const int lane = threadIdx.x & 31;
    int m, b, j, dep_lo, dep_hi;
    unsigned long long tk_next = 0, tk_end = 0;
    PROF_DECL
    for (;;) {
        const bool more = next_item(P, lane, m, b, j, dep_lo, dep_hi, tk_next, tk_end);
        PROF_MARK(0)
        if (!more) break;
        item_ctx c;
        open_item<MODE>(P, c, lane, m, b, j, dep_lo, dep_hi);
        PROF_MARK(1)
        // plan flag 0x40: no in-block edges and max in-degree <= RR_MAX_FAST_DEG
        if (!UNIT && P.K == 1 && (c.M.int_mask & 0x40)) register_fast_item<MODE>(P, c);
        else if (!UNIT && (c.M.int_mask & 0x40) && c.M.max_deg <= 2) {
            switch (c.M.max_deg) {
                case 0: fast_item_k<MODE, 0>(P, c); break;
                case 1: fast_item_k<MODE, 1>(P, c); break;
                default: fast_item_k<MODE, 2>(P, c); break;
            }
        } else if (UNIT && P.K == 1 && (c.M.int_mask & 0x40) && c.M.max_deg <= 2) {
            const bool vec = P.tile_major >= 2;
            switch (c.M.max_deg) {
                case 0: if (vec) unit_fast_item<0, true>(P, c); else unit_fast_item<0, false>(P, c); break;
                case 1: if (vec) unit_fast_item<1, true>(P, c); else unit_fast_item<1, false>(P, c); break;
                default: if (vec) unit_fast_item<2, true>(P, c); else unit_fast_item<2, false>(P, c); break;
            }
        } else general_item<MODE>(P, c);
        __syncwarp();
#ifdef RR_PROFILE
        if (!(c.M.int_mask & 0x40)) { prof_acc[7] += (unsigned long long)(clock64() - prof_t); prof_acc[5] += 1; }   // general-path items
#endif
        PROF_MARK(2)
#ifdef RR_PROFILE
        prof_acc[4] += c.prof_wait; prof_acc[6] += 1;
#endif
        if (P.jitter > 0) {   // stress tests: pseudo-random delay before publishing
            unsigned h = (unsigned)b * 2654435761u ^ (unsigned)j * 40503u ^ (unsigned)clock();
            __nanosleep((h ^ (h >> 13)) & ((1u << min(P.jitter, 14)) - 1u));
            __syncwarp();
        }
        if (lane == 0) st_release(P.done + (size_t)m * P.n_blocks + b, j + 1);
        __syncwarp();
        PROF_MARK(3)
    }
    PROF_FLUSH
}

// Persistent kernel of the TMA-staged path: 4 warps per CTA, one CTA per SM, each warp owns a private
// shared-memory region (tile + upstream row slots + its mbarrier).
template <int MODE>
__global__ void __launch_bounds__(128, 1) rr_wavefront_tma_kernel(const __grid_constant__ rr_route_params P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *region = smem + (size_t)warp * P.smem_region;
    if (lane == 0) mbar_init(smem_u32(region + P.smem_region - 16), 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
    __syncthreads();
    uint32_t phase = 0;
    bool store_pending = false;
    int m, b, j, dep_lo, dep_hi;
    unsigned long long tk_next = 0, tk_end = 0;
    while (next_item(P, lane, m, b, j, dep_lo, dep_hi, tk_next, tk_end)) {
        item_ctx c;
        open_item<MODE>(P, c, lane, m, b, j, dep_lo, dep_hi);
        bool done_item = false;
        if (c.M.int_mask & 0x40) {
            done_item = tma_item<MODE>(P, c, region, phase, store_pending);
            if (!done_item) { register_fast_item<MODE>(P, c); done_item = true; }
        }
        if (!done_item) general_item<MODE>(P, c);
        __syncwarp();
        if (lane == 0) st_release(P.done + (size_t)m * P.n_blocks + b, j + 1);
    }
    if (lane == 0) bulk_wait_all();
}

// Host-callable launcher (used by rr_api.cu).
cudaError_t rr_launch_wavefront(int mode, const rr_route_params &P, int grid, int block, cudaStream_t stream) {
    if (P.smem_region > 0) {   // TMA-staged path (tile-major working arrays, one substep per row)
        const size_t smem = (size_t)P.smem_region * 4;
        cudaError_t e;
        if (mode == RR_MODE_MUSKINGUM) {
            e = cudaFuncSetAttribute(rr_wavefront_tma_kernel<RR_MODE_MUSKINGUM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            rr_wavefront_tma_kernel<RR_MODE_MUSKINGUM><<<grid, 128, smem, stream>>>(P);
        } else if (mode == RR_MODE_RAPID) {
            e = cudaFuncSetAttribute(rr_wavefront_tma_kernel<RR_MODE_RAPID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            rr_wavefront_tma_kernel<RR_MODE_RAPID><<<grid, 128, smem, stream>>>(P);
        } else return cudaErrorInvalidValue;
        return cudaGetLastError();
    }
    switch (mode) {
        case RR_MODE_MUSKINGUM: rr_wavefront_kernel<RR_MODE_MUSKINGUM><<<grid, block, 0, stream>>>(P); break;
        case RR_MODE_RAPID: rr_wavefront_kernel<RR_MODE_RAPID><<<grid, block, 0, stream>>>(P); break;
        case RR_MODE_UNIT: rr_wavefront_kernel<RR_MODE_UNIT><<<grid, block, 0, stream>>>(P); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

int rr_wavefront_occupancy(int mode, int block) {
    int nb = 0;
    cudaError_t e;
    switch (mode) {
        case RR_MODE_MUSKINGUM: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, rr_wavefront_kernel<RR_MODE_MUSKINGUM>, block, 0); break;
        case RR_MODE_RAPID: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, rr_wavefront_kernel<RR_MODE_RAPID>, block, 0); break;
        default: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, rr_wavefront_kernel<RR_MODE_UNIT>, block, 0); break;
    }
    return e == cudaSuccess ? nb : -1;
}
