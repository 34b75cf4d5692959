// Device-side helpers shared by the wavefront kernels (rr_route.cu, rr_direct.cu): flag protocol, 256-bit memory
// operations, ticket decode.
#pragma once
#include <cuda_runtime.h>

#include "rr_route.cuh"

#ifndef RR_SPIN_NS0
#define RR_SPIN_NS0 64
#endif
#ifndef RR_SPIN_NSMAX
#define RR_SPIN_NSMAX 1024
#endif
#define RR_FULL_MASK 0xffffffffu

namespace rrdev {

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

// stress tests: a delay of 0 .. 2^bits ns that differs per (block, tile, site, SM clock)
__device__ __forceinline__ void jitter_delay(int jitter, int b, int j, int site) {
    if (jitter > 0) {
        unsigned h = (unsigned)b * 2654435761u ^ (unsigned)j * 40503u ^ (unsigned)site * 69069u ^ (unsigned)clock();
        h ^= h >> 13;
        __nanosleep(h & ((1u << min(jitter, 14)) - 1u));
    }
}

struct d4 { double a, b, c, d; };
// one aligned 32-byte sector written by another SM earlier in this launch: coherent load, ordered after the acquire;
// L2 fetches the whole 128-byte line (this and the next three sectors of the series) in one DRAM burst
__device__ __forceinline__ d4 ld_sector(const double *p) {
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

// ticket -> (member, block, tile, dependency range); false when the tickets are exhausted
__device__ __forceinline__ bool next_ticket(const rr_route_params &P, int lane, int &m, int &b, int &j, int &dep_lo, int &dep_hi) {
    unsigned long long t0 = 0;
    if (lane == 0) t0 = atomicAdd(P.ticket, 1ull);
    const unsigned long long tk = __shfl_sync(RR_FULL_MASK, t0, 0);
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

}  // namespace rrdev
