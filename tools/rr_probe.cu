// rr_probe.cu -- memory-system probe for the access patterns of the routing kernel (diagnostic, not on the
// product path).  The wavefront kernel moves its bytes in 32-byte sectors: each lane streams its own 512-byte
// series (a reach's lateral rows, an upstream reach's exchange row) with one 256-bit load per four time steps, and
// writes its own series the same way.  A plain copy is not the right ceiling for that, so this measures what HBM
// delivers for exactly this shape:
//   mode 0  coalesced copy: consecutive lanes touch consecutive 32-byte sectors (the torch copy of
//           MEASURED_PEAKS.json, expressed with the same 256-bit instructions)
//   mode 1  per-lane series, rows in order: lane l of item i streams row 32 i + l (the lateral tile pattern)
//   mode 2  per-lane series, rows scattered: lane l streams row perm[32 i + l] (the exchange-row pattern: upstream
//           reaches of neighbouring lanes live anywhere in the ring buffer) and writes row 32 i + l
//   mode 3 / 4  as 1 / 2 with the L2::128B prefetch hint on the loads (what the routing kernel uses)
//   mode 5  coalesced reads, per-lane series writes      (cost of the write shape alone)
//   mode 6  per-lane series reads (hinted), coalesced writes   (cost of the read shape alone)
// Every mode reads and writes `rows x row_doubles x 8` bytes once; the caller gets the kernel time.
#include <cuda_runtime.h>

#include <string>

#include <cstdint>

// diagnostic library of its own (tools/librr_probe.so, built by `make -C river_route_b200/csrc probe`): not part of
// the product library or its C ABI
static thread_local std::string g_probe_err;
static void rr_set_error(const std::string &msg) { g_probe_err = msg; }
extern "C" const char *rr_probe_last_error(void) { return g_probe_err.c_str(); }

namespace {

__device__ __forceinline__ void ld256(const double *p, double &a, double &b, double &c, double &d) {
    asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void ld256h(const double *p, double &a, double &b, double &c, double &d) {
    asm volatile("ld.global.nc.L2::128B.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void st256(double *p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(256) probe_kernel(const double *__restrict__ src, double *__restrict__ dst,
                                                    const int32_t *__restrict__ perm, int64_t rows, int row_doubles) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int groups = row_doubles / 4;
    for (int64_t item = warp; item * 32 < rows; item += n_warps) {
        const int64_t row = item * 32 + lane;
        if (MODE == 0) {
            // the item's 32 rows as one contiguous block; lane l takes sectors l, l + 32, ...
            const double *s = src + item * 32 * (int64_t)row_doubles;
            double *d = dst + item * 32 * (int64_t)row_doubles;
            const int64_t left = rows - item * 32;
            const int sectors = groups * (int)(left < 32 ? left : 32);
            for (int g = lane; g < sectors; g += 32 * 2) {
                double a0, b0, c0, d0, a1 = 0, b1 = 0, c1 = 0, d1 = 0;
                ld256(s + (int64_t)g * 4, a0, b0, c0, d0);
                const bool two = g + 32 < sectors;
                if (two) ld256(s + (int64_t)(g + 32) * 4, a1, b1, c1, d1);
                st256(d + (int64_t)g * 4, a0, b0, c0, d0);
                if (two) st256(d + (int64_t)(g + 32) * 4, a1, b1, c1, d1);
            }
        } else if (MODE == 5 || MODE == 6) {
            // one side coalesced over the item's contiguous 32-row block, the other side one series per lane; the item
            // is staged through registers group by group: lane l owns sector (g, l) of the block on the coalesced side
            // and group g of row l on the series side -- same bytes, different shape
            const int64_t left = rows - item * 32;
            if (left < 32) continue;
            const double *sb = src + item * 32 * (int64_t)row_doubles;
            double *db = dst + item * 32 * (int64_t)row_doubles;
            for (int g = 0; g < groups; ++g) {
                double a, b, c, d;
                if (MODE == 5) {
                    ld256(sb + ((int64_t)g * 32 + lane) * 4, a, b, c, d);
                    st256(db + (int64_t)lane * row_doubles + g * 4, a, b, c, d);
                } else {
                    ld256h(sb + (int64_t)lane * row_doubles + g * 4, a, b, c, d);
                    st256(db + ((int64_t)g * 32 + lane) * 4, a, b, c, d);
                }
            }
        } else if (row < rows) {
            const int64_t from = (MODE == 2 || MODE == 4) ? (int64_t)__ldg(perm + row) : row;
            const double *s = src + from * (int64_t)row_doubles;
            double *d = dst + row * (int64_t)row_doubles;
            for (int g = 0; g < groups; g += 2) {   // two groups in flight, like the routing kernel's pipeline
                double a0, b0, c0, d0, a1 = 0, b1 = 0, c1 = 0, d1 = 0;
                if (MODE >= 3) ld256h(s + g * 4, a0, b0, c0, d0); else ld256(s + g * 4, a0, b0, c0, d0);
                if (g + 1 < groups) { if (MODE >= 3) ld256h(s + (g + 1) * 4, a1, b1, c1, d1); else ld256(s + (g + 1) * 4, a1, b1, c1, d1); }
                st256(d + g * 4, a0, b0, c0, d0);
                if (g + 1 < groups) st256(d + (g + 1) * 4, a1, b1, c1, d1);
            }
        }
    }
}

}  // namespace

#define CK(call)                                                                \
    do {                                                                        \
        cudaError_t e_ = (call);                                                \
        if (e_ != cudaSuccess) {                                                \
            rr_set_error(std::string(#call) + ": " + cudaGetErrorString(e_));   \
            return 200;                                                         \
        }                                                                       \
    } while (0)

extern "C" int rr_probe_sector_bandwidth(int mode, int64_t rows, int32_t row_doubles, const double *src, double *dst,
                                         const int32_t *perm, int32_t reps, double *ms_best) {
    if (mode < 0 || mode > 6 || rows <= 0 || row_doubles <= 0 || row_doubles % 4 || !src || !dst || !ms_best ||
        ((mode == 2 || mode == 4) && !perm) || reps < 1) {
        rr_set_error("bad probe argument");
        return 100;
    }
    int dev = 0, sms = 148;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    double best = 1e300;
    for (int r = 0; r < reps + 1; ++r) {
        CK(cudaEventRecord(a));
        const int grid = sms * 8;
        switch (mode) {
            case 0: probe_kernel<0><<<grid, 256>>>(src, dst, perm, rows, row_doubles); break;
            case 1: probe_kernel<1><<<grid, 256>>>(src, dst, perm, rows, row_doubles); break;
            case 2: probe_kernel<2><<<grid, 256>>>(src, dst, perm, rows, row_doubles); break;
            case 3: probe_kernel<3><<<grid, 256>>>(src, dst, perm, rows, row_doubles); break;
            case 4: probe_kernel<4><<<grid, 256>>>(src, dst, perm, rows, row_doubles); break;
            case 5: probe_kernel<5><<<grid, 256>>>(src, dst, perm, rows, row_doubles); break;
            default: probe_kernel<6><<<grid, 256>>>(src, dst, perm, rows, row_doubles); break;
        }
        CK(cudaGetLastError());
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (r > 0 && ms < best) best = ms;   // first launch is the warm-up
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *ms_best = best;
    return 0;
}
