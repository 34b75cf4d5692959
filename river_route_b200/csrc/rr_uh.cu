// rr_uh.cu -- unit-hydrograph causal convolution for sm_100a.
//
// Replaces UnitHydrograph.convolve (river_route/uhkernels/UnitHydrograph.py:77-107):
//     out[t,b] = carry[t,b] (t < n_ks) + sum_{tau} kernel[tau,b] * lateral[t-tau,b]
// accumulated oldest contribution first, which is the order of the reference's exact
// direct form convolve_incrementally (:64-75; the FFT path agrees with it to 1e-12,
// tests/test_uhkernels.py:52-78).  Not a dense contraction: each basin has its own taps, so
// there is nothing for tensor cores; it is a register-blocked FIR (8 outputs x sliding input window per
// thread), one thread per basin, all arrays (time, basin) row-major so every warp access is a coalesced
// 256-byte row segment.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <string>

#include "rr_internal.h"

void rr_count_launch(int64_t k);

#define CK(call)                                                                \
    do {                                                                        \
        cudaError_t e_ = (call);                                                \
        if (e_ != cudaSuccess) {                                                \
            rr_set_error(std::string(#call) + ": " + cudaGetErrorString(e_));   \
            return 200;                                                         \
        }                                                                       \
    } while (0)

namespace {

// One thread per basin, TB outputs per iteration held in registers.  Taps are walked from the oldest
// input to the newest (tau descending), which is the accumulation order of convolve_incrementally; the
// TB inputs that pair with one tap slide by one position per tap, so each tap costs one new input load,
// one tap load and TB independent FMAs.  The tap loop is unrolled TB-fold with a rotating register
// window, so no register moves are needed.  Taps beyond n_ks are zero (fma(0, x, acc) == acc).
// grid.y splits time into chunks when there are few basins.
template <int TB>
__global__ void __launch_bounds__(128) uh_conv_kernel(int64_t n, int n_ks, int64_t T, int64_t chunk,
                                                      const double *__restrict__ lat, int64_t ldl,
                                                      const double *__restrict__ ker, int64_t ldk,
                                                      const double *__restrict__ state, int64_t lds,
                                                      double *__restrict__ out, int64_t ldo) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const int64_t tb = (int64_t)blockIdx.y * chunk;
    const int64_t te = min(T, tb + chunk);
    const int n_groups = (n_ks + TB - 1) / TB;
    const double *lb = lat + b;
    auto input = [&](int64_t t) -> double { return (t >= 0 && t < T) ? __ldg(lb + t * ldl) : 0.0; };
    for (int64_t t0 = tb; t0 < te; t0 += TB) {
        double acc[TB], w[TB];
#pragma unroll
        for (int u = 0; u < TB; ++u) {
            const int64_t t = t0 + u;
            acc[u] = (t < n_ks && t < te) ? state[t * lds + b] : 0.0;      // UnitHydrograph.py:100
        }
        int tau = n_groups * TB - 1;                                        // oldest (possibly padded) tap
#pragma unroll
        for (int u = 0; u < TB; ++u) w[u] = input(t0 + u - tau);
        for (int g = 0; g < n_groups; ++g) {
#pragma unroll
            for (int p = 0; p < TB; ++p, --tau) {
                const double kt = tau < n_ks ? __ldg(ker + (int64_t)tau * ldk + b) : 0.0;
#pragma unroll
                for (int u = 0; u < TB; ++u) acc[u] = fma(kt, w[(u + p) % TB], acc[u]);
                w[p % TB] = input(t0 + TB - tau);                           // newest input of the next tap
            }
        }
#pragma unroll
        for (int u = 0; u < TB; ++u)
            if (t0 + u < te) out[(t0 + u) * ldo + b] = acc[u];
    }
}

// Same algorithm with the taps held in registers (NG groups of TB taps, n_ks <= NG * TB): the tap loop is fully
// unrolled, so the only loads in the inner loop are the sliding inputs.
template <int TB, int NG>
__global__ void __launch_bounds__(128) uh_conv_kernel_reg(int64_t n, int n_ks, int64_t T, int64_t chunk,
                                                          const double *__restrict__ lat, int64_t ldl,
                                                          const double *__restrict__ ker, int64_t ldk,
                                                          const double *__restrict__ state, int64_t lds,
                                                          double *__restrict__ out, int64_t ldo) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const int64_t tb = (int64_t)blockIdx.y * chunk;
    const int64_t te = min(T, tb + chunk);
    constexpr int NT = NG * TB;
    double kq[NT];
#pragma unroll
    for (int k = 0; k < NT; ++k) kq[k] = k < n_ks ? __ldg(ker + (int64_t)k * ldk + b) : 0.0;
    const double *lb = lat + b;
    auto input = [&](int64_t t) -> double { return (t >= 0 && t < T) ? __ldg(lb + t * ldl) : 0.0; };
    for (int64_t t0 = tb; t0 < te; t0 += TB) {
        double acc[TB], w[TB];
#pragma unroll
        for (int u = 0; u < TB; ++u) {
            const int64_t t = t0 + u;
            acc[u] = (t < n_ks && t < te) ? state[t * lds + b] : 0.0;      // UnitHydrograph.py:100
        }
#pragma unroll
        for (int u = 0; u < TB; ++u) w[u] = input(t0 + u - (NT - 1));
#pragma unroll
        for (int g = 0; g < NG; ++g) {
#pragma unroll
            for (int p = 0; p < TB; ++p) {
                const int tau = NT - 1 - (g * TB + p);                      // oldest tap first
#pragma unroll
                for (int u = 0; u < TB; ++u) acc[u] = fma(kq[tau], w[(u + p) % TB], acc[u]);
                w[p % TB] = input(t0 + TB - tau);
            }
        }
#pragma unroll
        for (int u = 0; u < TB; ++u)
            if (t0 + u < te) out[(t0 + u) * ldo + b] = acc[u];
    }
}


// Long kernels (more than 16 taps): the taps live in shared memory ([tap][thread], read back by the thread that wrote
// them, so no barrier is needed) and each iteration produces 16 outputs.  Round 1 kept up to 48 taps in registers
// (168 registers, 12 warps per SM) with 8 outputs per iteration, and every iteration re-loaded its whole input window:
// 56 loads of 8 bytes per 8 outputs, i.e. 56 B of L1/L2 traffic per output for 16 B of algorithmic traffic.  With 16
// outputs per iteration the window costs 63 loads per 16 outputs, the tap costs one LDS per 16 DFMA, and ~90 registers
// give 16 warps per SM.  Same accumulation order: old state first, then inputs from the oldest to the newest.
// SAFE: the iteration's whole input window lies inside [0, T) -- no per-load bounds predicates.
template <int TB>
__global__ void __launch_bounds__(128, 4) uh_conv_kernel_smem(int64_t n, int n_ks, int n_groups, int64_t T, int64_t chunk,
                                                              const double *__restrict__ lat, int64_t ldl,
                                                              const double *__restrict__ ker, int64_t ldk,
                                                              const double *__restrict__ state, int64_t lds,
                                                              double *__restrict__ out, int64_t ldo) {
    extern __shared__ double sk[];                         // [n_groups * TB][128]
    const int tid = threadIdx.x;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + tid;
    if (b >= n) return;
    const int NT = n_groups * TB;
    for (int k = 0; k < NT; ++k) sk[k * 128 + tid] = k < n_ks ? __ldg(ker + (int64_t)k * ldk + b) : 0.0;
    const int64_t tb = (int64_t)blockIdx.y * chunk;
    const int64_t te = min(T, tb + chunk);
    const double *lb = lat + b;
    for (int64_t t0 = tb; t0 < te; t0 += TB) {
        double acc[TB], w[TB];
#pragma unroll
        for (int u = 0; u < TB; ++u) {
            const int64_t t = t0 + u;
            acc[u] = (t < n_ks && t < te) ? state[t * lds + b] : 0.0;      // UnitHydrograph.py:100
        }
        int tau = NT - 1;                                                   // oldest (possibly padded) tap
        const bool safe = t0 - tau >= 0 && t0 + TB <= T;
        if (safe) {
            const double *p0 = lb + (t0 - tau) * ldl;                       // input of output 0 under the oldest tap
#pragma unroll
            for (int u = 0; u < TB; ++u) w[u] = __ldg(p0 + u * ldl);
            const double *pn = p0 + TB * ldl;                               // next input to enter the window
            for (int g = 0; g < n_groups; ++g) {
#pragma unroll
                for (int p = 0; p < TB; ++p, --tau) {
                    const double kt = sk[tau * 128 + tid];
#pragma unroll
                    for (int u = 0; u < TB; ++u) acc[u] = fma(kt, w[(u + p) % TB], acc[u]);
                    if (tau > 0) w[p] = __ldg(pn);                          // t0 + TB - tau < t0 + TB <= T
                    pn += ldl;
                }
            }
        } else {
            auto input = [&](int64_t t) -> double { return (t >= 0 && t < T) ? __ldg(lb + t * ldl) : 0.0; };
#pragma unroll
            for (int u = 0; u < TB; ++u) w[u] = input(t0 + u - tau);
            for (int g = 0; g < n_groups; ++g) {
#pragma unroll
                for (int p = 0; p < TB; ++p, --tau) {
                    const double kt = sk[tau * 128 + tid];
#pragma unroll
                    for (int u = 0; u < TB; ++u) acc[u] = fma(kt, w[(u + p) % TB], acc[u]);
                    w[p] = input(t0 + TB - tau);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < TB; ++u)
            if (t0 + u < te) out[(t0 + u) * ldo + b] = acc[u];
    }
}

// Carry-over state: rows 0..n_ks-2 = the full convolution at times T..T+n_ks-2, last row = 0
// (UnitHydrograph.py:103-105).  Reads of the old state run ahead of the writes (row T+j > j).
__global__ void __launch_bounds__(128) uh_state_kernel(int64_t n, int n_ks, int64_t T,
                                                       const double *__restrict__ lat, int64_t ldl,
                                                       const double *__restrict__ ker, int64_t ldk,
                                                       double *state, int64_t lds) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    for (int j = 0; j < n_ks - 1; ++j) {
        const int64_t t = T + j;
        double acc = t < n_ks ? state[t * lds + b] : 0.0;
        const int64_t s_lo = t - n_ks + 1 > 0 ? t - n_ks + 1 : 0;
        for (int64_t s = s_lo; s < T; ++s)
            acc = fma(__ldg(ker + (t - s) * ldk + b), __ldg(lat + s * ldl + b), acc);
        state[(int64_t)j * lds + b] = acc;
    }
    state[(int64_t)(n_ks - 1) * lds + b] = 0.0;
}

// The same with the taps and the last NT - 1 inputs in registers (n_ks <= NT = NG * 8): every kernel row and every
// input row is read once instead of once per (state row, tap) pair.  Same accumulation order: old state first,
// then inputs from the oldest to the newest.
template <int NG>
__global__ void __launch_bounds__(128) uh_state_kernel_reg(int64_t n, int n_ks, int64_t T,
                                                           const double *__restrict__ lat, int64_t ldl,
                                                           const double *__restrict__ ker, int64_t ldk,
                                                           double *state, int64_t lds) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    constexpr int NT = NG * 8;
    double kq[NT], x[NT];   // x[i] = input at time T - NT + i (zero before the call started)
#pragma unroll
    for (int k = 0; k < NT; ++k) {
        kq[k] = k < n_ks ? __ldg(ker + (int64_t)k * ldk + b) : 0.0;
        const int64_t t = T - NT + k;
        x[k] = t >= 0 ? __ldg(lat + t * ldl + b) : 0.0;
    }
#pragma unroll
    for (int j = 0; j < NT - 1; ++j) {
        if (j < n_ks - 1) {
            // an old state row that is still ahead of the call's end (only when T < n_ks); row T + j is overwritten
            // later than it is read here because T >= 1
            const int64_t t = T + j;
            double acc = t < n_ks ? state[t * lds + b] : 0.0;
            // inputs s = T - NT + i, tap = T + j - s = NT + j - i; taps >= n_ks are zero (kq padded), taps <= NT - 1
#pragma unroll
            for (int i = j + 1; i < NT; ++i) acc = fma(kq[NT + j - i], x[i], acc);
            state[(int64_t)j * lds + b] = acc;
        }
    }
    state[(int64_t)(n_ks - 1) * lds + b] = 0.0;
}

}  // namespace

extern "C" int rr_uh_convolve_dev(int64_t n, int64_t n_ks, int64_t T, const double *lateral, int64_t ldl,
                                  const double *kernel, int64_t ldk, double *state, int64_t lds, double *out,
                                  int64_t ldo, void *stream_) {
    if (n <= 0 || n_ks <= 0 || T <= 0) { rr_set_error("n, n_ks and T must be positive"); return 100; }
    if (!lateral || !kernel || !state || !out) { rr_set_error("null argument"); return 100; }
    if (n_ks > 0x7fffffff) { rr_set_error("kernel too long"); return 100; }
    cudaStream_t stream = (cudaStream_t)stream_;
    const int threads = 128;
    const unsigned gx = (unsigned)((n + threads - 1) / threads);
    // enough time chunks to fill the machine when there are few basins; a chunk re-reads NK-1 rows
    int sms = 148;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t want_y = std::max<int64_t>(1, ((int64_t)sms * 16 + gx - 1) / gx);
    int64_t chunk = std::max<int64_t>(std::max<int64_t>(8 * n_ks, 64), (T + want_y - 1) / want_y);
    chunk = std::min<int64_t>(((chunk + 7) / 8) * 8, std::max<int64_t>(T, 1));
    const unsigned gy = (unsigned)((T + chunk - 1) / chunk);
    dim3 grid(gx, gy);
    const int nk = (int)n_ks;
    const int groups16 = (nk + 15) / 16;
    const size_t smem_taps = (size_t)groups16 * 16 * 128 * sizeof(double);
    static const bool reg_only = getenv("RR_UH_REGISTERS") != nullptr;          // A/B switch for measurements
    if (nk > 16 && smem_taps <= 192 * 1024 && !reg_only) {
        int64_t chunk16 = std::min<int64_t>(((chunk + 15) / 16) * 16, std::max<int64_t>(T, 1));
        dim3 grid16(gx, (unsigned)((T + chunk16 - 1) / chunk16));
        CK(cudaFuncSetAttribute(uh_conv_kernel_smem<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_taps));
        uh_conv_kernel_smem<16><<<grid16, threads, smem_taps, stream>>>(n, nk, groups16, T, chunk16, lateral, ldl, kernel, ldk, state, lds,
                                                                      out, ldo);
    } else
#define RR_UH_REG(NG) uh_conv_kernel_reg<8, NG><<<grid, threads, 0, stream>>>(n, nk, T, chunk, lateral, ldl, kernel, ldk, state, lds, out, ldo)
    switch ((nk + 7) / 8) {
        case 1: RR_UH_REG(1); break;
        case 2: RR_UH_REG(2); break;
        case 3: RR_UH_REG(3); break;
        case 4: RR_UH_REG(4); break;
        case 5: RR_UH_REG(5); break;
        case 6: RR_UH_REG(6); break;
        default: uh_conv_kernel<8><<<grid, threads, 0, stream>>>(n, nk, T, chunk, lateral, ldl, kernel, ldk, state, lds, out, ldo);
    }
#undef RR_UH_REG
    CK(cudaGetLastError());
#define RR_UH_STATE(NG) uh_state_kernel_reg<NG><<<gx, threads, 0, stream>>>(n, nk, T, lateral, ldl, kernel, ldk, state, lds)
    switch ((nk + 7) / 8) {
        case 1: RR_UH_STATE(1); break;
        case 2: RR_UH_STATE(2); break;
        case 3: RR_UH_STATE(3); break;
        case 4: RR_UH_STATE(4); break;
        case 5: RR_UH_STATE(5); break;
        case 6: RR_UH_STATE(6); break;
        default: uh_state_kernel<<<gx, threads, 0, stream>>>(n, nk, T, lateral, ldl, kernel, ldk, state, lds);
    }
#undef RR_UH_STATE
    CK(cudaGetLastError());
    rr_count_launch(2);
    return 0;
}

extern "C" int rr_uh_convolve_host(int64_t n, int64_t n_ks, int64_t T, const double *lateral, int64_t ldl,
                                   const double *kernel, int64_t ldk, double *state, int64_t lds, double *out,
                                   int64_t ldo) {
    if (n <= 0 || n_ks <= 0 || T <= 0) { rr_set_error("n, n_ks and T must be positive"); return 100; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        rr_set_error("no CUDA device available: librr_b200 has no CPU fallback");
        return 201;
    }
    const int64_t ldd = ((n + 31) / 32) * 32;
    double *d_lat = nullptr, *d_ker = nullptr, *d_state = nullptr, *d_out = nullptr;
    CK(cudaMalloc((void **)&d_lat, sizeof(double) * (size_t)T * ldd));
    CK(cudaMalloc((void **)&d_out, sizeof(double) * (size_t)T * ldd));
    CK(cudaMalloc((void **)&d_ker, sizeof(double) * (size_t)n_ks * ldd));
    CK(cudaMalloc((void **)&d_state, sizeof(double) * (size_t)n_ks * ldd));
    int rc = 0;
    auto fail = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && !rc) { rr_set_error(std::string(what) + ": " + cudaGetErrorString(e)); rc = 200; }
    };
    fail(cudaMemcpy2D(d_lat, ldd * 8, lateral, ldl * 8, n * 8, T, cudaMemcpyHostToDevice), "H2D lateral");
    fail(cudaMemcpy2D(d_ker, ldd * 8, kernel, ldk * 8, n * 8, n_ks, cudaMemcpyHostToDevice), "H2D kernel");
    fail(cudaMemcpy2D(d_state, ldd * 8, state, lds * 8, n * 8, n_ks, cudaMemcpyHostToDevice), "H2D state");
    if (!rc) rc = rr_uh_convolve_dev(n, n_ks, T, d_lat, ldd, d_ker, ldd, d_state, ldd, d_out, ldd, nullptr);
    if (!rc) {
        fail(cudaMemcpy2D(out, ldo * 8, d_out, ldd * 8, n * 8, T, cudaMemcpyDeviceToHost), "D2H out");
        fail(cudaMemcpy2D(state, lds * 8, d_state, ldd * 8, n * 8, n_ks, cudaMemcpyDeviceToHost), "D2H state");
    }
    cudaFree(d_lat); cudaFree(d_out); cudaFree(d_ker); cudaFree(d_state);
    return rc;
}
