// rr_spmm.cu -- grid-runoff -> catchment lateral inflow for sm_100a.
//
// Replaces the SpMM core and the in-place tail of runoff_to_qlateral
// (river_route/runoff.py:292-337): y[t,r] = sum_j w[j] * x[t, col[j]] over CSR row r in stored
// (ascending column) order -- scipy's csr_matvecs row-wise axpy order -- then cumulative ->
// incremental, optional clip at zero, NaN -> 0 and optional multiplication by catchment area.
// Memory-bound CSR SpMM with time as the dense dimension: one thread per river so the (T, n)
// output rows are written as coalesced 256-byte segments; grid cells of neighbouring catchments
// are neighbours in the grid row, which stays L2 resident (a 0.25 degree global row is 4 MB).
#include <cuda_runtime.h>

#include <algorithm>
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

constexpr int FAST_NNZ = 8;   // row entries cached in registers
constexpr int TSTEP = 4;      // time steps in flight per thread

template <typename XT>
__device__ __forceinline__ double row_dot(const XT *__restrict__ xrow, const int32_t *__restrict__ indices,
                                          const double *__restrict__ w, int p0, int nnz, const int32_t (&col)[FAST_NNZ],
                                          const double (&wt)[FAST_NNZ]) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < FAST_NNZ; ++k)
        if (k < nnz) acc = fma(wt[k], (double)__ldg(xrow + col[k]), acc);
    for (int k = FAST_NNZ; k < nnz; ++k)
        acc = fma(__ldg(w + p0 + k), (double)__ldg(xrow + __ldg(indices + p0 + k)), acc);
    return acc;
}

template <typename XT>
__global__ void __launch_bounds__(128) weights_kernel(int64_t n_rivers, int64_t T, int64_t chunk,
                                                      const int32_t *__restrict__ indptr,
                                                      const int32_t *__restrict__ indices,
                                                      const double *__restrict__ w, const XT *__restrict__ x,
                                                      int64_t ldx, double *__restrict__ y, int64_t ldy,
                                                      int cumulative, int force_positive,
                                                      const double *__restrict__ area) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rivers) return;
    const int64_t tb = (int64_t)blockIdx.y * chunk;
    const int64_t te = min(T, tb + chunk);
    const int p0 = __ldg(indptr + r);
    const int nnz = __ldg(indptr + r + 1) - p0;
    int32_t col[FAST_NNZ];
    double wt[FAST_NNZ];
#pragma unroll
    for (int k = 0; k < FAST_NNZ; ++k) {
        col[k] = k < nnz ? __ldg(indices + p0 + k) : 0;
        wt[k] = k < nnz ? __ldg(w + p0 + k) : 0.0;
    }
    const double a = area ? __ldg(area + r) : 1.0;
    // cumulative input: the chunk needs the aggregated value of the step before it (runoff.py:310-312)
    double prev = 0.0;
    if (cumulative && tb > 0) prev = row_dot(x + (tb - 1) * ldx, indices, w, p0, nnz, col, wt);
    for (int64_t t = tb; t < te; t += TSTEP) {
        double v[TSTEP];
#pragma unroll
        for (int u = 0; u < TSTEP; ++u)
            v[u] = (t + u < te) ? row_dot(x + (t + u) * ldx, indices, w, p0, nnz, col, wt) : 0.0;
#pragma unroll
        for (int u = 0; u < TSTEP; ++u) {
            if (t + u < te) {
                double q = v[u];
                if (cumulative) { if (t + u > 0) q = v[u] - prev; prev = v[u]; }
                if (force_positive && q < 0.0) q = 0.0;   // :313-314
                if (q != q) q = 0.0;                      // :331-333
                if (area) q *= a;                         // :335-336
                y[(t + u) * ldy + r] = q;
            }
        }
    }
}

}  // namespace

extern "C" int rr_weights_transform_dev(int64_t n_rivers, int64_t T, const int32_t *indptr, const int32_t *indices,
                                        const double *w, const void *x, int x_is_f32, int64_t ldx, double *y,
                                        int64_t ldy, int cumulative, int force_positive, const double *area,
                                        void *stream_) {
    if (n_rivers <= 0 || T <= 0) { rr_set_error("n_rivers and T must be positive"); return 100; }
    if (!indptr || !indices || !w || !x || !y) { rr_set_error("null argument"); return 100; }
    cudaStream_t stream = (cudaStream_t)stream_;
    const int threads = 128;
    const unsigned gx = (unsigned)((n_rivers + threads - 1) / threads);
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want_y = std::max<int64_t>(1, ((int64_t)sms * 16 + gx - 1) / gx);
    int64_t chunk = std::max<int64_t>(32, (T + want_y - 1) / want_y);
    chunk = std::min<int64_t>(((chunk + TSTEP - 1) / TSTEP) * TSTEP, std::max<int64_t>(T, 1));
    dim3 grid(gx, (unsigned)((T + chunk - 1) / chunk));
    if (x_is_f32)
        weights_kernel<float><<<grid, threads, 0, stream>>>(n_rivers, T, chunk, indptr, indices, w, (const float *)x,
                                                            ldx, y, ldy, cumulative, force_positive, area);
    else
        weights_kernel<double><<<grid, threads, 0, stream>>>(n_rivers, T, chunk, indptr, indices, w,
                                                             (const double *)x, ldx, y, ldy, cumulative,
                                                             force_positive, area);
    CK(cudaGetLastError());
    rr_count_launch(1);
    return 0;
}

extern "C" int rr_weights_transform_host(int64_t n_rivers, int64_t n_points, int64_t T, const int32_t *indptr,
                                         const int32_t *indices, const double *w, const void *x, int x_is_f32,
                                         int64_t ldx, double *y, int64_t ldy, int cumulative, int force_positive,
                                         const double *area) {
    if (n_rivers <= 0 || T <= 0 || n_points <= 0) { rr_set_error("sizes must be positive"); return 100; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        rr_set_error("no CUDA device available: librr_b200 has no CPU fallback");
        return 201;
    }
    const size_t es = x_is_f32 ? 4 : 8;
    const int64_t nnz = indptr[n_rivers];
    const int64_t ldyd = ((n_rivers + 31) / 32) * 32;
    int32_t *d_ptr = nullptr, *d_idx = nullptr;
    double *d_w = nullptr, *d_y = nullptr, *d_area = nullptr;
    void *d_x = nullptr;
    CK(cudaMalloc((void **)&d_ptr, sizeof(int32_t) * (size_t)(n_rivers + 1)));
    CK(cudaMalloc((void **)&d_idx, sizeof(int32_t) * (size_t)std::max<int64_t>(nnz, 1)));
    CK(cudaMalloc((void **)&d_w, sizeof(double) * (size_t)std::max<int64_t>(nnz, 1)));
    CK(cudaMalloc((void **)&d_x, es * (size_t)T * n_points));
    CK(cudaMalloc((void **)&d_y, sizeof(double) * (size_t)T * ldyd));
    if (area) CK(cudaMalloc((void **)&d_area, sizeof(double) * (size_t)n_rivers));
    int rc = 0;
    auto fail = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && !rc) { rr_set_error(std::string(what) + ": " + cudaGetErrorString(e)); rc = 200; }
    };
    fail(cudaMemcpy(d_ptr, indptr, sizeof(int32_t) * (size_t)(n_rivers + 1), cudaMemcpyHostToDevice), "H2D indptr");
    fail(cudaMemcpy(d_idx, indices, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice), "H2D indices");
    fail(cudaMemcpy(d_w, w, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice), "H2D weights");
    fail(cudaMemcpy2D(d_x, n_points * es, x, ldx * es, n_points * es, T, cudaMemcpyHostToDevice), "H2D runoff");
    if (area) fail(cudaMemcpy(d_area, area, sizeof(double) * (size_t)n_rivers, cudaMemcpyHostToDevice), "H2D area");
    if (!rc)
        rc = rr_weights_transform_dev(n_rivers, T, d_ptr, d_idx, d_w, d_x, x_is_f32, n_points, d_y, ldyd, cumulative,
                                      force_positive, d_area, nullptr);
    if (!rc) fail(cudaMemcpy2D(y, ldy * 8, d_y, ldyd * 8, n_rivers * 8, T, cudaMemcpyDeviceToHost), "D2H qlateral");
    cudaFree(d_ptr); cudaFree(d_idx); cudaFree(d_w); cudaFree(d_x); cudaFree(d_y);
    if (d_area) cudaFree(d_area);
    return rc;
}
