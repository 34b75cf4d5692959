// How long does one warp take to fetch "one group" of upstream entries from L2 -- 32 lanes x NUP lines x 4 sectors of 32 B,
// every lane its own 128-byte line(s), lines 512 B apart -- with weak, .cg and strong (relaxed.gpu) 256-bit loads?
// The buffer is written by a previous kernel (resident in L2, not in L1).  Prints ns per group fetch for 1 warp alone and
// for 16 warps per SM on all SMs fetching at once.
#include <cstdio>
#include <cuda_runtime.h>
struct d4 { double a, b, c, d; };
template <int KIND>
__device__ __forceinline__ d4 ld(const double *p) {
    d4 v;
    if (KIND == 0) asm volatile("ld.global.L2::128B.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.d) : "l"(p) : "memory");
    if (KIND == 1) asm volatile("ld.global.cg.L2::128B.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.d) : "l"(p) : "memory");
    if (KIND == 2) asm volatile("ld.relaxed.gpu.global.L2::128B.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.d) : "l"(p) : "memory");
    if (KIND == 3) {   // two 128-bit strong loads
        asm volatile("ld.relaxed.gpu.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.a), "=d"(v.b) : "l"(p) : "memory");
        asm volatile("ld.relaxed.gpu.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.c), "=d"(v.d) : "l"(p + 2) : "memory");
    }
    if (KIND == 4) asm volatile("ld.volatile.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.d) : "l"(p) : "memory");
    return v;
}
__global__ void fill(double *buf, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) buf[i] = (double)i;
}
template <int KIND, int NUP>
__global__ void fetch(const double *buf, size_t n_groups_per_warp, int iters, unsigned long long *out, double *sink) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    // each warp walks its own region: group g of the warp = 32 lanes x NUP lines x 64 doubles pitch
    const double *base = buf + (size_t)warp * n_groups_per_warp * 32 * NUP * 64;
    double acc = 0;
    unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const double *g = base + (size_t)(it % n_groups_per_warp) * 32 * NUP * 64;
        d4 v[NUP][4];
#pragma unroll
        for (int k = 0; k < NUP; ++k)
#pragma unroll
            for (int s = 0; s < 4; ++s) v[k][s] = ld<KIND>(g + ((size_t)k * 32 + lane) * 64 + 4 * s);
#pragma unroll
        for (int k = 0; k < NUP; ++k)
#pragma unroll
            for (int s = 0; s < 4; ++s) acc += v[k][s].a + v[k][s].d;
        // make the next iteration depend on this one (a group fetch is on the critical path in the wavefront)
        if (acc == -1.0) base += 1;
    }
    unsigned long long t1 = clock64();
    if (lane == 0) out[warp] = t1 - t0;
    if (acc == -1.0) *sink = acc;
}
// the same fetch with every lane's lines on a different 2 MB page, pages drawn from the whole buffer (5 GB): what the level-sorted
// tile layout does to a block whose upstream reaches sit in a dozen blocks of other levels
template <int KIND, int NUP>
__global__ void fetch_scattered(const double *buf, size_t n_pages, int span_pages, int iters, unsigned long long *out, double *sink) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    double acc = 0;
    unsigned long long t0 = clock64();
    size_t h = (size_t)warp * 7919 + 13;
    for (int it = 0; it < iters; ++it) {
        d4 v[NUP][4];
#pragma unroll
        for (int k = 0; k < NUP; ++k) {
            // page: a pseudo-random one out of span_pages pages following a per-iteration base; line inside the page varies too
            const size_t page = (h + (size_t)((lane * 2 + k) % span_pages) * 97) % n_pages;
            const double *g = buf + page * (size_t)(1 << 18) + (size_t)((it * 131 + lane * 17 + k * 5) & 4095) * 64;
#pragma unroll
            for (int s = 0; s < 4; ++s) v[k][s] = ld<KIND>(g + 4 * s);
        }
#pragma unroll
        for (int k = 0; k < NUP; ++k)
#pragma unroll
            for (int s = 0; s < 4; ++s) acc += v[k][s].a + v[k][s].d;
        h = h * 6364136223846793005ull + 1442695040888963407ull + (acc == -1.0 ? 1 : 0);
        h >>= 11;
    }
    unsigned long long t1 = clock64();
    if (lane == 0) out[warp] = t1 - t0;
    if (acc == -1.0) *sink = acc;
}
template <int KIND, int NUP>
static void run_scattered(const double *buf, size_t n_pages, int span, int blocks, int threads, unsigned long long *out, double *sink, double ghz) {
    const int iters = 2000;
    fetch_scattered<KIND, NUP><<<blocks, threads>>>(buf, n_pages, span, iters, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    double s = 0; int nw = blocks * threads / 32;
    for (int w = 0; w < nw; ++w) s += (double)out[w];
    printf("{\"load\": \"relaxed.gpu v4, scattered\", \"pages_per_fetch\": %d, \"warps\": %d, \"ns_per_group_fetch\": %.1f, \"err\": \"%s\"}\n", span, nw, s / nw / iters / ghz, cudaGetErrorString(e));
}
template <int KIND, int NUP>
static void run(const double *buf, const char *name, int blocks, int threads, unsigned long long *out, double *sink, double ghz) {
    const int iters = 2000;
    const size_t groups = 64;   // 64 groups x 32 x NUP x 512 B = 1-2 MB per warp: L2-resident, never the same line twice in L1's lifetime... (iters wrap: lines repeat after 64 groups)
    fetch<KIND, NUP><<<blocks, threads>>>(buf, groups, iters, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    double s = 0; int nw = blocks * threads / 32;
    for (int w = 0; w < nw; ++w) s += (double)out[w];
    printf("{\"load\": \"%s\", \"upstreams\": %d, \"warps\": %d, \"ns_per_group_fetch\": %.1f, \"err\": \"%s\"}\n", name, NUP, nw, s / nw / iters / ghz, cudaGetErrorString(e));
}
int main() {
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz / 1e6;
    const size_t n = (size_t)sms * 16 * 64 * 32 * 2 * 64;   // all warps x groups x lanes x NUP x pitch
    double *buf, *sink; unsigned long long *out;
    cudaMalloc(&buf, n * 8); cudaMalloc(&sink, 8);
    cudaMallocManaged(&out, sizeof(unsigned long long) * sms * 16);
    fill<<<sms * 8, 256>>>(buf, n);
    cudaDeviceSynchronize();
    printf("{\"sm_clock_ghz\": %.3f, \"buffer_mb\": %.0f}\n", ghz, n * 8 / 1e6);
    for (int cfg = 0; cfg < 2; ++cfg) {
        const int blocks = cfg == 0 ? 1 : sms * 2, threads = cfg == 0 ? 32 : 256;
        run<0, 2>(buf, "weak", blocks, threads, out, sink, ghz);
        run<1, 2>(buf, "cg", blocks, threads, out, sink, ghz);
        run<2, 2>(buf, "relaxed.gpu v4", blocks, threads, out, sink, ghz);
        run<3, 2>(buf, "relaxed.gpu 2 x v2", blocks, threads, out, sink, ghz);
        run<4, 2>(buf, "volatile v4", blocks, threads, out, sink, ghz);
        run<0, 1>(buf, "weak", blocks, threads, out, sink, ghz);
        run<2, 1>(buf, "relaxed.gpu v4", blocks, threads, out, sink, ghz);
        for (int span : {1, 2, 4, 8, 16, 64}) run_scattered<2, 2>(buf, n * 8 / (2u << 20), span, blocks, threads, out, sink, ghz);
    }
    // a few hundred warps, like the waiting chain of a small network
    for (int span : {1, 4, 16, 64}) run_scattered<2, 2>(buf, n * 8 / (2u << 20), span, sms, 64, out, sink, ghz);
    return 0;
}
