// Latency floor of the wavefront's hand-over: two warps on different SMs bounce a counter through global memory.
//   mode 0: st.relaxed.gpu / ld.relaxed.gpu spin on one 8-byte word                 (the sentinel protocol's hand-over)
//   mode 1: st.release.gpu / ld.acquire.gpu on a flag                               (the flag protocol's hand-over)
//   mode 2: mode 0, plus a 32-byte data sector written before the word and read after it was seen (data + flag, no fence)
//   mode 3: writer stores 4 x 32 B (one line) with st.relaxed.gpu.v2, reader polls the last 8 bytes, then loads the line
// Prints ns per one-way hop.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pingpong pingpong.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void pingpong(unsigned long long *buf, int iters, int mode, int peer, unsigned long long *out_ns) {
    // block 0 and block `peer` play; everyone else exits.  buf: [0..15] line A (0 -> peer), [16..31] line B (peer -> 0)
    const int me = blockIdx.x == 0 ? 0 : (blockIdx.x == peer ? 1 : -1);
    if (me < 0 || threadIdx.x != 0) return;
    unsigned long long *mine = buf + (me == 0 ? 0 : 16), *theirs = buf + (me == 0 ? 16 : 0);
    unsigned long long t0 = 0, t1 = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int it = 1; it <= iters; ++it) {
        if (me == 0) {
            if (mode == 1) st_release(mine + 15, it);
            else {
                if (mode >= 2) for (int k = 0; k < (mode == 3 ? 15 : 3); ++k) st_relaxed(mine + k, it);
                st_relaxed(mine + 15, it);
            }
            while ((mode == 1 ? ld_acquire(theirs + 15) : ld_relaxed(theirs + 15)) < (unsigned long long)it) { }
            if (mode >= 2) { unsigned long long s = 0; for (int k = 0; k < (mode == 3 ? 15 : 3); ++k) s += ld_relaxed(theirs + k); if (s == 1) out_ns[2] = s; }
        } else {
            while ((mode == 1 ? ld_acquire(theirs + 15) : ld_relaxed(theirs + 15)) < (unsigned long long)it) { }
            if (mode >= 2) { unsigned long long s = 0; for (int k = 0; k < (mode == 3 ? 15 : 3); ++k) s += ld_relaxed(theirs + k); if (s == 1) out_ns[2] = s; }
            if (mode == 1) st_release(mine + 15, it);
            else {
                if (mode >= 2) for (int k = 0; k < (mode == 3 ? 15 : 3); ++k) st_relaxed(mine + k, it);
                st_relaxed(mine + 15, it);
            }
        }
    }
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (me == 0) out_ns[0] = t1 - t0;
}

// background load: every other block streams a buffer (reads) to load the memory system like the wavefront does
__global__ void pingpong_loaded(unsigned long long *buf, int iters, int mode, int peer, unsigned long long *out_ns, const double *bg, size_t bg_n,
                                volatile int *stop) {
    const int me = blockIdx.x == 0 ? 0 : (blockIdx.x == peer ? 1 : -1);
    if (me < 0) {
        double acc = 0;
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        while (!*stop) { acc += bg[i % bg_n]; i += (size_t)gridDim.x * blockDim.x; }
        if (acc == 1.2345) out_ns[3] = 1;
        return;
    }
    if (threadIdx.x != 0) return;
    unsigned long long *mine = buf + (me == 0 ? 0 : 16), *theirs = buf + (me == 0 ? 16 : 0);
    unsigned long long t0 = 0, t1 = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int it = 1; it <= iters; ++it) {
        if (me == 0) {
            if (mode == 1) st_release(mine + 15, it); else st_relaxed(mine + 15, it);
            while ((mode == 1 ? ld_acquire(theirs + 15) : ld_relaxed(theirs + 15)) < (unsigned long long)it) { }
        } else {
            while ((mode == 1 ? ld_acquire(theirs + 15) : ld_relaxed(theirs + 15)) < (unsigned long long)it) { }
            if (mode == 1) st_release(mine + 15, it); else st_relaxed(mine + 15, it);
        }
    }
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (me == 0) { out_ns[0] = t1 - t0; *stop = 1; __threadfence(); }
}

int main() {
    unsigned long long *buf, *out;
    cudaMalloc(&buf, 4096);
    cudaMallocManaged(&out, 64);
    const int iters = 20000;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int mode = 0; mode < 4; ++mode)
        for (int peer : {1, 2, sms / 2, sms - 1}) {
            cudaMemset(buf, 0, 4096);
            out[0] = 0;
            pingpong<<<sms, 32>>>(buf, iters, mode, peer, out);
            cudaError_t e = cudaDeviceSynchronize();
            printf("{\"mode\": %d, \"peer_block\": %d, \"ns_per_hop\": %.1f, \"err\": \"%s\"}\n", mode, peer, (double)out[0] / (2.0 * iters), cudaGetErrorString(e));
        }
    // with the rest of the machine streaming from DRAM
    double *bg;
    size_t bg_n = (size_t)1 << 28;   // 2 GB
    cudaMalloc(&bg, bg_n * 8);
    cudaMemset(bg, 0, bg_n * 8);
    int *stop;
    cudaMalloc(&stop, 4);
    for (int mode = 0; mode < 2; ++mode) {
        cudaMemset(buf, 0, 4096);
        cudaMemset(stop, 0, 4);
        out[0] = 0;
        pingpong_loaded<<<sms * 2, 256>>>(buf, iters, mode, sms / 2, out, bg, bg_n, stop);
        cudaError_t e = cudaDeviceSynchronize();
        printf("{\"mode\": %d, \"loaded\": true, \"ns_per_hop\": %.1f, \"err\": \"%s\"}\n", mode, (double)out[0] / (2.0 * iters), cudaGetErrorString(e));
    }
    return 0;
}
