// Micro-benchmark: latency of a round of independent loads when the rounds cycle through P distinct 2 MB pages per SM
// (is the micro-TLB the reason small activation reads inside the persistent decoder kernel take microseconds?).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__global__ void __launch_bounds__(128, 1) k(const uint4* buf, size_t page_stride16, int pages, int rounds, unsigned* out, long long* cyc) {
    const int tid = threadIdx.x;
    unsigned acc = 0;
    long long total = 0;
    for (int r = 0; r < rounds; ++r) {
        // every round touches ONE page (like one LL buffer / bias vector), the page changes from round to round
        const uint4* p = buf + (size_t)((r * 7 + blockIdx.x) % pages) * page_stride16 + (size_t)(r % 16) * 4096;
        __syncthreads();
        const long long t0 = clock64();
        uint4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[i].x), "=r"(v[i].y), "=r"(v[i].z), "=r"(v[i].w) : "l"(p + tid + i * 128) : "memory");
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += v[i].x ^ v[i].w;
        total += clock64() - t0;
    }
    out[blockIdx.x * 128 + tid] = acc;
    if (tid == 0) cyc[blockIdx.x] = total;
}
int main() {
    int n_sms = 0; cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t page = 2u << 20;
    const int max_pages = 512;
    uint4* buf; unsigned* out; long long* cyc;
    cudaMalloc(&buf, page * max_pages); cudaMemset(buf, 1, page * max_pages); cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4096);
    for (int pages : {1, 4, 8, 16, 24, 32, 64, 128, 512}) {
        const int rounds = 2000;
        k<<<n_sms, 128>>>(buf, page / 16, pages, rounds, out, cyc);   // warm
        k<<<n_sms, 128>>>(buf, page / 16, pages, rounds, out, cyc);
        cudaDeviceSynchronize();
        long long h[256]; cudaMemcpy(h, cyc, n_sms * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < n_sms; ++i) avg += h[i]; avg /= n_sms * (double)rounds;
        printf("pages per SM working set = %3d: %.0f cycles per round of 8 x 16 B loads per thread\n", pages, avg);
    }
    return 0;
}
