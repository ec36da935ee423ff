// Micro-benchmark: every CTA (one per SM) reads the SAME small buffer from L2 - the activation broadcast between the
// stages of the persistent decoder step kernel.  Compares load flavours and sizes.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int MODE>
__device__ __forceinline__ uint4 ld16(const uint4* p) {
    uint4 r;
    if (MODE == 0) { u64 a, b; asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory"); r.x = (uint32_t)a; r.y = (uint32_t)(a >> 32); r.z = (uint32_t)b; r.w = (uint32_t)(b >> 32); }
    else if (MODE == 1) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    else if (MODE == 2) asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    else asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
// MODE 0..3: register loads, B per thread in flight; MODE 4: one cp.async.bulk of the whole buffer into shared memory
template <int MODE, int B, int ROT = 0>
__global__ void __launch_bounds__(128, 1) k(const uint4* buf, int n16, int iters, unsigned* out, long long* cyc) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    const int tid = threadIdx.x;
    if (MODE == 4 && tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    unsigned acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 4) {
            if (tid == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n16 * 16) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + 128)), "l"(buf), "r"(n16 * 16), "r"(smem_u32(bar)) : "memory");
            }
            uint32_t ok = 0;
            while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(it & 1) : "memory");
            acc += smem[128 + tid * 16];
            __syncthreads();
        } else {
            for (int base = 0; base < n16; base += 128 * B) {
                uint4 r[B];
#pragma unroll
                for (int i = 0; i < B; ++i) { const int e = base + i * 128 + tid; if (e < n16) r[i] = ld16<MODE>(buf + (ROT ? (e + blockIdx.x * 64) % n16 : e)); }
#pragma unroll
                for (int i = 0; i < B; ++i) { const int e = base + i * 128 + tid; if (e < n16) acc += r[i].x ^ r[i].y ^ r[i].z ^ r[i].w; }
            }
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * 128 + tid] = acc;
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE, int B, int ROT = 0> void run(const char* name, const uint4* buf, int bytes, unsigned* out, long long* cyc, int n_sms) {
    const int iters = 200;
    cudaFuncSetAttribute(k<MODE, B, ROT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    k<MODE, B, ROT><<<n_sms, 128, 220 * 1024>>>(buf, bytes / 16, iters, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    long long h[256]; cudaMemcpy(h, cyc, n_sms * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < n_sms; ++i) mx = h[i] > mx ? h[i] : mx;
    const double us = mx / 1.965e3 / iters;
    printf("%-34s %6d B x %3d CTAs: %7.2f us per sweep  (%.1f GB/s per SM, %.2f TB/s aggregate)\n", name, bytes, n_sms, us, bytes / us / 1e3, (double)bytes * n_sms / us / 1e6);
}
int main() {
    int n_sms = 0; cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, 0);
    uint4* buf; unsigned* out; long long* cyc;
    cudaMalloc(&buf, 1 << 20); cudaMemset(buf, 1, 1 << 20); cudaFuncSetAttribute(k<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4096);
    for (int bytes : {12800, 25600, 51200, 102400, 204800}) {
        run<0, 13>("ld.relaxed.gpu.v2.u64 x13", buf, bytes, out, cyc, n_sms);
        run<0, 13, 1>("ld.relaxed.gpu.v2.u64 x13 rotated", buf, bytes, out, cyc, n_sms);
        run<0, 4>("ld.relaxed.gpu.v2.u64 x4", buf, bytes, out, cyc, n_sms);
        run<1, 13>("ld.global.cg.v4 x13", buf, bytes, out, cyc, n_sms);
        run<2, 13>("ld.global.nc.v4 x13", buf, bytes, out, cyc, n_sms);
        run<3, 13>("ld.volatile.v4 x13", buf, bytes, out, cyc, n_sms);
        run<1, 26>("ld.global.cg.v4 x26", buf, bytes, out, cyc, n_sms);
        run<4, 1>("cp.async.bulk -> smem", buf, bytes, out, cyc, n_sms);
        run<1, 13>("ld.global.cg.v4 x13, 1 CTA", buf, bytes, out, cyc, 1);
        run<4, 1>("cp.async.bulk -> smem, 1 CTA", buf, bytes, out, cyc, 1);
    }
    return 0;
}
