// Micro-benchmark: latency of a software grid barrier on B200 (one CTA per SM, cooperative launch), alone and with a
// background stream of cp.async.bulk loads per CTA (the situation inside the persistent decoder step kernel).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bench_barrier tools/bench_barrier.cu && /tmp/bench_barrier
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_volatile(const unsigned* p) { unsigned v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// variant 0: red.release + ld.relaxed spin + fence.acq_rel      (decoder_mega.cu today)
// variant 1: red.release + ld.acquire spin
// variant 2: __threadfence + atomicAdd (relaxed) + volatile spin + __threadfence
// variant 3: red.relaxed (no release) + ld.relaxed spin, no fences  (lower bound: pure signalling latency)
// variant 4: per-CTA flag lines: st.release own flag; CTA 0's 148 threads poll all flags then st.release a go word; others poll go
template <int VARIANT>
__device__ __forceinline__ void grid_sync(unsigned* bar, unsigned& k, int nctas, int tid) {
    asm volatile("bar.sync 1, 256;" ::: "memory");
    ++k;
    if (VARIANT == 4) {
        unsigned* flags = bar + 64;                      // flags[cta * 32]
        unsigned* go = bar + 32;
        if (blockIdx.x == 0) {
            if (tid < nctas) {
                if (tid == 0) { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
                else while (ld_relaxed(flags + tid * 32) < k) {}
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (tid == 0) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(go), "r"(k) : "memory"); }
        } else if (tid == 0) {
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x * 32), "r"(k) : "memory");
            while (ld_relaxed(go) < k) {}
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
        }
    } else if (tid == 0) {
        const unsigned target = k * (unsigned)nctas;
        if (VARIANT == 0) {
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
            while (ld_relaxed(bar) < target) {}
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
        } else if (VARIANT == 1) {
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
            while (ld_acquire(bar) < target) {}
        } else if (VARIANT == 2) {
            __threadfence();
            atomicAdd(bar, 1u);
            while (ld_volatile(bar) < target) {}
            __threadfence();
        } else {
            asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
            while (ld_relaxed(bar) < target) {}
        }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
}

template <int VARIANT>
__global__ void __launch_bounds__(288, 1) barrier_kernel(unsigned* bar, int iters, const uint8_t* stream, size_t stream_bytes_per_cta,
                                                         int background, float* payload, unsigned long long* out_ns) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint8_t* ring = smem + 1024;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr int NS = 16, SLOT = 8192;
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 8) {
        if (tid != 256 || !background) return;
        // keep NS bulk copies in flight; nobody consumes them, the producer just recycles the slots
        const uint8_t* src = stream + (size_t)blockIdx.x * stream_bytes_per_cta;
        const size_t n = stream_bytes_per_cta / SLOT;
        uint32_t phase = 0;
        volatile unsigned* stop = bar + 16;
        for (size_t i = 0;; ++i) {
            const int s = (int)(i % NS);
            if (i >= NS) {
                uint32_t ok = 0;
                while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&full[s])), "r"(phase) : "memory");
                if (s == NS - 1) phase ^= 1;
            }
            if (*stop) break;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(SLOT) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(ring + s * SLOT)), "l"(src + (i % n) * SLOT), "r"(SLOT), "r"(smem_u32(&full[s])) : "memory");
        }
        // drain
        return;
    }
    unsigned k = 0;
    grid_sync<VARIANT>(bar, k, gridDim.x, tid);
    const unsigned long long t0 = gtimer();
    float acc = 0.f;
    for (int i = 0; i < iters; ++i) {
        // a token amount of cross-CTA traffic per barrier: every CTA writes one float, reads its neighbour's
        if (tid == 0) __stcg(payload + blockIdx.x * 32, (float)i);
        grid_sync<VARIANT>(bar, k, gridDim.x, tid);
        if (tid == 0) acc += __ldcg(payload + ((blockIdx.x + 1) % gridDim.x) * 32);
    }
    const unsigned long long t1 = gtimer();
    if (tid == 0) {
        out_ns[blockIdx.x] = t1 - t0;
        if (acc < 0) out_ns[blockIdx.x] = 0;
        if (blockIdx.x == 0) bar[16] = 1;               // tell the background producers to stop
    }
}

template <int VARIANT>
static void run(const char* name, int background, unsigned* bar, const uint8_t* stream, size_t per_cta, float* payload,
                unsigned long long* out, int n_sms) {
    const int iters = 2000;
    const size_t smem = 1024 + 16 * 8192;
    cudaFuncSetAttribute(barrier_kernel<VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemset(bar, 0, 64 * 1024);
    int it = iters;
    void* params[] = {&bar, &it, &stream, &per_cta, &background, &payload, &out};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)barrier_kernel<VARIANT>, dim3(n_sms), dim3(288), params, smem, 0);
    if (e != cudaSuccess) { printf("%s: launch %s\n", name, cudaGetErrorString(e)); return; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    unsigned long long h[256];
    cudaMemcpy(h, out, n_sms * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    unsigned long long mx = 0;
    for (int i = 0; i < n_sms; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%-58s background=%d  %.3f us per barrier\n", name, background, mx / 1000.0 / iters);
}

int main() {
    int n_sms = 0;
    cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned* bar; float* payload; unsigned long long* out; uint8_t* stream;
    const size_t per_cta = (size_t)24 << 20;            // 24 MB per CTA = 3.5 GB total, ~0.5 ms of HBM streaming
    cudaMalloc(&bar, 64 * 1024); cudaMalloc(&payload, 64 * 1024); cudaMalloc(&out, 4096);
    cudaMalloc(&stream, per_cta * n_sms);
    cudaMemset(stream, 1, per_cta * n_sms);
    for (int bg = 0; bg < 2; ++bg) {
        run<0>("0: red.release + ld.relaxed spin + fence.acq_rel", bg, bar, stream, per_cta, payload, out, n_sms);
        run<1>("1: red.release + ld.acquire spin", bg, bar, stream, per_cta, payload, out, n_sms);
        run<2>("2: threadfence + atomicAdd + volatile spin + threadfence", bg, bar, stream, per_cta, payload, out, n_sms);
        run<3>("3: relaxed red + relaxed spin, no fences (lower bound)", bg, bar, stream, per_cta, payload, out, n_sms);
        run<4>("4: per-CTA flags gathered by CTA 0 + go word", bg, bar, stream, per_cta, payload, out, n_sms);
    }
    return 0;
}
