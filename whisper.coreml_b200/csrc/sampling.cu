#include "sampling_dev.cuh"

namespace b200 {

struct BlockSync { __device__ __forceinline__ void operator()() const { __syncthreads(); } };

__global__ void __launch_bounds__(256) sample_partial_kernel(const SampleArgs a) {
    sample_partial_body<256>(a, blockIdx.x, blockIdx.y, threadIdx.x, BlockSync());
}

void sample_partial(const SampleArgs& a, cudaStream_t s) {
    dim3 grid(SAMPLE_CHUNKS, a.nb);
    sample_partial_kernel<<<grid, 256, 0, s>>>(a);
    B200_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------
// GreedyDecoder.update (:303-318) / BeamSearchDecoder.update (:350-409), one CTA.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) beam_update_kernel(const BeamUpdateArgs a) {
    __shared__ int stage[DEC_MAX_BEAMS * DEC_TOK_LD];
    beam_update_body<256, false>(a, stage, nullptr, threadIdx.x, BlockSync());
}

void beam_update(const BeamUpdateArgs& a, cudaStream_t s) {
    beam_update_kernel<<<1, 256, 0, s>>>(a);
    B200_LAUNCH_CHECK();
}

// phase 1 in every CTA; the last CTA to publish its partial (release fence + counter) merges them and updates the beams
__global__ void __launch_bounds__(256) sample_update_kernel(const SampleArgs a, const BeamUpdateArgs u) {
    __shared__ int stage[DEC_MAX_BEAMS * DEC_TOK_LD];
    __shared__ int s_last;
    sample_partial_body<256>(a, blockIdx.x, blockIdx.y, threadIdx.x, BlockSync());
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned n = gridDim.x * gridDim.y;
        s_last = atomicAdd(&a.part->arrivals, 1u) == n - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) a.part->arrivals = 0;
    beam_update_body<256, false>(u, stage, nullptr, threadIdx.x, BlockSync());
}

void sample_and_update(const SampleArgs& a, const BeamUpdateArgs& u, cudaStream_t s) {
    dim3 grid(SAMPLE_CHUNKS, a.nb);
    sample_update_kernel<<<grid, 256, 0, s>>>(a, u);
    B200_LAUNCH_CHECK();
}

__global__ void __launch_bounds__(1024) no_speech_kernel(const float* __restrict__ x, int V, int tok, DecodeState* st) {
    __shared__ float rm[32], rs[32];
    float m = -INFINITY, s = 0.f;
    for (int v = threadIdx.x; v < V; v += 1024) online_add(m, s, x[v]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) online_merge(m, s, __shfl_xor_sync(0xffffffffu, m, o), __shfl_xor_sync(0xffffffffu, s, o));
    if ((threadIdx.x & 31) == 0) { rm[threadIdx.x >> 5] = m; rs[threadIdx.x >> 5] = s; }
    __syncthreads();
    if (threadIdx.x < 32) {
        m = rm[threadIdx.x]; s = rs[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) online_merge(m, s, __shfl_xor_sync(0xffffffffu, m, o), __shfl_xor_sync(0xffffffffu, s, o));
        if (threadIdx.x == 0) st->no_speech_prob = expf(x[tok] - m) / s;
    }
}
void no_speech_prob(const float* logits, int n_vocab, int no_speech, DecodeState* st, cudaStream_t s) {
    no_speech_kernel<<<1, 1024, 0, s>>>(logits, n_vocab, no_speech, st);
    B200_LAUNCH_CHECK();
}

}  // namespace b200
