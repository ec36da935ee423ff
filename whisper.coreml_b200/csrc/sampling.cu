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

// batched windows: blockIdx.z = window; the last CTA of a window to publish its partial updates that window's beams
__global__ void __launch_bounds__(256) sample_update_batch_kernel(const SampleBatchArgs b) {
    __shared__ int stage[DEC_MAX_BEAMS * DEC_TOK_LD];
    __shared__ int s_last;
    const int w = blockIdx.z;
    unsigned long long* dbg = b.dbg && !b.st[w].done ? b.dbg + ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 : nullptr;   // (replays past the end leave no marks)
    auto mark = [&](int i) { if (dbg && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); dbg[i] = t; } };
    mark(0);
    SampleArgs a;
    a.logits = b.logits + (long)w * b.row_stride_w * b.ld_logits; a.ld_logits = (long)b.row_stride_b * b.ld_logits;
    a.tokens = b.tokens + (long)w * b.slot_stride * DEC_TOK_LD; a.st = b.st + w; a.spec = b.spec; a.nb = b.nb; a.k = b.k;
    a.part = b.part + w; a.cand_lp = b.cand_lp + w * DEC_MAX_BEAMS * SAMPLE_MAX_K; a.cand_tok = b.cand_tok + w * DEC_MAX_BEAMS * SAMPLE_MAX_K;
    sample_partial_body<256>(a, blockIdx.x, blockIdx.y, threadIdx.x, BlockSync());
    mark(1);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned n = gridDim.x * gridDim.y;
        s_last = atomicAdd(&a.part->arrivals, 1u) == n - 1;
    }
    __syncthreads();
    mark(2);
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) a.part->arrivals = 0;
    BeamUpdateArgs u;
    u.part = a.part; u.timestamp_begin = b.spec.timestamp_begin; u.update = 1; u.cand_lp = a.cand_lp; u.cand_tok = a.cand_tok; u.nb = b.nb; u.k = b.k;
    u.tokens = b.tokens + (long)w * b.slot_stride * DEC_TOK_LD; u.table = b.table + (long)w * b.slot_stride * 448;
    u.fin_tokens = b.fin_tokens + (long)w * DEC_MAX_BEAMS * DEC_TOK_LD; u.st = b.st + w; u.eot = b.spec.eot; u.n_text_ctx = b.n_text_ctx;
    beam_update_body<256, false>(u, stage, nullptr, threadIdx.x, BlockSync());
    mark(3);
}

void sample_and_update_batch(const SampleBatchArgs& a, cudaStream_t s) {
    dim3 grid(SAMPLE_CHUNKS, a.nb, a.W);
    sample_update_batch_kernel<<<grid, 256, 0, s>>>(a);
    B200_LAUNCH_CHECK();
}

__global__ void __launch_bounds__(1024) no_speech_kernel(const float* __restrict__ x, int V, int tok, DecodeState* st, long ld) {
    __shared__ float rm[32], rs[32];
    x += (long)blockIdx.x * ld; st += blockIdx.x;                      // one window per CTA
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
void no_speech_prob_batch(const float* logits, long ld_logits, int n_vocab, int no_speech, DecodeState* st, int W, cudaStream_t s) {
    no_speech_kernel<<<W, 1024, 0, s>>>(logits, n_vocab, no_speech, st, ld_logits);
    B200_LAUNCH_CHECK();
}
void no_speech_prob(const float* logits, int n_vocab, int no_speech, DecodeState* st, cudaStream_t s) {
    no_speech_kernel<<<1, 1024, 0, s>>>(logits, n_vocab, no_speech, st, 0);
    B200_LAUNCH_CHECK();
}

}  // namespace b200
