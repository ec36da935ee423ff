// C ABI Part 2: batched windows and the device-resident decode loop.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <tuple>
#include <vector>

#include "api_batch.cuh"
#include "ops.cuh"
#include "sampling.cuh"
#include "gemm.cuh"
#include "state.cuh"
#include "timing.cuh"
#include "whisper_b200.h"

namespace b200 {

struct DecodeCtx {
    DecodeSpec spec;
    std::vector<uint8_t> h_suppress;
    uint8_t* d_suppress = nullptr; size_t suppress_cap = 0;
};
static DecodeCtx g_dc;

DecodeSpec decode_spec() { return g_dc.spec; }

// x256[r] = tok_emb[tok[r]] + pos_emb[r] for r < n, zero rows after (decoder.py:202,214); mask = causal with
// columns >= n masked (decoder.py:212-213)
__global__ void prefill_inputs_kernel(const bf16* __restrict__ tok_emb, const float* __restrict__ pos_emb,
                                      const int* __restrict__ tokens, int n, int d, float* x, float* mask) {
    const int r = blockIdx.x;
    const int tok = r < n ? tokens[r] : 0;
    for (int c = threadIdx.x; c < d; c += blockDim.x)
        x[(long)r * d + c] = r < n ? __bfloat162float(tok_emb[(long)tok * d + c]) + pos_emb[(long)r * d + c] : 0.f;
    for (int c = threadIdx.x; c < PREFILL_CTX; c += blockDim.x)
        mask[r * PREFILL_CTX + c] = (c <= r && c < n) ? 0.f : -INFINITY;
}

struct AlignCtx {
    float *tmp = nullptr, *mat = nullptr, *neg = nullptr, *logits = nullptr, *probs = nullptr;
    bf16* rows = nullptr;
    int *path = nullptr, *targets = nullptr, *d_tok = nullptr;
    uint8_t* scratch = nullptr;
    size_t tmp_cap = 0;
    bool ready = false;
};
static AlignCtx g_al;

__global__ void negate_kernel(const float* __restrict__ in, float* __restrict__ out, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = -in[i];
}
__global__ void token_prob_kernel(const float* __restrict__ logits, long ld, int eot, const int* __restrict__ targets,
                                  float* __restrict__ out);

void decode_clear_graphs() { batch_clear_graphs(); }

}  // namespace b200

using namespace b200;

extern "C" {

void b200SetDecodeSpec(int sot, int eot, int no_timestamps, int timestamp_begin, int no_speech, const int* suppress,
                       int n_suppress, const int* blank, int n_blank) {
    State& s = S();
    DecodeCtx& c = g_dc;
    if (!s.V) { record_error("b200SetDecodeSpec: load the decoder first (n_vocab unknown)"); return; }
    if (eot < 0 || eot >= s.V || timestamp_begin < 0 || timestamp_begin >= s.V) { record_error("b200SetDecodeSpec: eot %d / timestamp_begin %d outside the vocabulary (%d)", eot, timestamp_begin, s.V); return; }
    if (n_blank > 4) { record_error("b200SetDecodeSpec: %d blank tokens (at most 4 are supported)", n_blank); return; }
    if ((timestamp_begin + SAMPLE_TEXT_CHUNKS - 1) / SAMPLE_TEXT_CHUNKS > SAMPLE_CHUNK_TOKENS_MAX || s.V - timestamp_begin > SAMPLE_CHUNK_TOKENS_MAX) {
        record_error("b200SetDecodeSpec: vocabulary chunks of %d text / %d timestamp tokens exceed %d", (timestamp_begin + SAMPLE_TEXT_CHUNKS - 1) / SAMPLE_TEXT_CHUNKS,
                     s.V - timestamp_begin, SAMPLE_CHUNK_TOKENS_MAX);
        return;
    }
    use_device();
    B200_CHECK(cudaDeviceSynchronize());
    batch_clear_graphs();                                               // captured step graphs hold the old spec (and d_suppress) by value
    c.spec.sot = sot; c.spec.eot = eot; c.spec.no_timestamps = no_timestamps; c.spec.timestamp_begin = timestamp_begin;
    c.spec.no_speech = no_speech; c.spec.n_vocab = s.V;
    for (int i = 0; i < 4; ++i) c.spec.blank[i] = i < n_blank ? blank[i] : -1;
    c.h_suppress.assign((size_t)s.V, 0);
    for (int i = 0; i < n_suppress; ++i)
        if (suppress[i] >= 0 && suppress[i] < s.V) c.h_suppress[suppress[i]] = 1;
    // the flag buffer is kept and overwritten in place (captured graphs hold its address); it only moves when the vocabulary grows
    if ((!c.d_suppress || c.suppress_cap < (size_t)s.V) && !dev_alloc(&c.d_suppress, (size_t)s.V)) return;
    c.suppress_cap = std::max(c.suppress_cap, (size_t)s.V);
    B200_CHECK(cudaMemcpy(c.d_suppress, c.h_suppress.data(), (size_t)s.V, cudaMemcpyHostToDevice));
    c.spec.d_suppress = c.d_suppress;
}

void encoderPredictWindows(const float* d_mel, long total_frames, const int* seeks, int n_windows) {
    encoderPredictWindowsContent(d_mel, total_frames, total_frames, seeks, n_windows);
}

void encoderPredictWindowsContent(const float* d_mel, long total_frames, long content_frames, const int* seeks, int n_windows) {
    State& s = S();
    if (!s.enc_loaded) { record_error("encoderPredictWindows: encoder not loaded"); return; }
    if (n_windows < 1) return;
    use_device();
    B200_CHECK(cudaStreamSynchronize(cudaStreamLegacy));                // d_mel may have been produced on the caller's stream
    if (!ensure_encoder_capacity(n_windows)) return;
    B200_CHECK(cudaMemcpyAsync(s.d_seeks, seeks, (size_t)n_windows * sizeof(int), cudaMemcpyHostToDevice, s.stream));
    {
        StageTimer t(ST_ENCODER);
        run_encoder(d_mel, total_frames, content_frames, n_windows);
    }
    s.cur_window = 0;
    B200_CHECK(cudaStreamSynchronize(s.stream));
}

void crossKVPredictWindows(int n_windows) {
    State& s = S();
    if (!s.ckv_loaded || !s.enc_loaded) { record_error("crossKVPredictWindows: encoder / crossKV not loaded"); return; }
    if (n_windows < 1 || n_windows > s.n_windows) { record_error("crossKVPredictWindows: %d windows requested, %d encoded", n_windows, s.n_windows); return; }
    use_device();
    {
        StageTimer t(ST_CROSSKV);
        run_cross_kv(n_windows);
    }
    B200_CHECK(cudaStreamSynchronize(s.stream));
}

}  // extern "C"

extern "C" {

int b200DecodeWindowsEx(const int* windows, int n_windows, const int* initial_tokens, int n_initial, int beam_size, int n_group,
                        float temperature, unsigned long long seed, int sample_len, int without_timestamps,
                        int max_initial_timestamp_index, int* out_tokens, int* out_lengths, float* out_sum_logprobs,
                        float* out_no_speech, int* out_steps) {
    State& s = S();
    DecodeCtx& c = g_dc;
    if (!s.dec1_loaded || !s.dec256_loaded || !s.ckv_loaded) { record_error("b200DecodeWindows: decoder256 / decoder1 / crossKV not loaded"); return 0; }
    const int nb = beam_size > 0 ? beam_size : (n_group > 0 ? n_group : 1);
    if (nb > STEP_MAX_BEAMS) { record_error("b200DecodeWindows: %d beams / samples (at most %d)", nb, STEP_MAX_BEAMS); return 0; }
    if (beam_size > 0 && temperature != 0.f) { record_error("b200DecodeWindows: beam search runs at temperature 0 (whisper/transcribe.py:196-202)"); return 0; }
    if (temperature < 0.f) { record_error("b200DecodeWindows: temperature %g", temperature); return 0; }
    if (n_initial < 1 || n_initial > PREFILL_CTX) { record_error("b200DecodeWindows: n_initial %d outside [1, 256]", n_initial); return 0; }
    if (n_windows < 1 || sample_len < 1) return 0;
    for (int w = 0; w < n_windows; ++w)
        if (windows[w] < 0 || windows[w] >= (s.ckv_cap > 0 ? s.ckv_cap : 1)) { record_error("b200DecodeWindows: window %d outside [0, %d)", windows[w], s.ckv_cap); return 0; }
    use_device();
    if (c.spec.n_vocab != s.V || c.spec.eot < 0) { record_error("decode: call b200SetDecodeSpec (n_vocab %d) first", s.V); return 0; }
    if (!batch_available()) { record_error("b200DecodeWindows: the step kernel does not support these dimensions (n_state %d, %d layers)", s.d, s.Ld); return 0; }
    // windows are spread over concurrent decode lanes, every lane advancing its windows in ONE batched step kernel (api_batch.cu)
    return decode_windows_batch(windows, n_windows, initial_tokens, n_initial, beam_size, n_group, temperature, seed, sample_len,
                                without_timestamps, max_initial_timestamp_index, out_tokens, out_lengths, out_sum_logprobs, out_no_speech, out_steps);
}

int b200DecodeWindows(const int* windows, int n_windows, const int* initial_tokens, int n_initial, int beam_size, int sample_len,
                      int without_timestamps, int max_initial_timestamp_index, int* out_tokens, int* out_lengths,
                      float* out_sum_logprobs, float* out_no_speech, int* out_steps) {
    return b200DecodeWindowsEx(windows, n_windows, initial_tokens, n_initial, beam_size, 1, 0.f, 0ull, sample_len, without_timestamps,
                               max_initial_timestamp_index, out_tokens, out_lengths, out_sum_logprobs, out_no_speech, out_steps);
}

int b200DecodeWindow(const int* initial_tokens, int n_initial, int beam_size, int sample_len, int without_timestamps,
                     int max_initial_timestamp_index, int* out_tokens, int* out_lengths, float* out_sum_logprobs,
                     float* out_no_speech) {
    const int w = S().cur_window;
    return b200DecodeWindows(&w, 1, initial_tokens, n_initial, beam_size, sample_len, without_timestamps, max_initial_timestamp_index,
                             out_tokens, out_lengths, out_sum_logprobs, out_no_speech, nullptr);
}

void decoder1StepFused(const int* tokens_hist, int n_hist, int sample_begin, int text_offset, int without_timestamps,
                       int max_initial_timestamp_index, float* out_logprob, int* out_token) {
    State& s = S();
    if (!s.dec1_loaded || !s.dec256_loaded || !s.ckv_loaded) { record_error("decoder1StepFused: decoders not loaded"); return; }
    if (n_hist < 1 || n_hist > N_TEXT_CTX || text_offset != n_hist - 1) { record_error("decoder1StepFused: n_hist %d / text_offset %d", n_hist, text_offset); return; }
    use_device();
    if (g_dc.spec.n_vocab != s.V || g_dc.spec.eot < 0) { record_error("decode: call b200SetDecodeSpec (n_vocab %d) first", s.V); return; }
    if (!batch_available()) { record_error("decoder1StepFused: the step kernel does not support these dimensions"); return; }
    step_fused_abi(tokens_hist, n_hist, sample_begin, text_offset, without_timestamps, max_initial_timestamp_index, out_logprob, out_token);
}

int b200AlignTokens(const int* tokens, int n_tokens, int n_skip, int num_frames, int medfilt_width, int* out_i, int* out_j,
                    float* out_matrix, float* out_text_token_probs) {
    State& s = S();
    AlignCtx& a = g_al;
    if (!s.dec256_loaded || !s.ckv_loaded) { record_error("b200AlignTokens: decoder256 / crossKV not loaded"); return 0; }
    if (s.n_align < 1) { record_error("b200AlignTokens: loadDecoder256 was called with n_alignment_head = 0"); return 0; }
    const int F = num_frames / 2, n_rows = n_tokens - 1 - n_skip, n_text = n_tokens - n_skip - 2;
    if (n_tokens > PREFILL_CTX || n_rows < 1 || F < 1 || F > N_AUDIO_CTX) { record_error("b200AlignTokens: n_tokens %d n_skip %d num_frames %d", n_tokens, n_skip, num_frames); return 0; }
    use_device();
    cudaStream_t st = s.stream;
    const size_t need = (size_t)2 * s.n_align * PREFILL_CTX * N_AUDIO_CTX;
    if (!a.ready || a.tmp_cap < need) {
        bool ok = true;
        ok &= dev_alloc(&a.tmp, need); a.tmp_cap = need;
        ok &= dev_alloc(&a.mat, (size_t)PREFILL_CTX * N_AUDIO_CTX); ok &= dev_alloc(&a.neg, (size_t)PREFILL_CTX * N_AUDIO_CTX);
        ok &= dev_alloc(&a.logits, (size_t)PREFILL_CTX * ((s.V + 7) / 8 * 8)); ok &= dev_alloc(&a.probs, (size_t)PREFILL_CTX);
        ok &= dev_alloc(&a.rows, (size_t)PREFILL_CTX * s.d); ok &= dev_alloc(&a.path, (size_t)2 * (PREFILL_CTX + N_AUDIO_CTX) + 1);
        ok &= dev_alloc(&a.targets, (size_t)PREFILL_CTX); ok &= dev_alloc(&a.d_tok, (size_t)PREFILL_CTX);
        ok &= dev_alloc(&a.scratch, dtw_scratch_bytes(PREFILL_CTX, N_AUDIO_CTX));
        a.ready = ok;
        if (!ok) return 0;
    }
    StageTimer timer(ST_ALIGN);
    B200_CHECK(cudaMemcpyAsync(a.d_tok, tokens, (size_t)n_tokens * sizeof(int), cudaMemcpyHostToDevice, st));
    prefill_inputs_kernel<<<PREFILL_CTX, 256, 0, st>>>(s.tok_emb, s.pos_emb, a.d_tok, n_tokens, s.d, s.px, s.pmask);
    B200_LAUNCH_CHECK();
    run_prefill(0, true, n_tokens);                                     // model(tokens[None]) (timing.py:185, model.py:110-119)
    alignment_matrix_dev(s.pchw, s.n_align, PREFILL_CTX, n_tokens, F, n_skip, medfilt_width, a.tmp, a.mat, false, st);
    negate_kernel<<<cdiv(n_rows * F, 256), 256, 0, st>>>(a.mat, a.neg, (long)n_rows * F);   // dtw(-matrix) (timing.py:205)
    B200_LAUNCH_CHECK();
    const int cap = n_rows + F;
    B200_CHECK(cudaMemsetAsync(a.path + 2 * cap, 0, sizeof(int), st));
    dtw_dev(a.neg, n_rows, F, a.path, a.path + cap, a.path + 2 * cap, a.scratch, st);
    if (out_text_token_probs && n_text > 0) {                           // token_probs over [:eot] (timing.py:187-190)
        f32_to_bf16(s.pout + (size_t)n_skip * s.d, a.rows, (long)n_text * s.d, st);
        GemmParams g = gemm_plain(a.rows, s.tok_emb, a.logits, n_text, s.V, s.d);
        g.c_fp32 = 1; g.ldc = (s.V + 7) / 8 * 8;                        // padded rows keep the epilogue stores vectorised
        gemm_tcgen05(g, st);
        B200_CHECK(cudaMemcpyAsync(a.targets, tokens + n_skip + 1, (size_t)n_text * sizeof(int), cudaMemcpyHostToDevice, st));
        token_prob_kernel<<<n_text, 1024, 0, st>>>(a.logits, (s.V + 7) / 8 * 8, g_dc.spec.eot >= 0 ? g_dc.spec.eot : s.V, a.targets, a.probs);
        B200_LAUNCH_CHECK();
        B200_CHECK(cudaMemcpyAsync(out_text_token_probs, a.probs, (size_t)n_text * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    int len = 0;
    B200_CHECK(cudaMemcpyAsync(&len, a.path + 2 * cap, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (out_matrix) B200_CHECK(cudaMemcpyAsync(out_matrix, a.mat, (size_t)n_rows * F * sizeof(float), cudaMemcpyDeviceToHost, st));
    B200_CHECK(cudaStreamSynchronize(st));
    if (len > 0 && len <= cap) {
        B200_CHECK(cudaMemcpy(out_i, a.path, (size_t)len * sizeof(int), cudaMemcpyDeviceToHost));
        B200_CHECK(cudaMemcpy(out_j, a.path + cap, (size_t)len * sizeof(int), cudaMemcpyDeviceToHost));
    }
    return len;
}

void b200GetStageTimes(float* out_ms7, int reset) { stage_times(out_ms7, reset != 0); }

}  // extern "C"
