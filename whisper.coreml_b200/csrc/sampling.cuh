// Device-side logit filters, log-softmax, top-k and greedy / beam-search bookkeeping
// (whisper/decoding.py:299-431, 450-532), so that a whole DecodingTask._main_loop (:707-737) runs
// without a host round trip per token.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int DEC_MAX_BEAMS = 8;
constexpr int DEC_TOK_LD = 449;        // n_text_ctx + 1

struct DecodeSpec {                    // token ids (whisper/tokenizer.py), set by b200SetDecodeSpec
    int sot = -1, eot = -1, no_timestamps = -1, timestamp_begin = -1, no_speech = -1, n_vocab = 0;
    int blank[4] = {-1, -1, -1, -1};
    const uint8_t* d_suppress = nullptr;   // [n_vocab] 1 = SuppressTokens member
};

struct DecodeState {                   // one per decode, in device memory
    int L;                             // tokens per beam so far
    int pos;                           // text_offset of the next decoder1 step (= L - 1)
    int step;                          // loop index i of _main_loop
    int done;
    int n_finished;
    int sample_begin, sample_len, beam_mode, without_timestamps, max_initial_ts, suppress_blank;
    float no_speech_prob;
    float temperature;                 // 0: argmax / beam search; > 0: Categorical(logits / temperature) per row (decoding.py:307-310)
    unsigned seed_lo, seed_hi, stream;  // counter-based RNG key and the decode's stream id (window), see sampling_dev.cuh
    float sum_lp[DEC_MAX_BEAMS];
    float fin_score[DEC_MAX_BEAMS];
    int fin_len[DEC_MAX_BEAMS];
};

// Phase 1: the vocabulary is cut into SAMPLE_TEXT_CHUNKS text chunks + 1 timestamp chunk per beam (one CTA each, so
// a 5-beam step fills the GPU); each CTA applies the logit filters and leaves (max, sum exp) and its top-k.
constexpr int SAMPLE_TEXT_CHUNKS = 28, SAMPLE_CHUNKS = SAMPLE_TEXT_CHUNKS + 1, SAMPLE_MAX_K = DEC_MAX_BEAMS + 1;
constexpr int SAMPLE_CHUNK_TOKENS_MAX = 2048;   // tokens one (chunk, beam) CTA can hold (sampling_dev.cuh: SP_CHUNK_MAX)
struct SamplePartials {                // device scratch
    float m[DEC_MAX_BEAMS][SAMPLE_CHUNKS], s[DEC_MAX_BEAMS][SAMPLE_CHUNKS];
    float topv[DEC_MAX_BEAMS][SAMPLE_CHUNKS][SAMPLE_MAX_K];
    float topx[DEC_MAX_BEAMS][SAMPLE_CHUNKS];      // temperature > 0: the UNPERTURBED logit of the chunk's winner (topv holds logit / T + Gumbel noise)
    int topi[DEC_MAX_BEAMS][SAMPLE_CHUNKS][SAMPLE_MAX_K];
    unsigned arrivals;                 // CTAs of the fused kernel that have written their partial (re-armed by the last one)
};
struct SampleArgs {
    const float* logits; long ld_logits;   // [nb][V]
    const int* tokens;                 // [nb][DEC_TOK_LD]
    DecodeState* st;
    DecodeSpec spec;
    int nb, k;                         // k = nb + 1 candidates per beam (1 for greedy)
    SamplePartials* part;
    float* cand_lp; int* cand_tok;     // [nb][k] (written by the update kernel; kept for decoder1StepFused)
};
void sample_partial(const SampleArgs& a, cudaStream_t s);

// Phase 2 (one CTA): merge the partials into log-softmax + top-k per beam, then GreedyDecoder.update /
// BeamSearchDecoder.update.  `update` = 0 stops after the candidates (decoder1StepFused).
struct BeamUpdateArgs {
    const SamplePartials* part; int timestamp_begin; int update;
    float* cand_lp; int* cand_tok; int nb, k;
    int* tokens;                       // [nb][DEC_TOK_LD], permuted + extended in place
    int* table;                        // [nb][448] KV slot table, permuted in place
    int* fin_tokens;                   // [DEC_MAX_BEAMS][DEC_TOK_LD]
    DecodeState* st;
    int eot, n_text_ctx;
};
void beam_update(const BeamUpdateArgs& a, cudaStream_t s);

// ---- batched windows: one launch samples and updates every window of a batched decoder step --------------------------------
// Window w owns rows [w * slot_stride, w * slot_stride + nb) of tokens / table, st[w], part[w], fin_tokens[w], cand_*[w]; the
// logits of its beam b are row w * row_stride_w + b * row_stride_b (0 for b right after the prompt: all beams share the row).
struct SampleBatchArgs {
    const float* logits; long ld_logits; int row_stride_w, row_stride_b;
    int* tokens; int* table; int* fin_tokens;      // [W * slot_stride][DEC_TOK_LD], [W * slot_stride][448], [W][DEC_MAX_BEAMS][DEC_TOK_LD]
    DecodeState* st; SamplePartials* part;         // [W]
    float* cand_lp; int* cand_tok;                 // [W][DEC_MAX_BEAMS * SAMPLE_MAX_K]
    DecodeSpec spec;
    int W, nb, k, slot_stride, n_text_ctx;
    unsigned long long* dbg;                       // optional %globaltimer marks [CTA][8] (tools/step_timeline.py)
};
void sample_and_update_batch(const SampleBatchArgs& a, cudaStream_t s);
// st[w].no_speech_prob for W windows from logits rows w
void no_speech_prob_batch(const float* logits, long ld_logits, int n_vocab, int no_speech, DecodeState* st, int W, cudaStream_t s);

// st->no_speech_prob = softmax(logits)[no_speech]   (whisper/decoding.py:716-720)
void no_speech_prob(const float* logits, int n_vocab, int no_speech, DecodeState* st, cudaStream_t s);

}  // namespace b200
