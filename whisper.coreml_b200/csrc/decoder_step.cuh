// decoder1: one beam-batched token step (whisper/decoder.py:241-257 + :261-327, coreml.mm:404-444).
// Memory-bound: every weight byte is read once per step and shared by all beams.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int STEP_MAX_BEAMS = 8;     // beams are the N=8 dimension of mma.m16n8k16

struct StepGemv {
    const bf16* w_frag;        // fragment-major weights (export.py:to_frag), ceil(N/16) x K/32 blocks of 1 KB
    const float* bias;         // [N] or nullptr
    int N, K;
    // input vector per beam: either LN(x_f32) (ln_g != nullptr), plain x_f32, or x_bf16
    const float* x_f32; long ld_x;
    const float* ln_g; const float* ln_b; float eps;
    const bf16* x_bf16;
    int nb;
    int gelu;
    const float* residual; long ld_res;    // fp32 [nb][ld_res] added to the output (may alias out_f32)
    float* out_f32; bf16* out_bf16; long ld_out;
    const int* d_skip;         // device flag: non-zero -> the launch is a no-op (decode already complete)
};
void step_gemv(const StepGemv& g, cudaStream_t s);

struct StepSelfAttn {
    const float* qkv;          // [nb][3d] fp32 (q | k | v), biases applied, q pre-scaled
    bf16* cache_k; bf16* cache_v;          // layer slices of the KV cache: [slots][448][d]
    int* table;                // [slots][448] physical slot of logical (beam, position) (see api.cu: rearrange_mkv)
    const float* mask;         // (449) additive fp32 on the device or nullptr (all visible)
    int text_offset;           // number of cached positions; the new row is written at this index
    const int* d_text_offset;  // if set, read the offset from device memory instead (graph replay)
    const int* d_skip;
    int nb, n_head, d;
    bf16* out;                 // [nb][d]
};
void step_self_attn(const StepSelfAttn& a, cudaStream_t s);

struct StepCrossAttn {
    const float* q;            // [nb][d] fp32
    const bf16* ck; const bf16* cv;        // [H][1500][64] of this layer / window
    int nb, n_head, d, n_keys;
    float* part;               // scratch: [H][splits][8][66] (m, l, o[64])
    int* counters;             // [H], zero between launches
    bf16* out;                 // [nb][d]
    const int* d_skip;
};
void step_cross_attn(const StepCrossAttn& a, cudaStream_t s);

// x[b, :] = tok_emb[token[b], :] + pos_emb[pos, :]   (whisper/decoder.py:202), bf16 table, fp32 out
// tokens: [nb][token_stride]; the token at column `pos` (or *d_pos) of each row is embedded at position `pos`.
void step_embed(const bf16* tok_emb, const float* pos_emb, const int* tokens, long token_stride, int pos, const int* d_pos,
                const int* d_skip, int nb, int d, float* x, cudaStream_t s);

}  // namespace b200
