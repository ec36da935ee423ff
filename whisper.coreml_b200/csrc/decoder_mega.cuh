// Persistent decoder1 step kernel (decoder_mega.cu): interface.
#pragma once
#include "common.cuh"
#include "sampling.cuh"
#include "state.cuh"

namespace b200 {

constexpr int MEGA_MAX_LAYERS = 32;
constexpr int MEGA_DBG_LD = 640;      // timeline marks per CTA (2 per stage + 1)

struct MegaLayer {
    const bf16 *qkv, *attn_out, *cross_q, *cross_out, *mlp1, *mlp2;                  // fragment-major weights
    const float *qkv_b, *attn_out_b, *cross_q_b, *cross_out_b, *mlp1_b, *mlp2_b;
    const float *ln1_w, *ln1_b, *ln2_w, *ln2_b, *ln3_w, *ln3_b;
};
struct MegaModel {                      // lives in device memory, built when both decoders are loaded
    int d, H, Ld, V, n_tiles_vocab;
    const bf16* tok_emb; const bf16* tok_emb_frag; const float* pos_emb; const float* ln_w; const float* ln_b;
    MegaLayer layers[MEGA_MAX_LAYERS];
};
struct MegaArgs {
    const bf16* ckv_frag;               // fragment-major cross K / V^T of the current window
    int nb;
    // shared-memory geometry (mega_* helpers below): bf16 activation rows [xs_rows][xs_cols + 32] first, then scratch,
    // then the ring of n_slots 40 KB slots at ring_offset; sa_cap = cached positions a self-attention unit stages
    int xs_cols, xs_rows, ring_offset, n_slots, sa_cap;
    // activations handed from stage to stage as LL words {payload, epoch} (see decoder_mega.cu)
    uint2 *ll_qkv;                      // [8][3d]   fp32   q | k | v of the new token
    uint2 *ll_att, *ll_catt;            // [8][d/2]  bf16x2 self- / cross-attention output
    uint2 *ll_x1, *ll_x2, *ll_x3;       // [8][d]    fp32   residual stream after self-attention / cross-attention / MLP
    uint2 *ll_x1b, *ll_x2b, *ll_x3b;    // [8][d/2]  bf16x2 the same values for the LayerNorm prologues (half the words to poll)
    uint2 *ll_q;                        // [8][d]    fp32   cross-attention query
    uint2 *ll_cap;                      // [H][7][8][66] fp32 cross-attention partials (max, sum, o[64]) per key split
    uint2 *ll_hid;                      // [8][2d]   bf16x2 MLP hidden activations
    float* logits; long ld_logits;
    bf16* mkv; long kv_stride;          // KV cache [2Ld][slots][448][d] of this decode lane; elements per [slots][448][d] plane
    int* table; int* tokens;            // KV slot table [8][448]; token histories [8][DEC_TOK_LD] (nullptr with x_in)
    // device-resident decode loop (do_sampling): logit filters + top-k + beam update run in the kernel's tail
    int do_sampling, k; DecodeState* st; DecodeSpec spec; SamplePartials* sp; float* cand_lp; int* cand_tok; int* fin_tokens;
    const int* d_pos; const int* d_done;   // device-resident decode loop: text_offset and completion flag, else nullptr
    const float* mask;                  // reference ABI: additive (449) mask on the device, else nullptr
    const float* x_in;                  // reference ABI: embedded tokens fp32 [nb][d] on the device, else nullptr
    int text_offset;                    // used when d_pos == nullptr
    int no_vocab;                       // stop after the last layer (prompt positions whose logits nobody reads)
    unsigned* barrier;                  // [0] grid-barrier arrivals, [1] CTAs that have left the kernel
    unsigned* seq;                      // launch sequence number (device memory; the kernel increments it)
    long long dbg_delay;                // experiment: cycles the producer waits before it starts streaming
    unsigned long long* dbg;            // optional stage timeline: [n_ctas][MEGA_DBG_LD] %globaltimer values (0 = not reached)
};
void mega_set_model(const MegaModel& m);     // copies the descriptor into the kernel's constant memory
int mega_xs_rows(int nb);
int mega_slots(int nb);
size_t mega_ring_offset(int xs_cols, int nb);
size_t mega_smem_bytes(int xs_cols, int nb);
int mega_sa_cap(int xs_cols, int nb);
bool mega_launch(const MegaArgs& a, int n_ctas, cudaStream_t s);

}  // namespace b200
