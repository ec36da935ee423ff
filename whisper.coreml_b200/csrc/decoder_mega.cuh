// Persistent decoder1 step kernel (decoder_mega.cu): interface.
#pragma once
#include "common.cuh"
#include "sampling.cuh"
#include "state.cuh"

namespace b200 {

constexpr int MEGA_MAX_LAYERS = 32;

struct MegaLayer {
    const bf16 *qkv, *attn_out, *cross_q, *cross_out, *mlp1, *mlp2;                  // fragment-major weights
    const float *qkv_b, *attn_out_b, *cross_q_b, *cross_out_b, *mlp1_b, *mlp2_b;
    const float *ln1_w, *ln1_b, *ln2_w, *ln2_b, *ln3_w, *ln3_b;
    bf16 *cache_k, *cache_v;                                                          // [slots][448][d]
};
struct MegaModel {                      // lives in device memory, built when both decoders are loaded
    int d, H, Ld, V, n_tiles_vocab;
    const bf16* tok_emb; const bf16* tok_emb_frag; const float* pos_emb; const float* ln_w; const float* ln_b;
    MegaLayer layers[MEGA_MAX_LAYERS];
};
struct MegaArgs {
    const MegaModel* model;
    const bf16* ckv_frag;               // fragment-major cross K / V^T of the current window
    int nb, k, xs_cols;
    int n_slots;                        // ring slots in use (<= 20): bytes in flight per SM = n_slots * 8 KB
    float* xb[2]; float* part_qkv; float* part_q; bf16* attn; bf16* hid; float* part_m2; float* logits; long ld_logits;
    float* ca_part; int* ca_counters;
    int* table; int* tokens; DecodeState* st; DecodeSpec spec; SamplePartials* sp; float* cand_lp; int* cand_tok; int* fin_tokens;
    const float* mask;                  // reference ABI: additive (449) mask on the device, else nullptr
    const float* x_in;                  // reference ABI: embedded tokens fp32 [nb][d] on the device, else nullptr
    int text_offset;                    // used when st == nullptr
    int do_sampling;
    unsigned* barrier;                  // [0] arrivals, [1] generation
    unsigned long long* dbg;            // optional: %globaltimer of CTA 0 after every grid barrier (stage timeline), [0] = count
};
size_t mega_smem_bytes(int xs_cols);
bool mega_launch(const MegaArgs& a, int n_ctas, cudaStream_t s);

}  // namespace b200
