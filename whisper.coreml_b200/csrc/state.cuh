// Process-global model state behind the C ABI (the analogue of the MLMultiArray globals of
// coreml/coreml.mm:18-23): weights, activation workspaces, Xa, CK/CV and the 448-slot KV cache.
#pragma once
#include <string>
#include <vector>

#include "common.cuh"
#include "weights.cuh"

namespace b200 {

constexpr int N_AUDIO_CTX = 1500, N_FRAMES = 3000, N_TEXT_CTX = 448, PREFILL_CTX = 256;
constexpr int STEP_MAX_BEAMS = 8;         // KV-cache slots / beams of one decode (loadDecoder256's beam_size)
constexpr int CROSS_KEYS_PAD = 1504;      // 1500 audio keys padded to a multiple of 32 for the fragment-major copy

struct EncLayer {
    const float *attn_ln_w, *attn_ln_b, *qkv_b, *out_b, *mlp_ln_w, *mlp_ln_b, *mlp1_b, *mlp2_b;
    const bf16 *qkv_w, *out_w, *mlp1_w, *mlp2_w;
};
struct DecLinear { const bf16* w; const bf16* frag; const float* b; };
struct DecLayer {
    const float *attn_ln_w, *attn_ln_b, *cross_ln_w, *cross_ln_b, *mlp_ln_w, *mlp_ln_b;
    DecLinear qkv, attn_out, cross_q, cross_out, mlp1, mlp2;
};


struct State {
    int device = 0;
    cudaStream_t stream = nullptr;    // library stream (non-blocking, capturable); created by use_device()

    // ---- encoder -------------------------------------------------------------------------------
    bool enc_loaded = false;
    WeightFile enc_w;
    int d = 0, H = 0, Le = 0, n_mels = 0, cpad = 0;
    const bf16 *conv1_w = nullptr, *conv2_w = nullptr;
    const float *conv1_b = nullptr, *conv2_b = nullptr, *pos = nullptr, *ln_post_w = nullptr, *ln_post_b = nullptr;
    std::vector<EncLayer> enc_layers;
    int w_cap = 0;                    // windows the workspaces are sized for
    int n_windows = 0;                // windows of the last encoder call
    bf16 *melrows = nullptr, *h1 = nullptr, *y = nullptr, *qkv = nullptr, *att = nullptr, *hid = nullptr, *xa = nullptr;
    float* x = nullptr;
    float* mel_stage = nullptr;       // (n_mels, 3000) fp32 staging for encoderPredict
    int* d_seeks = nullptr;

    // ---- crossKV -------------------------------------------------------------------------------
    bool ckv_loaded = false;
    WeightFile ckv_w;
    int Ld = 0;
    const bf16* ckv_wt = nullptr; const float* ckv_b = nullptr;
    bf16* ckv = nullptr;              // [w_cap][Ld][2][H][1500][64]  row-major copy (prefill / alignment attention)
    bf16* ckv_frag = nullptr;         // [w_cap][Ld][2][H][64 * 1504]  fragment-major copy (decoder step kernel)
    int ckv_cap = 0;
    int cur_window = 0;

    // ---- decoder -------------------------------------------------------------------------------
    WeightFile dec_w;                 // shared by decoder256 and decoder1
    bool dec256_loaded = false, dec1_loaded = false;
    int V = 0, bs = 0, n_align = 0;
    const bf16 *tok_emb = nullptr, *tok_emb_frag = nullptr;
    const float *pos_emb = nullptr, *ln_w = nullptr, *ln_b = nullptr;
    std::vector<DecLayer> dec_layers;
    bf16* mkv = nullptr;              // [2Ld][bs][448][d]
    int* table = nullptr;             // [bs][448] logical (beam, pos) -> physical slot
    std::vector<int> align_heads;     // (layer, head) pairs in CHW row order; empty = default
    int* d_dump_slot = nullptr;       // [Ld][H] CHW row of each head or -1
    // prefill workspace (256 rows)
    float *px = nullptr, *pmask = nullptr, *pchw = nullptr, *pout = nullptr;
    bf16 *py = nullptr, *pqkv = nullptr, *patt = nullptr, *phid = nullptr, *pq = nullptr;
    // reference-ABI step (decoder1Predict): logits out, mask / embedded tokens in
    float *slogits = nullptr, *smask = nullptr, *sxin = nullptr;
    int smem_optin = 0;
    int n_sms = 0;
    float* pin_logits = nullptr;      // pinned host staging
    float* pin_x = nullptr;

    size_t ckv_window_elems() const { return (size_t)Ld * 2 * H * N_AUDIO_CTX * 64; }
    size_t ckv_frag_window_elems() const { return (size_t)Ld * 2 * H * CROSS_KEYS_PAD * 64; }
    const bf16* ckf_ptr(int w, int l) const { return ckv_frag + w * ckv_frag_window_elems() + (size_t)(l * 2) * H * CROSS_KEYS_PAD * 64; }
    const bf16* cvf_ptr(int w, int l) const { return ckv_frag + w * ckv_frag_window_elems() + (size_t)(l * 2 + 1) * H * CROSS_KEYS_PAD * 64; }
    bf16* ck_ptr(int w, int l) const { return ckv + w * ckv_window_elems() + (size_t)(l * 2) * H * N_AUDIO_CTX * 64; }
    bf16* cv_ptr(int w, int l) const { return ckv + w * ckv_window_elems() + (size_t)(l * 2 + 1) * H * N_AUDIO_CTX * 64; }
    bf16* mk_ptr(int l) const { return mkv + (size_t)(2 * l) * bs * N_TEXT_CTX * d; }
    bf16* mv_ptr(int l) const { return mkv + (size_t)(2 * l + 1) * bs * N_TEXT_CTX * d; }
};

State& S();

template <typename T>
bool dev_alloc(T** p, size_t n, bool zero = false) {
    if (*p) { cudaFree(*p); *p = nullptr; }
    cudaError_t e = cudaMalloc((void**)p, n * sizeof(T) > 0 ? n * sizeof(T) : 256);
    if (e != cudaSuccess) { record_error("cudaMalloc(%zu bytes): %s", n * sizeof(T), cudaGetErrorString(e)); *p = nullptr; return false; }
    if (zero) B200_CHECK(cudaMemset(*p, 0, n * sizeof(T)));
    return true;
}
template <typename T>
void dev_free(T** p) { if (*p) { B200_CHECK(cudaFree(*p)); *p = nullptr; } }

void use_device();

// per-stage device time (CUDA events), the analogue of whisper/coreml.py:9-13
enum Stage { ST_MEL = 0, ST_ENCODER, ST_CROSSKV, ST_DECODER256, ST_DECODER1, ST_SAMPLING, ST_ALIGN, ST_COUNT };
struct StageTimer {
    StageTimer(int stage);
    ~StageTimer();
    int stage; cudaEvent_t e0, e1;
};
void stage_times(float* out_ms, bool reset);

// sub-model bodies (api.cu)
bool ensure_encoder_capacity(int n_windows);
void run_encoder(const float* d_mel, long total_frames, long valid_frames, int n_windows);   // seeks already in S().d_seeks
void run_cross_kv(int n_windows);
void run_prefill(int beam_idx, bool want_chw, int rows);                            // px/pmask -> pout (+ pchw), KV rows -> slot
}  // namespace b200
