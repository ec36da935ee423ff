// Non-GEMM device ops shared by the sub-models.
#pragma once
#include "common.cuh"

namespace b200 {

// y[m, :] = LayerNorm(x[m, :]) * gamma + beta     (fp32 in; bf16 and/or fp32 out), one warp per row.
void layernorm(const float* x, const float* gamma, const float* beta, float eps, bf16* y_bf16, float* y_f32, int M,
               int d, cudaStream_t s);
// log-mel (n_mels, total_frames) fp32 -> time-major, zero-framed bf16 rows for the conv stem:
// out[(w*3002 + 1 + t), c] = mel[c, seeks[w] + t] (0 beyond total_frames), c < c_pad; rows 0 and 3001 are zero.
void mel_to_rows(const float* mel, long total_frames, long valid_frames, const int* d_seeks, int n_windows, int n_mels, int c_pad,
                 bf16* out, cudaStream_t s);
void f32_to_bf16(const float* in, bf16* out, long n, cudaStream_t s);
void bf16_to_f32(const bf16* in, float* out, long n, cudaStream_t s);
// dst[r*ld_dst + c] = src[r*ld_src + c], r < rows, c < cols (cols % 8 == 0), bf16
void copy_rows_bf16(const bf16* src, long ld_src, bf16* dst, long ld_dst, int rows, int cols, cudaStream_t s);

struct AttnParams {
    const bf16* Q; long ldq, q_head_stride, q_batch_stride;     // Q[b][h][i][c] = Q + b*bs + h*hs + i*ld + c
    const bf16* K; long ldk, k_head_stride, k_batch_stride;
    const bf16* V; long ldv, v_head_stride, v_batch_stride;
    bf16* O; long ldo, o_head_stride, o_batch_stride;
    int n_q, n_k, n_head, batch;
    const float* mask; long ld_mask;                  // additive fp32 [n_q, ld_mask] or nullptr
    float* qk_dump; const int* dump_slot;             // raw QK (pre-mask, pre-softmax) of head h is written to
    long dump_ld, dump_slot_stride;                   //   qk_dump + dump_slot[h]*dump_slot_stride + i*dump_ld + j (slot < 0: skip)
};
// softmax(Q K^T + mask) V per head; head dim 64; no scaling inside: the reference folds 64^-0.5 into k
// (encoder.py:38) or into the query weights (decoder.py:16-20), and so does the exporter.
void attention_tc(const AttnParams& p, cudaStream_t s);      // tcgen05 flash attention (attention.cu)
void attention_simt(const AttnParams& p, cudaStream_t s);    // general SIMT kernel (masks, QK dump)
void attention_qk_dump(const AttnParams& p, cudaStream_t s); // raw QK of the heads with a dump slot only (batch 1)

}  // namespace b200
