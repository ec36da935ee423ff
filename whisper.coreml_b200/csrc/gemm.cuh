// tcgen05 / TMEM / TMA GEMM used by the encoder, crossKV and decoder256 sub-models.
#pragma once
#include "common.cuh"

namespace b200 {

// C[row(m), n] = act(sum_k A[m,k] * B[n,k] + bias[n]) + add[(m % add_rows), n]
//
// A is addressed through up to three 3-D TMA tensor maps (inner = K slice, rows, batch):
//   k-block kb reads map[kb / kblocks_per_map] at inner coordinate (kb % kblocks_per_map) * 64.
//   A plain row-major matrix uses one map with batch = 1.  The conv stem uses three maps (one per
//   filter tap, base pointer shifted by one input row each) so that no im2col copy is ever made.
// The M dimension is `batch` groups of `rows_per_batch` rows; tiles never straddle groups.
// Output row of (b, t) is b * c_batch_rows + c_row0 + t.
struct GemmParams {
    const bf16* A[3];           // base pointer of each A map
    int num_a_maps;             // 1 or 3
    int kblocks_per_map;        // K blocks (of 64) taken from each map
    long a_inner;               // extent of the inner (K) dimension of each A map (elements)
    long a_row_stride;          // elements between consecutive rows
    long a_batch_stride;        // elements between batches
    int rows_per_batch;
    int batch;
    const bf16* B;              // [N, ldb] row-major (nn.Linear weight layout)
    long ldb;
    int N;
    int K;                      // total K (multiple of 64 after padding by the exporter)
    const float* bias;          // [N] or nullptr
    int gelu;
    const float* add;           // fp32 [add_rows, ld_add] or nullptr (residual / positional embedding)
    int add_rows;
    long ld_add;
    void* C;
    int c_fp32;                 // 1: fp32 output, 0: bf16 output
    long ldc;
    int c_batch_rows;
    int c_row0;
    int c_split;                // 1: column n goes to (n / 64) * c_split_stride + (n % 64) (head-major output, ldc = 64)
    long c_split_stride;
    // optional second bf16 output for the crossKV projection: per 64-column group (layer, k|v, head) a fragment-major
    // copy for the decoder step kernel - K as [keys/16][2][1 KB] tiles, V transposed as [4][c2_keys/32][1 KB] tiles
    // (decoder_batch.cu).  Group g is a V group when (g / c2_heads) is odd.  Rows are keys (row index within the batch).
    bf16* C2;
    long c2_batch_stride;
    int c2_heads;
    int c2_keys;                // padded key count (multiple of 32)
};

void gemm_tcgen05(const GemmParams& p, cudaStream_t stream);
// Plain helper: C[M,N] = A[M,K] B[N,K]^T with contiguous A (lda = K).
GemmParams gemm_plain(const bf16* A, const bf16* B, void* C, int M, int N, int K);

// SIMT checker (tests only)
void gemm_simt(const bf16* A, const bf16* B, const float* bias, void* C, int M, int N, int K, int c_fp32, int gelu,
               cudaStream_t stream);

}  // namespace b200
