// Word-timestamp kernels: median filter, DTW with the reference's CPU tie rule, and the
// find_alignment pre-processing (whisper/timing.py:19-54, 57-105, 194-204).
#pragma once
#include "common.cuh"

namespace b200 {

// y = median filter of odd `width` along the last axis with reflect padding; x, y: [rows][len] fp32 (device)
void median_filter_dev(const float* x, float* y, long rows, int len, int width, cudaStream_t s);
// DTW over the device cost matrix x [N][M]; path written to d_path (2 * (N + M) ints: i's then j's) in
// forward order, its length to d_len.  scratch: >= (N + 1) * (M + 1) bytes when the trace does not fit in smem.
void dtw_dev(const float* x, int N, int M, int* d_path_i, int* d_path_j, int* d_len, uint8_t* scratch, cudaStream_t s);
size_t dtw_scratch_bytes(int N, int M);

// raw QK [heads][ld_tok rows][1500] -> alignment matrix [n_tok - 1 - n_skip][F], negated for DTW if `negate`
void alignment_matrix_dev(float* chw, int heads, int ld_tok, int n_tok, int F, int n_skip, int width, float* tmp,
                          float* out, bool negate, cudaStream_t s);

}  // namespace b200
