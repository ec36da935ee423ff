#include "timing.cuh"

#include <vector>

#include "gemm.cuh"
#include "ops.cuh"
#include "state.cuh"
#include "whisper_b200.h"

namespace b200 {

// ---------------------------------------------------------------------------------------------------
// median filter (timing.py:19-54): reflect padding without edge repeat, middle of the sorted window
// ---------------------------------------------------------------------------------------------------
constexpr int MED_MAX_W = 63;

__device__ __forceinline__ float median_at(const float* __restrict__ xr, int t, int len, int width) {
    float win[MED_MAX_W];
    const int pad = width >> 1;
    for (int k = 0; k < width; ++k) {
        int s = t + k - pad;
        if (s < 0) s = -s;
        if (s >= len) s = 2 * (len - 1) - s;
        const float v = xr[s];
        int p = k - 1;                                   // insertion sort
        while (p >= 0 && win[p] > v) { win[p + 1] = win[p]; --p; }
        win[p + 1] = v;
    }
    return win[pad];
}

template <int W>
__device__ __forceinline__ float median_fixed(const float* __restrict__ xr, int t, int len) {
    float win[W];
    constexpr int pad = W / 2;
#pragma unroll
    for (int k = 0; k < W; ++k) {
        int s = t + k - pad;
        if (s < 0) s = -s;
        if (s >= len) s = 2 * (len - 1) - s;
        win[k] = xr[s];
    }
#pragma unroll
    for (int i = 0; i < W; ++i)                          // odd-even transposition network, fully unrolled
#pragma unroll
        for (int j = (i & 1); j + 1 < W; j += 2) {
            const float a = win[j], b = win[j + 1];
            win[j] = fminf(a, b); win[j + 1] = fmaxf(a, b);
        }
    return win[pad];
}

__global__ void median_kernel(const float* __restrict__ x, float* __restrict__ y, long rows, int len, int width) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * len) return;
    const long r = i / len; const int t = (int)(i % len);
    const float* xr = x + r * len;
    if (len <= width / 2) { y[i] = xr[t]; return; }      // timing.py:22-24
    y[i] = width == 7 ? median_fixed<7>(xr, t, len) : median_at(xr, t, len, width);
}

void median_filter_dev(const float* x, float* y, long rows, int len, int width, cudaStream_t s) {
    if (width < 1 || width > MED_MAX_W || (width & 1) == 0) { record_error("medianFilter: width %d must be odd and <= %d", width, MED_MAX_W); return; }
    const long n = rows * len;
    if (n <= 0) return;
    median_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, y, rows, len, width);
    B200_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------
// DTW (timing.py:57-105).  One CTA, thread r owns row r of the cost matrix and walks the anti-
// diagonals; a row only needs its left neighbour (own register) and the two latest values of the
// row above (shared memory).  The 2-bit trace lives in shared memory when it fits (<= 200 KB), so
// the serial backtrace never touches DRAM.
// ---------------------------------------------------------------------------------------------------
constexpr int DTW_SMEM_TRACE = 200 * 1024;

template <bool SMEM_TRACE>
__global__ void __launch_bounds__(1024) dtw_kernel(const float* __restrict__ x, int N, int M, int* __restrict__ out_i,
                                                   int* __restrict__ out_j, int* __restrict__ out_len, uint8_t* gtrace) {
    extern __shared__ __align__(16) uint8_t dsm[];
    float* ring = reinterpret_cast<float*>(dsm);         // [3][N + 1]
    uint8_t* trace = SMEM_TRACE ? dsm + (size_t)3 * (N + 1) * sizeof(float) : gtrace;   // [(N * M + 3) / 4] 2-bit cells
    const int nthr = blockDim.x;
    const int rows_per = (N + nthr - 1) / nthr;          // > 1 only when N > 1024
    (void)rows_per;
    const int r = threadIdx.x + 1;                       // 1-based row
    const bool has_row = r <= N;
    const size_t tbytes = ((size_t)N * M + 3) / 4;
    for (size_t i = threadIdx.x; i < tbytes; i += nthr) trace[i] = 0;
    for (int i = threadIdx.x; i < 3 * (N + 1); i += nthr) ring[i] = INFINITY;
    __syncthreads();
    float left = INFINITY;                               // cost[r][c - 1]
    const float* xr = x + (size_t)(r - 1) * M;
    float xnext = (has_row && r == 1) ? xr[0] : 0.f;
    for (int d = 2; d <= N + M; ++d) {                   // d = r + c
        float* cur = ring + (d % 3) * (N + 1);
        const float* p1 = ring + ((d + 2) % 3) * (N + 1);   // values of diagonal d - 1
        const float* p2 = ring + ((d + 1) % 3) * (N + 1);   // values of diagonal d - 2
        const int c = d - r;
        if (has_row && c >= 1 && c <= M) {
            const float xv = xnext;
            float c0, c1;
            if (r == 1) { c0 = c == 1 ? 0.f : INFINITY; c1 = INFINITY; }
            else { c0 = c == 1 ? INFINITY : p2[r - 1]; c1 = p1[r - 1]; }
            const float c2 = left;
            float cc; int t;
            if (c0 < c1 && c0 < c2) { cc = c0; t = 0; }
            else if (c1 < c0 && c1 < c2) { cc = c1; t = 1; }
            else { cc = c2; t = 2; }
            const float v = (float)((double)xv + (double)cc);   // fp64 add stored as fp32 (timing.py:85,102)
            cur[r] = v; left = v;
            const size_t cell = (size_t)(r - 1) * M + (c - 1);
            // rows are owned by one thread, but 4 cells share a byte across the row boundary only when M % 4 != 0
            atomicOr(reinterpret_cast<unsigned int*>(trace + (cell >> 2 & ~(size_t)3)),
                     (unsigned int)t << (((cell >> 2) & 3) * 8 + (cell & 3) * 2));
        }
        if (has_row && c + 1 >= 1 && c + 1 <= M) xnext = xr[c];   // prefetch x[r-1][c] for the next diagonal
        __syncthreads();
    }
    if (threadIdx.x == 0) {                              // backtrace (timing.py:57-79)
        int i = N, j = M, n = 0;
        while (i > 0 || j > 0) {
            out_i[n] = i - 1; out_j[n] = j - 1; ++n;
            int t;
            if (j == 0) t = 1; else if (i == 0) t = 2;
            else { const size_t cell = (size_t)(i - 1) * M + (j - 1); t = (trace[cell >> 2] >> ((cell & 3) * 2)) & 3; }
            if (t == 0) { --i; --j; } else if (t == 1) --i; else --j;
        }
        for (int a = 0, b = n - 1; a < b; ++a, --b) {
            int tmp = out_i[a]; out_i[a] = out_i[b]; out_i[b] = tmp;
            tmp = out_j[a]; out_j[a] = out_j[b]; out_j[b] = tmp;
        }
        *out_len = n;
    }
}

size_t dtw_scratch_bytes(int N, int M) { return ((size_t)N * M + 3) / 4 + 16; }

void dtw_dev(const float* x, int N, int M, int* d_path_i, int* d_path_j, int* d_len, uint8_t* scratch, cudaStream_t s) {
    if (N < 1 || M < 1 || N > 1024) { record_error("dtw: N=%d M=%d unsupported (1 <= N <= 1024)", N, M); return; }
    const int threads = ((N + 31) / 32) * 32;
    const size_t ring = (size_t)3 * (N + 1) * sizeof(float);
    const size_t tbytes = (((size_t)N * M + 3) / 4 + 3) & ~(size_t)3;
    if (ring + tbytes <= DTW_SMEM_TRACE) {
        static size_t attr = 0;
        if (ring + tbytes > attr) {
            B200_CHECK(cudaFuncSetAttribute(dtw_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DTW_SMEM_TRACE));
            attr = DTW_SMEM_TRACE;
        }
        dtw_kernel<true><<<1, threads, ring + tbytes, s>>>(x, N, M, d_path_i, d_path_j, d_len, nullptr);
    } else {
        dtw_kernel<false><<<1, threads, ring, s>>>(x, N, M, d_path_i, d_path_j, d_len, scratch);
    }
    B200_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------
// find_alignment numerics (timing.py:194-204)
// ---------------------------------------------------------------------------------------------------
// w[h][t][0..F) = softmax over frames (qk_scale = 1), one warp per (head, token) row
__global__ void align_softmax_kernel(float* __restrict__ chw, int heads, int ld_tok, int n_tok, int F) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= heads * n_tok) return;
    float* p = chw + ((size_t)(row / n_tok) * ld_tok + row % n_tok) * N_AUDIO_CTX;
    float m = -INFINITY;
    for (int j = lane; j < F; j += 32) m = fmaxf(m, p[j]);
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < F; j += 32) { const float e = expf(p[j] - m); p[j] = e; s += e; }
    s = warp_sum(s);
    for (int j = lane; j < F; j += 32) p[j] = p[j] / s;
}
// z-score over tokens (std_mean(dim=-2, unbiased=False)), one thread per (head, frame); out [heads][n_tok][F]
__global__ void align_norm_kernel(const float* __restrict__ chw, int heads, int ld_tok, int n_tok, int F, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= heads * F) return;
    const int h = i / F, f = i % F;
    const float* p = chw + (size_t)h * ld_tok * N_AUDIO_CTX + f;
    float mean = 0.f;
    for (int t = 0; t < n_tok; ++t) mean += p[(size_t)t * N_AUDIO_CTX];
    mean /= n_tok;
    float var = 0.f;
    for (int t = 0; t < n_tok; ++t) { const float dd = p[(size_t)t * N_AUDIO_CTX] - mean; var = fmaf(dd, dd, var); }
    const float sd = sqrtf(var / n_tok);
    for (int t = 0; t < n_tok; ++t) out[((size_t)h * n_tok + t) * F + f] = (p[(size_t)t * N_AUDIO_CTX] - mean) / sd;
}
// mean over heads of rows [n_skip, n_tok - 1), optionally negated
__global__ void align_mean_kernel(const float* __restrict__ w, int heads, int n_tok, int F, int n_skip, int n_rows, bool negate,
                                  float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows * F) return;
    const int r = i / F, f = i % F;
    float s = 0.f;
    for (int h = 0; h < heads; ++h) s += w[((size_t)h * n_tok + n_skip + r) * F + f];
    s /= heads;
    out[i] = negate ? -s : s;
}

void alignment_matrix_dev(float* chw, int heads, int ld_tok, int n_tok, int F, int n_skip, int width, float* tmp, float* out,
                          bool negate, cudaStream_t s) {
    const int n_rows = n_tok - 1 - n_skip;
    if (n_rows < 1 || heads < 1 || F < 1) { record_error("alignment: nothing to align (tokens %d, skip %d, heads %d)", n_tok, n_skip, heads); return; }
    align_softmax_kernel<<<cdiv(heads * n_tok, 8), 256, 0, s>>>(chw, heads, ld_tok, n_tok, F);
    B200_LAUNCH_CHECK();
    float* z = tmp; float* zm = tmp + (size_t)heads * n_tok * F;
    align_norm_kernel<<<cdiv(heads * F, 128), 128, 0, s>>>(chw, heads, ld_tok, n_tok, F, z);
    B200_LAUNCH_CHECK();
    median_filter_dev(z, zm, (long)heads * n_tok, F, width, s);
    align_mean_kernel<<<cdiv(n_rows * F, 256), 256, 0, s>>>(zm, heads, n_tok, F, n_skip, n_rows, negate, out);
    B200_LAUNCH_CHECK();
}

// softmax over [0, eot) of logits row r, probability of tokens[n_skip + r + 1]... see b200AlignTokens
__global__ void token_prob_kernel(const float* __restrict__ logits, long ld, int eot, const int* __restrict__ targets,
                                  float* __restrict__ out) {
    __shared__ float red[32];
    const float* x = logits + (size_t)blockIdx.x * ld;
    float m = -INFINITY;
    for (int v = threadIdx.x; v < eot; v += 1024) m = fmaxf(m, x[v]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = red[0];
    for (int i = 1; i < 32; ++i) m = fmaxf(m, red[i]);
    __syncthreads();
    float s = 0.f;
    for (int v = threadIdx.x; v < eot; v += 1024) s += expf(x[v] - m);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int i = 0; i < 32; ++i) tot += red[i];
        const int tk = targets[blockIdx.x];
        out[blockIdx.x] = (tk >= 0 && tk < eot) ? expf(x[tk] - m) / tot : 0.f;
    }
}

}  // namespace b200

using namespace b200;

extern "C" {

void medianFilter(const float* x, float* y, long rows, int len, int width) {
    use_device();
    const size_t n = (size_t)rows * len;
    if (!n) return;
    float *dx = nullptr, *dy = nullptr;
    if (!dev_alloc(&dx, n) || !dev_alloc(&dy, n)) { dev_free(&dx); return; }
    cudaStream_t st = S().stream;
    B200_CHECK(cudaMemcpyAsync(dx, x, n * sizeof(float), cudaMemcpyHostToDevice, st));
    median_filter_dev(dx, dy, rows, len, width, st);
    B200_CHECK(cudaMemcpyAsync(y, dy, n * sizeof(float), cudaMemcpyDeviceToHost, st));
    B200_CHECK(cudaStreamSynchronize(st));
    dev_free(&dx); dev_free(&dy);
}

int dtw(const float* x, int N, int M, int* out_i, int* out_j) {
    use_device();
    if (N < 1 || M < 1) return 0;
    float* dx = nullptr; int* dp = nullptr; uint8_t* scratch = nullptr;
    const size_t n = (size_t)N * M;
    if (!dev_alloc(&dx, n) || !dev_alloc(&dp, (size_t)2 * (N + M) + 1) || !dev_alloc(&scratch, dtw_scratch_bytes(N, M))) {
        dev_free(&dx); dev_free(&dp); dev_free(&scratch); return 0;
    }
    cudaStream_t st = S().stream;
    B200_CHECK(cudaMemcpyAsync(dx, x, n * sizeof(float), cudaMemcpyHostToDevice, st));
    B200_CHECK(cudaMemsetAsync(dp + 2 * (N + M), 0, sizeof(int), st));
    dtw_dev(dx, N, M, dp, dp + (N + M), dp + 2 * (N + M), scratch, st);
    int len = 0;
    B200_CHECK(cudaMemcpyAsync(&len, dp + 2 * (N + M), sizeof(int), cudaMemcpyDeviceToHost, st));
    B200_CHECK(cudaStreamSynchronize(st));
    if (len > 0 && len <= N + M) {
        B200_CHECK(cudaMemcpy(out_i, dp, (size_t)len * sizeof(int), cudaMemcpyDeviceToHost));
        B200_CHECK(cudaMemcpy(out_j, dp + (N + M), (size_t)len * sizeof(int), cudaMemcpyDeviceToHost));
    }
    dev_free(&dx); dev_free(&dp); dev_free(&scratch);
    return len;
}

}  // extern "C"
