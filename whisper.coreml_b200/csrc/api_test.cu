// Test hooks exported through the C ABI (tests/ only; see include/whisper_b200.h).
#include <algorithm>
#include "api_batch.cuh"
#include "gemm.cuh"
#include "ops.cuh"
#include "whisper_b200.h"

using namespace b200;

extern "C" void b200TestGemm(const void* dA, const void* dB, const float* dBias, void* dC, int M, int N, int K,
                             int out_fp32, int gelu, int use_simt) {
    if (use_simt) {
        gemm_simt((const bf16*)dA, (const bf16*)dB, dBias, dC, M, N, K, out_fp32, gelu, 0);
    } else {
        GemmParams p = gemm_plain((const bf16*)dA, (const bf16*)dB, dC, M, N, K);
        p.bias = dBias; p.gelu = gelu; p.c_fp32 = out_fp32;
        gemm_tcgen05(p, 0);
    }
    B200_CHECK(cudaStreamSynchronize(0));
}

namespace b200 { extern int g_gemm_force; }
namespace b200 { void encoder_clear_graphs(); }
extern "C" void b200TestGemmTile(int sel) { b200::g_gemm_force = sel; b200::encoder_clear_graphs(); }    // (a captured encoder graph holds the old tile choice)

// average device time (ms, CUDA events around `iters` back-to-back launches) of one GEMM configuration;
// mode bits: 1 = bias, 2 = GELU, 4 = fp32 output with an fp32 residual added (else bf16 output)
extern "C" float b200TestGemmTime(const void* dA, const void* dB, void* dC, int M, int N, int K, int mode, int iters) {
    GemmParams p = gemm_plain((const bf16*)dA, (const bf16*)dB, dC, M, N, K);
    float *bias = nullptr, *add = nullptr, *cf = nullptr;
    if (mode & 1) { cudaMalloc(&bias, (size_t)N * 4); cudaMemset(bias, 0, (size_t)N * 4); p.bias = bias; }
    p.gelu = (mode & 2) ? 1 : 0;
    if (mode & 4) {
        cudaMalloc(&add, (size_t)M * N * 4); cudaMemset(add, 0, (size_t)M * N * 4);
        cudaMalloc(&cf, (size_t)M * N * 4);
        p.add = add; p.add_rows = M; p.ld_add = N; p.C = cf; p.c_fp32 = 1;
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) gemm_tcgen05(p, 0);
    cudaEventRecord(e0, 0);
    for (int i = 0; i < iters; ++i) gemm_tcgen05(p, 0);
    cudaEventRecord(e1, 0);
    B200_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (bias) cudaFree(bias);
    if (add) cudaFree(add);
    if (cf) cudaFree(cf);
    return ms / (iters > 0 ? iters : 1);
}

// clock64 marks of CTA 0 of one GEMM launch (mode bits as b200TestGemmTime), out[tile * 16 + k]: epilogue warp 2: 0 tile
// entered, 1 accumulator ready, 2+3c / 3+3c / 4+3c chunk c loads issued / landed / chunk done, 14 accumulator released;
// 15 = MMA thread: last MMA of the tile issued.  Returns the number of tiles CTA 0 processed.
namespace b200 { extern unsigned long long* g_gemm_dbg; }
extern "C" int b200TestGemmTimeline(const void* dA, const void* dB, void* dC, int M, int N, int K, int mode, unsigned long long* out, int cap_tiles) {
    GemmParams p = gemm_plain((const bf16*)dA, (const bf16*)dB, dC, M, N, K);
    float *bias = nullptr, *add = nullptr, *cf = nullptr;
    if (mode & 1) { cudaMalloc(&bias, (size_t)N * 4); cudaMemset(bias, 0, (size_t)N * 4); p.bias = bias; }
    p.gelu = (mode & 2) ? 1 : 0;
    if (mode & 4) {
        cudaMalloc(&add, (size_t)M * N * 4); cudaMemset(add, 0, (size_t)M * N * 4);
        cudaMalloc(&cf, (size_t)M * N * 4);
        p.add = add; p.add_rows = M; p.ld_add = N; p.C = cf; p.c_fp32 = 1;
    }
    unsigned long long* d = nullptr;
    cudaMalloc(&d, (size_t)cap_tiles * 16 * 8); cudaMemset(d, 0, (size_t)cap_tiles * 16 * 8);
    for (int i = 0; i < 2; ++i) gemm_tcgen05(p, 0);
    g_gemm_dbg = d;
    gemm_tcgen05(p, 0);
    g_gemm_dbg = nullptr;
    B200_CHECK(cudaStreamSynchronize(0));
    B200_CHECK(cudaMemcpy(out, d, (size_t)cap_tiles * 16 * 8, cudaMemcpyDeviceToHost));
    cudaFree(d);
    if (bias) cudaFree(bias);
    if (add) cudaFree(add);
    if (cf) cudaFree(cf);
    const int tiles = ((M + 127) / 128) * ((N + 255) / 256);
    int n = 0;
    for (int t = 0; t < tiles; t += 148) ++n;
    return n < cap_tiles ? n : cap_tiles;
}

// ---- state read-back hooks (tests only) ---------------------------------------------------------
#include <vector>
#include "state.cuh"

// Xa of window w as fp32 (1500, d) on the host.
extern "C" void b200TestGetXa(float* out, int w) {
    State& s = S();
    if (!s.xa || w < 0 || w >= s.w_cap) { record_error("b200TestGetXa: no encoder output for window %d", w); return; }
    const size_t n = (size_t)N_AUDIO_CTX * s.d;
    float* tmp = nullptr;
    if (!dev_alloc(&tmp, n)) return;
    bf16_to_f32(s.xa + w * n, tmp, (long)n, s.stream);
    B200_CHECK(cudaMemcpy(out, tmp, n * sizeof(float), cudaMemcpyDeviceToHost));
    dev_free(&tmp);
}

// CK (Ld,H,64,1500) [pre-transposed like decoder.py:183] and CV (Ld,H,1500,64) of window w as fp32.
extern "C" void b200TestGetCrossKV(float* out_ck, float* out_cv, int w) {
    State& s = S();
    if (!s.ckv || w < 0 || w >= s.ckv_cap) { record_error("b200TestGetCrossKV: no crossKV for window %d", w); return; }
    const size_t n = s.ckv_window_elems();
    std::vector<bf16> h(n);
    B200_CHECK(cudaMemcpy(h.data(), s.ckv + w * n, n * sizeof(bf16), cudaMemcpyDeviceToHost));
    const size_t hs = (size_t)N_AUDIO_CTX * 64;
    for (int l = 0; l < s.Ld; ++l)
        for (int hh = 0; hh < s.H; ++hh) {
            const bf16* k = h.data() + ((size_t)(l * 2) * s.H + hh) * hs;
            const bf16* v = h.data() + ((size_t)(l * 2 + 1) * s.H + hh) * hs;
            float* ok = out_ck + ((size_t)l * s.H + hh) * hs;
            float* ov = out_cv + ((size_t)l * s.H + hh) * hs;
            for (int t = 0; t < N_AUDIO_CTX; ++t)
                for (int c = 0; c < 64; ++c) {
                    ok[(size_t)c * N_AUDIO_CTX + t] = __bfloat162float(k[(size_t)t * 64 + c]);
                    ov[(size_t)t * 64 + c] = __bfloat162float(v[(size_t)t * 64 + c]);
                }
        }
}

// Logical KV cache rows [0, n_rows) as fp32 (2Ld, bs, n_rows, d), resolved through the slot table.
extern "C" void b200TestGetKV(float* out, int n_rows) {
    State& s = S();
    if (!s.mkv) { record_error("b200TestGetKV: decoder not loaded"); return; }
    const size_t total = (size_t)2 * s.Ld * s.bs * N_TEXT_CTX * s.d;
    std::vector<bf16> h(total);
    std::vector<int> tab((size_t)STEP_MAX_BEAMS * N_TEXT_CTX);
    B200_CHECK(cudaMemcpy(h.data(), s.mkv, total * sizeof(bf16), cudaMemcpyDeviceToHost));
    B200_CHECK(cudaMemcpy(tab.data(), s.table, tab.size() * sizeof(int), cudaMemcpyDeviceToHost));
    for (int m = 0; m < 2 * s.Ld; ++m)
        for (int b = 0; b < s.bs; ++b)
            for (int t = 0; t < n_rows; ++t) {
                const int slot = tab[b * N_TEXT_CTX + t];
                const bf16* src = h.data() + (((size_t)m * s.bs + slot) * N_TEXT_CTX + t) * s.d;
                float* dst = out + (((size_t)m * s.bs + b) * n_rows + t) * s.d;
                for (int c = 0; c < s.d; ++c) dst[c] = __bfloat162float(src[c]);
            }
}

// softmax(Q K^T) V per head over a fused [batch][n_tok][3 * heads * 64] bf16 QKV buffer (device), output
// [batch][n_tok][heads * 64] bf16; tcgen05 kernel or the SIMT checker.
extern "C" void b200TestAttention(const void* dQKV, void* dO, int n_tok, int heads, int batch, int use_simt) {
    use_device();
    const long d = (long)heads * 64;
    AttnParams a{};
    a.Q = (const bf16*)dQKV; a.K = a.Q + d; a.V = a.Q + 2 * d;
    a.ldq = a.ldk = a.ldv = 3 * d; a.q_head_stride = a.k_head_stride = a.v_head_stride = 64;
    a.q_batch_stride = a.k_batch_stride = a.v_batch_stride = (long)n_tok * 3 * d;
    a.O = (bf16*)dO; a.ldo = d; a.o_head_stride = 64; a.o_batch_stride = (long)n_tok * d;
    a.n_q = a.n_k = n_tok; a.n_head = heads; a.batch = batch;
    B200_CHECK(cudaStreamSynchronize(cudaStreamLegacy));
    if (use_simt) attention_simt(a, S().stream); else attention_tc(a, S().stream);
    B200_CHECK(cudaStreamSynchronize(S().stream));
}

// average device time (ms) of `iters` back-to-back tcgen05 flash-attention launches over the same buffers
extern "C" float b200TestAttentionTime(const void* dQKV, void* dO, int n_tok, int heads, int batch, int iters) {
    use_device();
    const long d = (long)heads * 64;
    AttnParams a{};
    a.Q = (const bf16*)dQKV; a.K = a.Q + d; a.V = a.Q + 2 * d;
    a.ldq = a.ldk = a.ldv = 3 * d; a.q_head_stride = a.k_head_stride = a.v_head_stride = 64;
    a.q_batch_stride = a.k_batch_stride = a.v_batch_stride = (long)n_tok * 3 * d;
    a.O = (bf16*)dO; a.ldo = d; a.o_head_stride = 64; a.o_batch_stride = (long)n_tok * d;
    a.n_q = a.n_k = n_tok; a.n_head = heads; a.batch = batch;
    B200_CHECK(cudaStreamSynchronize(cudaStreamLegacy));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) attention_tc(a, S().stream);
    cudaEventRecord(e0, S().stream);
    for (int i = 0; i < iters; ++i) attention_tc(a, S().stream);
    cudaEventRecord(e1, S().stream);
    B200_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return ms / (iters > 0 ? iters : 1);
}

// clock64 marks of CTA (0,0,0) of one flash-attention launch: out[block * 16 + k], k = 0-4 softmax warp (enter, S ready, S loaded,
// P buffer free, P published), 5-7 MMA thread (loop top, next QK issued, P seen); returns the number of key blocks
namespace b200 { extern unsigned long long* g_fa_dbg; }
extern "C" int b200TestAttentionTimeline(const void* dQKV, void* dO, int n_tok, int heads, int batch, unsigned long long* out, int cap_blocks) {
    use_device();
    const int n_kb = (n_tok + 63) / 64;
    unsigned long long* d = nullptr;
    if (!dev_alloc(&d, (size_t)n_kb * 16, true)) return 0;
    g_fa_dbg = d;
    b200TestAttention(dQKV, dO, n_tok, heads, batch, 0);
    g_fa_dbg = nullptr;
    const int n = n_kb < cap_blocks ? n_kb : cap_blocks;
    B200_CHECK(cudaMemcpy(out, d, (size_t)n * 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    dev_free(&d);
    return n;
}

// Stage timeline of the persistent step kernel: enable = 1 allocates/clears the buffer (every following step overwrites
// it: mark k of CTA c = %globaltimer ns at out[c * DB_DBG_LD + k]; mark 0 = start, 2i+1 / 2i+2 = after the prologue /
// body of stage i); enable = 0 copies the last step's marks of up to `cap_ctas` CTAs to `out` and returns the CTA count.
extern "C" int b200TestStepTimeline(int enable, unsigned long long* out, int cap_ctas) {
    use_device();
    return batch_available() ? batch_timeline(enable, out, cap_ctas) : 0;
}
