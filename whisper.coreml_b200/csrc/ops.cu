// LayerNorm, layout/precision conversion and the general SIMT attention kernel.
#include "ops.cuh"

namespace b200 {

// ---------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row kept in registers (d <= 1536, d % 128 == 0), two-pass variance.  At most 85 registers
// per thread: three CTAs per SM, so the 375 CTAs of two windows (3000 rows) are ONE wave (at 88 registers they were 1.27 waves of 296).
// ---------------------------------------------------------------------------------------------------
template <int VPL>   // float4 per lane
__global__ void __launch_bounds__(256, 3) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, bf16* __restrict__ yb,
                                                        float* __restrict__ yf, int M, int d) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float4* xr = reinterpret_cast<const float4*>(x + (long)row * d);
    float4 v[VPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        v[i] = xr[lane + 32 * i];
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float rstd = rsqrtf(warp_sum(q) / d + eps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c4 = lane + 32 * i;
        const float4 g = reinterpret_cast<const float4*>(gamma)[c4], b = reinterpret_cast<const float4*>(beta)[c4];
        float4 o;
        o.x = v[i].x * rstd * g.x + b.x; o.y = v[i].y * rstd * g.y + b.y;
        o.z = v[i].z * rstd * g.z + b.z; o.w = v[i].w * rstd * g.w + b.w;
        if (yb) {
            uint2 pk; pk.x = pack_bf16(o.x, o.y); pk.y = pack_bf16(o.z, o.w);
            reinterpret_cast<uint2*>(yb + (long)row * d)[c4] = pk;
        }
        if (yf) reinterpret_cast<float4*>(yf + (long)row * d)[c4] = o;
    }
}

void layernorm(const float* x, const float* gamma, const float* beta, float eps, bf16* yb, float* yf, int M, int d,
               cudaStream_t s) {
    if (d % 128 != 0 || d > 1536) { record_error("layernorm: unsupported width %d", d); return; }
    const int grid = cdiv(M, 8);
    switch (d / 128) {
#define LN_CASE(V) case V: layernorm_kernel<V><<<grid, 256, 0, s>>>(x, gamma, beta, eps, yb, yf, M, d); break;
        LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(7) LN_CASE(8) LN_CASE(9) LN_CASE(10)
        LN_CASE(11) LN_CASE(12)
#undef LN_CASE
    }
    B200_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------
// mel (channel-major fp32) -> zero-framed time-major bf16 rows; 32x32 smem transpose.
// ---------------------------------------------------------------------------------------------------
// Frames at or past `valid_frames` (the file's content, whisper/transcribe.py:286-290: the window is sliced to the content and
// pad_or_trim ZERO-pads it) and past total_frames read as zeros.
__global__ void mel_to_rows_kernel(const float* __restrict__ mel, long total_frames, long valid_frames, const int* __restrict__ seeks,
                                   int n_mels, int c_pad, bf16* __restrict__ out) {
    __shared__ float tile[32][33];
    const int w = blockIdx.z;
    const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const long seek = seeks[w];
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {       // i: channel within tile, threadIdx.x: time
        const int c = c0 + i;
        const long f = seek + t0 + threadIdx.x;
        float v = 0.f;
        if (c < n_mels && t0 + threadIdx.x < 3000 && f < valid_frames) v = mel[(long)c * total_frames + f];
        tile[i][threadIdx.x] = v;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {       // i: time within tile, threadIdx.x: channel
        const int t = t0 + i, c = c0 + threadIdx.x;
        if (t < 3000 && c < c_pad) out[((long)w * 3002 + 1 + t) * c_pad + c] = __float2bfloat16(tile[threadIdx.x][i]);
    }
    if (blockIdx.x == 0 && threadIdx.y == 0) {                 // frame rows 0 and 3001
        const int c = c0 + threadIdx.x;
        if (c < c_pad) {
            out[((long)w * 3002) * c_pad + c] = __float2bfloat16(0.f);
            out[((long)w * 3002 + 3001) * c_pad + c] = __float2bfloat16(0.f);
        }
    }
}

void mel_to_rows(const float* mel, long total_frames, long valid_frames, const int* d_seeks, int n_windows, int n_mels, int c_pad, bf16* out,
                 cudaStream_t s) {
    dim3 grid(cdiv(3000, 32), cdiv(c_pad, 32), n_windows), block(32, 8);
    mel_to_rows_kernel<<<grid, block, 0, s>>>(mel, total_frames, valid_frames < total_frames ? valid_frames : total_frames, d_seeks, n_mels, c_pad, out);
    B200_LAUNCH_CHECK();
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, long n) {
    long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        const float4 v = *reinterpret_cast<const float4*>(in + i);
        uint2 pk; pk.x = pack_bf16(v.x, v.y); pk.y = pack_bf16(v.z, v.w);
        *reinterpret_cast<uint2*>(out + i) = pk;
    } else {
        for (; i < n; ++i) out[i] = __float2bfloat16(in[i]);
    }
}
void f32_to_bf16(const float* in, bf16* out, long n, cudaStream_t s) {
    if (n <= 0) return;
    f32_to_bf16_kernel<<<(unsigned)((n / 4 + 256) / 256), 256, 0, s>>>(in, out, n);
    B200_LAUNCH_CHECK();
}
__global__ void bf16_to_f32_kernel(const bf16* __restrict__ in, float* __restrict__ out, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __bfloat162float(in[i]);
}
void bf16_to_f32(const bf16* in, float* out, long n, cudaStream_t s) {
    if (n <= 0) return;
    bf16_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in, out, n);
    B200_LAUNCH_CHECK();
}

__global__ void copy_rows_bf16_kernel(const bf16* __restrict__ src, long ld_src, bf16* __restrict__ dst, long ld_dst,
                                      int rows, int cols8) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)rows * cols8) return;
    const int r = (int)(i / cols8), c = (int)(i % cols8) * 8;
    *reinterpret_cast<uint4*>(dst + r * ld_dst + c) = *reinterpret_cast<const uint4*>(src + r * ld_src + c);
}
void copy_rows_bf16(const bf16* src, long ld_src, bf16* dst, long ld_dst, int rows, int cols, cudaStream_t s) {
    const long n = (long)rows * (cols / 8);
    if (n <= 0) return;
    copy_rows_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, ld_src, dst, ld_dst, rows, cols / 8);
    B200_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------
// General SIMT attention: one thread per query row, K/V tiles of 32 keys staged in smem as fp32.
// Handles additive masks (incl. -inf), fully-masked rows (output 0 where torch would give NaN is
// avoided by the reference's masks: every prefill row keeps column 0), and the raw-QK dump used by
// the alignment heads (decoder.py:306-308).  Used for the decoder256 prefill (256 queries) and as the
// checker of the tcgen05 kernel.
// ---------------------------------------------------------------------------------------------------
constexpr int SA_Q = 64;     // queries per block
constexpr int SA_K = 32;     // keys per smem tile

__global__ void __launch_bounds__(SA_Q) attention_simt_kernel(const AttnParams p) {
    __shared__ float sk[SA_K][64];
    __shared__ float sv[SA_K][64];
    const int h = blockIdx.y, b = blockIdx.z;
    const int i = blockIdx.x * SA_Q + threadIdx.x;
    const bool active = i < p.n_q;
    float q[64], o[64];
    if (active) {
        const bf16* qp = p.Q + b * p.q_batch_stride + h * p.q_head_stride + (long)i * p.ldq;
#pragma unroll
        for (int c = 0; c < 64; c += 8) {
            const uint4 u = *reinterpret_cast<const uint4*>(qp + c);
            q[c] = bf16lo(u.x); q[c + 1] = bf16hi(u.x); q[c + 2] = bf16lo(u.y); q[c + 3] = bf16hi(u.y);
            q[c + 4] = bf16lo(u.z); q[c + 5] = bf16hi(u.z); q[c + 6] = bf16lo(u.w); q[c + 7] = bf16hi(u.w);
        }
    }
#pragma unroll
    for (int c = 0; c < 64; ++c) o[c] = 0.f;
    float m = -INFINITY, l = 0.f;
    const int slot = (p.qk_dump && p.dump_slot) ? p.dump_slot[h] : -1;
    float* dump = slot >= 0 && active ? p.qk_dump + (long)slot * p.dump_slot_stride + (long)i * p.dump_ld : nullptr;
    const float* mrow = (p.mask && active) ? p.mask + (long)i * p.ld_mask : nullptr;
    const bf16* kb = p.K + b * p.k_batch_stride + h * p.k_head_stride;
    const bf16* vb = p.V + b * p.v_batch_stride + h * p.v_head_stride;

    for (int j0 = 0; j0 < p.n_k; j0 += SA_K) {
        __syncthreads();
        for (int e = threadIdx.x; e < SA_K * 8; e += SA_Q) {      // 8 x 16-byte pieces per key row
            const int jj = e >> 3, c = (e & 7) * 8;
            uint4 uk = make_uint4(0, 0, 0, 0), uv = make_uint4(0, 0, 0, 0);
            if (j0 + jj < p.n_k) {
                uk = *reinterpret_cast<const uint4*>(kb + (long)(j0 + jj) * p.ldk + c);
                uv = *reinterpret_cast<const uint4*>(vb + (long)(j0 + jj) * p.ldv + c);
            }
            float* dk = &sk[jj][c]; float* dv = &sv[jj][c];
            dk[0] = bf16lo(uk.x); dk[1] = bf16hi(uk.x); dk[2] = bf16lo(uk.y); dk[3] = bf16hi(uk.y);
            dk[4] = bf16lo(uk.z); dk[5] = bf16hi(uk.z); dk[6] = bf16lo(uk.w); dk[7] = bf16hi(uk.w);
            dv[0] = bf16lo(uv.x); dv[1] = bf16hi(uv.x); dv[2] = bf16lo(uv.y); dv[3] = bf16hi(uv.y);
            dv[4] = bf16lo(uv.z); dv[5] = bf16hi(uv.z); dv[6] = bf16lo(uv.w); dv[7] = bf16hi(uv.w);
        }
        __syncthreads();
        if (!active) continue;
        const int nk = min(SA_K, p.n_k - j0);
        for (int jj = 0; jj < nk; ++jj) {
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < 64; ++c) s = fmaf(q[c], sk[jj][c], s);
            if (dump) dump[j0 + jj] = s;
            if (mrow) s += mrow[j0 + jj];
            if (s == -INFINITY) continue;
            const float mn = fmaxf(m, s);
            const float alpha = __expf(m - mn), pj = __expf(s - mn);     // m == -inf -> alpha = 0
            l = l * alpha + pj;
#pragma unroll
            for (int c = 0; c < 64; ++c) o[c] = fmaf(o[c], alpha, pj * sv[jj][c]);
            m = mn;
        }
    }
    if (!active) return;
    const float inv = l > 0.f ? 1.f / l : 0.f;
    bf16* op = p.O + b * p.o_batch_stride + h * p.o_head_stride + (long)i * p.ldo;
#pragma unroll
    for (int c = 0; c < 64; c += 8) {
        uint4 u;
        u.x = pack_bf16(o[c] * inv, o[c + 1] * inv); u.y = pack_bf16(o[c + 2] * inv, o[c + 3] * inv);
        u.z = pack_bf16(o[c + 4] * inv, o[c + 5] * inv); u.w = pack_bf16(o[c + 6] * inv, o[c + 7] * inv);
        *reinterpret_cast<uint4*>(op + c) = u;
    }
}

// Raw (pre-softmax) QK of the heads that have a dump slot - the alignment heads of the word-timestamp pass
// (decoder.py:306-308) - as its own small kernel, so that the attention itself can run on the tensor cores (attention_tc has
// no dump path): 64 keys x 32 queries per CTA, K and Q tiles as fp32 in shared memory, 8 outputs per thread.
__global__ void __launch_bounds__(256) attention_qk_dump_kernel(const AttnParams p) {
    const int h = blockIdx.z;
    const int slot = p.dump_slot[h];
    if (slot < 0) return;
    __shared__ float sk[64][65];
    __shared__ float sq[32][64];
    const int j0 = blockIdx.x * 64, i0 = blockIdx.y * 32, tid = threadIdx.x;
    const bf16* kb = p.K + h * p.k_head_stride;
    const bf16* qb = p.Q + h * p.q_head_stride;
    for (int e = tid; e < 64 * 8; e += 256) {                     // 8 x 16-byte pieces per key row
        const int jj = e >> 3, c = (e & 7) * 8;
        uint4 u = make_uint4(0, 0, 0, 0);
        if (j0 + jj < p.n_k) u = *reinterpret_cast<const uint4*>(kb + (long)(j0 + jj) * p.ldk + c);
        float* dk = &sk[jj][c];
        dk[0] = bf16lo(u.x); dk[1] = bf16hi(u.x); dk[2] = bf16lo(u.y); dk[3] = bf16hi(u.y);
        dk[4] = bf16lo(u.z); dk[5] = bf16hi(u.z); dk[6] = bf16lo(u.w); dk[7] = bf16hi(u.w);
    }
    {
        const int ii = tid >> 3, c = (tid & 7) * 8;                // 32 query rows x 8 pieces = 256 threads
        uint4 u = make_uint4(0, 0, 0, 0);
        if (i0 + ii < p.n_q) u = *reinterpret_cast<const uint4*>(qb + (long)(i0 + ii) * p.ldq + c);
        float* dq = &sq[ii][c];
        dq[0] = bf16lo(u.x); dq[1] = bf16hi(u.x); dq[2] = bf16lo(u.y); dq[3] = bf16hi(u.y);
        dq[4] = bf16lo(u.z); dq[5] = bf16hi(u.z); dq[6] = bf16lo(u.w); dq[7] = bf16hi(u.w);
    }
    __syncthreads();
    const int j = tid & 63, qg = tid >> 6;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
    for (int c = 0; c < 64; ++c) {
        const float kv = sk[j][c];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = fmaf(sq[qg * 8 + q][c], kv, acc[q]);
    }
    if (j0 + j < p.n_k) {
        float* dump = p.qk_dump + (long)slot * p.dump_slot_stride + j0 + j;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int i = i0 + qg * 8 + q;
            if (i < p.n_q) dump[(long)i * p.dump_ld] = acc[q];
        }
    }
}

void attention_qk_dump(const AttnParams& p, cudaStream_t s) {
    dim3 grid(cdiv(p.n_k, 64), cdiv(p.n_q, 32), p.n_head);
    attention_qk_dump_kernel<<<grid, 256, 0, s>>>(p);
    B200_LAUNCH_CHECK();
}

void attention_simt(const AttnParams& p, cudaStream_t s) {
    dim3 grid(cdiv(p.n_q, SA_Q), p.n_head, p.batch);
    attention_simt_kernel<<<grid, SA_Q, 0, s>>>(p);
    B200_LAUNCH_CHECK();
}

}  // namespace b200

