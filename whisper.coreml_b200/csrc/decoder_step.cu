#include "decoder_step.cuh"

namespace b200 {

__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ---------------------------------------------------------------------------------------------------
// Skinny GEMM y[b, n] = act(W[n, :] . x[b, :] + bias[n]) (+ residual), b < 8.
// One CTA per 16-row weight tile; the 8 warps split K in interleaved 32-wide chunks, so the CTA
// streams one contiguous (K/32) KB run of the fragment-major weight array.  The k-permutation inside
// a chunk is the same for A and B (dot products do not care), which lets both fragments be plain
// 128-bit loads: A from global, B from the bf16 copy of x in shared memory.
// ---------------------------------------------------------------------------------------------------
constexpr int GV_THREADS = 256;
constexpr int GV_XPAD = 32;           // bf16 elements; makes the 8 beam rows hit distinct banks

__global__ void __launch_bounds__(GV_THREADS, 3) step_gemv_kernel(const StepGemv g) {
    extern __shared__ __align__(16) uint8_t gv_smem[];
    bf16* xs = reinterpret_cast<bf16*>(gv_smem);                       // [8][K + GV_XPAD]
    const int ldx = g.K + GV_XPAD;
    float* red = reinterpret_cast<float*>(gv_smem + (size_t)8 * ldx * 2);   // [8 warps][128]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (g.d_skip && *g.d_skip) return;

    // ---- prologue: beam `warp` -> bf16 row in smem (rows >= nb are zero) ----
    {
        bf16* row = xs + (long)warp * ldx;
        if (warp >= g.nb) {
            for (int k = lane * 8; k < g.K; k += 256) *reinterpret_cast<uint4*>(row + k) = make_uint4(0, 0, 0, 0);
        } else if (g.x_bf16) {
            const bf16* src = g.x_bf16 + (long)warp * g.ld_x;
            for (int k = lane * 8; k < g.K; k += 256) *reinterpret_cast<uint4*>(row + k) = *reinterpret_cast<const uint4*>(src + k);
        } else if (g.ln_g) {
            const float4* src = reinterpret_cast<const float4*>(g.x_f32 + (long)warp * g.ld_x);
            const int nv = g.K >> 7;                                     // float4 per lane (K % 128 == 0, K <= 1536)
            float4 v[12];
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 12; ++i)
                if (i < nv) { v[i] = src[lane + 32 * i]; sum += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
            const float mean = warp_sum(sum) / g.K;
            float sq = 0.f;
#pragma unroll
            for (int i = 0; i < 12; ++i)
                if (i < nv) {
                    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
                    sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
                }
            const float rstd = rsqrtf(warp_sum(sq) / g.K + g.eps);
#pragma unroll
            for (int i = 0; i < 12; ++i)
                if (i < nv) {
                    const int c4 = lane + 32 * i;
                    const float4 ga = reinterpret_cast<const float4*>(g.ln_g)[c4], be = reinterpret_cast<const float4*>(g.ln_b)[c4];
                    uint2 pk;
                    pk.x = pack_bf16(v[i].x * rstd * ga.x + be.x, v[i].y * rstd * ga.y + be.y);
                    pk.y = pack_bf16(v[i].z * rstd * ga.z + be.z, v[i].w * rstd * ga.w + be.w);
                    *reinterpret_cast<uint2*>(row + c4 * 4) = pk;
                }
        } else {
            const float* src = g.x_f32 + (long)warp * g.ld_x;
            for (int k = lane * 4; k < g.K; k += 128) {
                const float4 v = *reinterpret_cast<const float4*>(src + k);
                uint2 pk; pk.x = pack_bf16(v.x, v.y); pk.y = pack_bf16(v.z, v.w);
                *reinterpret_cast<uint2*>(row + k) = pk;
            }
        }
    }
    __syncthreads();

    const int n_tiles = (g.N + 15) >> 4, n_kc = g.K >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    const bf16* xrow = xs + (long)gq * ldx + tq * 8;                     // B fragment source of this lane
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint4* wt = reinterpret_cast<const uint4*>(g.w_frag) + (long)tile * n_kc * 64 + lane;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 5
        for (int kc = warp; kc < n_kc; kc += 8) {
            const uint4 lo = ld_stream(wt + (long)kc * 64);             // rows g,   k = 32kc + 8t .. +7
            const uint4 hi = ld_stream(wt + (long)kc * 64 + 32);        // rows g+8
            const uint4 xb = *reinterpret_cast<const uint4*>(xrow + kc * 32);
            mma_16816(acc, lo.x, hi.x, lo.y, hi.y, xb.x, xb.y);
            mma_16816(acc, lo.z, hi.z, lo.w, hi.w, xb.z, xb.w);
        }
        // acc[0],acc[1]: (row g, beams 2t, 2t+1); acc[2],acc[3]: (row g+8, same beams)
        float* r = red + warp * 128;
        r[gq * 8 + tq * 2] = acc[0]; r[gq * 8 + tq * 2 + 1] = acc[1];
        r[(gq + 8) * 8 + tq * 2] = acc[2]; r[(gq + 8) * 8 + tq * 2 + 1] = acc[3];
        __syncthreads();
        if (threadIdx.x < 128) {
            const int b = threadIdx.x >> 4, rr = threadIdx.x & 15;
            const int n = tile * 16 + rr;
            if (b < g.nb && n < g.N) {
                float v = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) v += red[w * 128 + rr * 8 + b];
                if (g.bias) v += g.bias[n];
                if (g.gelu) v = gelu_erf(v);
                if (g.residual) v += g.residual[(long)b * g.ld_res + n];
                if (g.out_f32) g.out_f32[(long)b * g.ld_out + n] = v;
                if (g.out_bf16) g.out_bf16[(long)b * g.ld_out + n] = __float2bfloat16(v);
            }
        }
        __syncthreads();
    }
}

static int g_sms = 0;
static int num_sms() {
    if (!g_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev); }
    return g_sms;
}

void step_gemv(const StepGemv& g, cudaStream_t s) {
    if (g.K % 32 != 0 || g.nb > STEP_MAX_BEAMS || (g.ln_g && (g.K % 128 != 0 || g.K > 1536))) {
        record_error("step_gemv: unsupported shape N=%d K=%d nb=%d", g.N, g.K, g.nb); return;
    }
    const size_t smem = (size_t)8 * (g.K + GV_XPAD) * 2 + 8 * 128 * 4;
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        B200_CHECK(cudaFuncSetAttribute(step_gemv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    const int tiles = (g.N + 15) / 16;
    const int grid = tiles < num_sms() * 4 ? tiles : num_sms() * 4;
    step_gemv_kernel<<<grid, GV_THREADS, smem, s>>>(g);
    B200_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------
// Self-attention of the new token against the cached positions of its beam (through the slot table).
// grid (head, beam), 128 threads.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) step_self_attn_kernel(const StepSelfAttn a) {
    __shared__ float sq[64], sknew[64], svnew[64];
    __shared__ float ss[456];
    __shared__ float sred[4][64];
    __shared__ float sstat[8];
    const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (a.d_skip && *a.d_skip) return;
    const int t = a.d_text_offset ? *a.d_text_offset : a.text_offset, d = a.d;
    const float* row = a.qkv + (long)b * 3 * d + h * 64;
    if (tid < 64) {
        sq[tid] = row[tid];
        const bf16 kb = __float2bfloat16(row[d + tid]), vb = __float2bfloat16(row[2 * d + tid]);
        sknew[tid] = __bfloat162float(kb); svnew[tid] = __bfloat162float(vb);
        const long off = ((long)b * 448 + t) * d + h * 64 + tid;          // new row lives in physical slot b
        a.cache_k[off] = kb; a.cache_v[off] = vb;
    }
    if (h == 0 && tid == 0) a.table[b * 448 + t] = b;
    __syncthreads();
    const float q0 = sq[2 * lane], q1 = sq[2 * lane + 1];
    const int* tab = a.table + b * 448;
    for (int j = warp; j < t; j += 4) {
        const int slot = tab[j];
        const uint32_t kk = *reinterpret_cast<const uint32_t*>(a.cache_k + ((long)slot * 448 + j) * d + h * 64 + 2 * lane);
        float s = warp_sum(q0 * bf16lo(kk) + q1 * bf16hi(kk));
        if (lane == 0) ss[j] = s + (a.mask ? a.mask[j] : 0.f);
    }
    if (warp == 0) {
        float s = warp_sum(q0 * sknew[2 * lane] + q1 * sknew[2 * lane + 1]);
        if (lane == 0) ss[t] = s + (a.mask ? a.mask[448] : 0.f);
    }
    __syncthreads();
    float m = -INFINITY;
    for (int j = tid; j <= t; j += 128) m = fmaxf(m, ss[j]);
    m = warp_max(m);
    if (lane == 0) sstat[warp] = m;
    __syncthreads();
    m = fmaxf(fmaxf(sstat[0], sstat[1]), fmaxf(sstat[2], sstat[3]));
    float l = 0.f;
    for (int j = tid; j <= t; j += 128) { const float p = __expf(ss[j] - m); ss[j] = p; l += p; }
    l = warp_sum(l);
    if (lane == 0) sstat[4 + warp] = l;
    __syncthreads();
    l = (sstat[4] + sstat[5]) + (sstat[6] + sstat[7]);
    float o0 = 0.f, o1 = 0.f;
    for (int j = warp; j < t; j += 4) {
        const int slot = tab[j];
        const uint32_t vv = *reinterpret_cast<const uint32_t*>(a.cache_v + ((long)slot * 448 + j) * d + h * 64 + 2 * lane);
        const float p = ss[j];
        o0 = fmaf(p, bf16lo(vv), o0); o1 = fmaf(p, bf16hi(vv), o1);
    }
    if (warp == 0) { const float p = ss[t]; o0 = fmaf(p, svnew[2 * lane], o0); o1 = fmaf(p, svnew[2 * lane + 1], o1); }
    sred[warp][2 * lane] = o0; sred[warp][2 * lane + 1] = o1;
    __syncthreads();
    if (tid < 64) {
        const float o = (sred[0][tid] + sred[1][tid]) + (sred[2][tid] + sred[3][tid]);
        a.out[(long)b * d + h * 64 + tid] = __float2bfloat16(o / l);
    }
}

void step_self_attn(const StepSelfAttn& a, cudaStream_t s) {
    dim3 grid(a.n_head, a.nb);
    step_self_attn_kernel<<<grid, 128, 0, s>>>(a);
    B200_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------
// Cross-attention of all beams against the window's 1500 audio keys: the K/V of a head are read once
// for all beams.  grid (head, split); each CTA covers a slice of the keys and leaves a partial
// (max, sum, o[64]) per beam; the last CTA of a head to finish merges the partials in split order.
// ---------------------------------------------------------------------------------------------------
constexpr int CA_THREADS = 256;
constexpr int CA_SPLITS = 7;
constexpr int CA_KEYS = 216;          // keys per split: 7 * 216 >= 1500

__global__ void __launch_bounds__(CA_THREADS) step_cross_attn_kernel(const StepCrossAttn a) {
    __shared__ __align__(16) float sq[8][64];
    __shared__ float sp[8][CA_KEYS];
    __shared__ float sred[8][8][64];
    __shared__ float sm[8], sl[8];
    __shared__ int s_last;
    const int h = blockIdx.x, sp_idx = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (a.d_skip && *a.d_skip) return;
    const int j0 = sp_idx * CA_KEYS, nk = min(CA_KEYS, a.n_keys - j0);
    for (int e = tid; e < 8 * 64; e += CA_THREADS) {
        const int b = e >> 6, c = e & 63;
        sq[b][c] = b < a.nb ? a.q[(long)b * a.d + h * 64 + c] : 0.f;
    }
    __syncthreads();
    const bf16* kh = a.ck + ((long)h * a.n_keys + j0) * 64;
    const bf16* vh = a.cv + ((long)h * a.n_keys + j0) * 64;
    if (tid < nk) {
        float k[64];
        const uint4* kp = reinterpret_cast<const uint4*>(kh + (long)tid * 64);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint4 u = kp[i];
            k[8 * i] = bf16lo(u.x); k[8 * i + 1] = bf16hi(u.x); k[8 * i + 2] = bf16lo(u.y); k[8 * i + 3] = bf16hi(u.y);
            k[8 * i + 4] = bf16lo(u.z); k[8 * i + 5] = bf16hi(u.z); k[8 * i + 6] = bf16lo(u.w); k[8 * i + 7] = bf16hi(u.w);
        }
        for (int b = 0; b < a.nb; ++b) {
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < 64; c += 4) {
                const float4 qv = *reinterpret_cast<const float4*>(&sq[b][c]);
                s = fmaf(qv.x, k[c], s); s = fmaf(qv.y, k[c + 1], s); s = fmaf(qv.z, k[c + 2], s); s = fmaf(qv.w, k[c + 3], s);
            }
            sp[b][tid] = s;
        }
    }
    __syncthreads();
    if (warp < a.nb) {                                  // warp b: partial softmax of beam b
        float m = -INFINITY;
        for (int j = lane; j < nk; j += 32) m = fmaxf(m, sp[warp][j]);
        m = warp_max(m);
        float l = 0.f;
        for (int j = lane; j < nk; j += 32) { const float p = __expf(sp[warp][j] - m); sp[warp][j] = p; l += p; }
        l = warp_sum(l);
        if (lane == 0) { sm[warp] = m; sl[warp] = l; }
    }
    __syncthreads();
    float acc[8][2];
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[b][0] = acc[b][1] = 0.f;
    for (int j = warp; j < nk; j += 8) {
        const uint32_t vv = *reinterpret_cast<const uint32_t*>(vh + (long)j * 64 + 2 * lane);
        const float v0 = bf16lo(vv), v1 = bf16hi(vv);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const float p = sp[b][j];
            acc[b][0] = fmaf(p, v0, acc[b][0]); acc[b][1] = fmaf(p, v1, acc[b][1]);
        }
    }
#pragma unroll
    for (int b = 0; b < 8; ++b) { sred[warp][b][2 * lane] = acc[b][0]; sred[warp][b][2 * lane + 1] = acc[b][1]; }
    __syncthreads();
    float* part = a.part + ((long)h * CA_SPLITS + sp_idx) * 8 * 66;
    for (int e = tid; e < a.nb * 64; e += CA_THREADS) {
        const int b = e >> 6, c = e & 63;
        float o = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) o += sred[w][b][c];
        part[b * 66 + 2 + c] = o;
    }
    if (tid < a.nb) { part[tid * 66] = sm[tid]; part[tid * 66 + 1] = sl[tid]; }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&a.counters[h], 1) == (int)gridDim.y - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const float* ph = a.part + (long)h * CA_SPLITS * 8 * 66;
    for (int e = tid; e < a.nb * 64; e += CA_THREADS) {
        const int b = e >> 6, c = e & 63;
        float M = -INFINITY;
        for (int s = 0; s < (int)gridDim.y; ++s) M = fmaxf(M, __ldcg(ph + (s * 8 + b) * 66));
        float L = 0.f, O = 0.f;
        for (int s = 0; s < (int)gridDim.y; ++s) {
            const float w = __expf(__ldcg(ph + (s * 8 + b) * 66) - M);
            L = fmaf(__ldcg(ph + (s * 8 + b) * 66 + 1), w, L);
            O = fmaf(__ldcg(ph + (s * 8 + b) * 66 + 2 + c), w, O);
        }
        a.out[(long)b * a.d + h * 64 + c] = __float2bfloat16(O / L);
    }
    if (tid == 0) a.counters[h] = 0;
}

void step_cross_attn(const StepCrossAttn& a, cudaStream_t s) {
    const int splits = cdiv(a.n_keys, CA_KEYS);
    if (splits > CA_SPLITS || a.nb > 8) { record_error("step_cross_attn: n_keys=%d nb=%d unsupported", a.n_keys, a.nb); return; }
    dim3 grid(a.n_head, splits);
    step_cross_attn_kernel<<<grid, CA_THREADS, 0, s>>>(a);
    B200_LAUNCH_CHECK();
}

__global__ void step_embed_kernel(const bf16* __restrict__ tok_emb, const float* __restrict__ pos_emb,
                                  const int* __restrict__ tokens, long token_stride, int pos, const int* __restrict__ d_pos,
                                  const int* __restrict__ d_skip, int d, float* __restrict__ x) {
    if (d_skip && *d_skip) return;
    if (d_pos) pos = *d_pos;
    const int b = blockIdx.x;
    const int tok = tokens[b * token_stride + pos];
    for (int c = threadIdx.x; c < d; c += blockDim.x)
        x[(long)b * d + c] = __bfloat162float(tok_emb[(long)tok * d + c]) + pos_emb[(long)pos * d + c];
}
void step_embed(const bf16* tok_emb, const float* pos_emb, const int* tokens, long token_stride, int pos, const int* d_pos,
                const int* d_skip, int nb, int d, float* x, cudaStream_t s) {
    step_embed_kernel<<<nb, 256, 0, s>>>(tok_emb, pos_emb, tokens, token_stride, pos, d_pos, d_skip, d, x);
    B200_LAUNCH_CHECK();
}

}  // namespace b200
