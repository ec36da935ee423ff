// decoder1 as ONE persistent kernel per token step (whisper/decoder.py:241-257, 261-327; decoding.py:707-737).
//
// The step is a chain of ~35 small dependent stages (per layer: LN+QKV, self-attention, out-proj, LN+cross-q,
// cross-attention, cross-out, LN+MLP1, MLP2; then LN+vocabulary, sampling, beam update) over 360 MB of weights
// and caches that are each read exactly once.  Launching the stages as kernels costs 5-10 us of latency apiece
// (profiles/r1_launches_v1_summary.csv); here one CTA per SM stays resident and
//
//   * warp 8 (producer) walks the CTA's static byte schedule - its share of every stage's weights and of the
//     window's cross K/V, all stored fragment-major by the exporter so a share is one contiguous run - and streams
//     it with cp.async.bulk into a 160 KB ring of 8 KB slots, running as far ahead of the math as the ring allows:
//     HBM stays busy across stage boundaries and grid barriers;
//   * warps 0-7 (consumers) wait for a slot, feed it to mma.m16n8k16 (weights = A fragments straight from the
//     slot, the <= 8 beams = the N dimension, activations = B fragments from a bf16 copy in shared memory),
//     reduce across warps, apply the stage epilogue and meet the other CTAs at a grid barrier.
//
// Stages with K split over two CTAs (MLP2) leave raw partial sums that the next stage's prologue adds in a
// fixed order, so results do not depend on timing.  Every cross-CTA activation is read with ld.global.cg.
#include "decoder_mega.cuh"

#include "sampling_dev.cuh"

namespace b200 {

constexpr int MG_CONSUMERS = 256, MG_THREADS = 288;
constexpr int MG_SLOT = 8192, MG_NSLOTS = 20;
constexpr int MG_XS_PAD = 32;
constexpr int MG_SPLIT_TILES = 14;                    // 16-key tiles per cross-attention split (224 keys)
constexpr int MG_SPLIT_KEYS = MG_SPLIT_TILES * 16;

__device__ __forceinline__ void mg_mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
struct ConsumerSync { __device__ __forceinline__ void operator()() const { consumer_sync(); } };

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---- the CTA's position in the slot ring; producer and consumers advance identical copies -----------------
struct Ring {
    int slot; uint32_t phase;
    int n;
    __device__ __forceinline__ void advance() { if (++slot == n) { slot = 0; phase ^= 1; } }
};

// ---- static schedule: which (source, bytes) chunks this CTA consumes, in order ---------------------------------
// f(const bf16* src, int n_blocks) is called for every chunk (<= 8 blocks of 1 KB) of the GEMV units
// [tile][k-slice] this CTA owns in a stage.  n_kc = K / 32 blocks per tile.
template <class F>
__device__ __forceinline__ void for_gemv_units(const bf16* w, int n_tiles, int n_kc, int ks_split, int vcta, int nctas, F f) {
    if (ks_split == 1) {
        const int u0 = (int)((long)vcta * n_tiles / nctas), u1 = (int)((long)(vcta + 1) * n_tiles / nctas);
        for (int t = u0; t < u1; ++t) f(t, 0, w + (long)t * n_kc * 512, n_kc);
    } else {                                          // two K halves: even CTAs take the low half
        const int ks = vcta & 1, g = n_kc >> 1, half = nctas >> 1;
        for (int t = vcta >> 1; t < n_tiles; t += half) f(t, ks, w + ((long)t * n_kc + ks * g) * 512, g);
    }
}

struct MegaSmem {
    uint8_t* ring; bf16* xs; float* red; float* sp; float* sq; uint64_t* full; uint64_t* empty; float* stat;
    int ldx;
};

// ---- consumer: one GEMV stage -------------------------------------------------------------------------------------
// Units (16 output rows x all beams) arrive 8 blocks per slot; warp w multiplies block w of each slot.  Threads
// 0..127 own one (beam, row) output each: its additive term (bias, residual) is fetched BEFORE the MMA loop so that
// latency hides behind the weight stream.
enum { EPI_F32 = 0, EPI_GELU_BF16 = 1, EPI_PARTIAL = 2 };
struct GemvStage {
    const bf16* w; int n_tiles, n_kc, ks_split, vcta;
    int epi;
    const float* bias;          // [N] or nullptr
    const float* res_in;        // fp32 [8][ld_out] residual read with ld.cg, or nullptr
    float* out_f32; bf16* out_bf16; long ld_out;
    int n_valid;                // outputs >= n_valid are not stored
};

__device__ __noinline__ void stage_gemv(const MegaSmem sm, Ring& ring, const GemvStage g, int nb, int nctas, int warp, int lane) {
    const int gq = lane >> 2, tq = lane & 3;
    const int tid = warp * 32 + lane, ob = tid >> 4, orow = tid & 15;
    const bf16* xrow = sm.xs + (long)gq * sm.ldx + tq * 8;
    for_gemv_units(g.w, g.n_tiles, g.n_kc, g.ks_split, g.vcta, nctas, [&](int t, int ks, const bf16*, int n_blocks) {
        const int n = t * 16 + orow;
        const int kb0 = g.ks_split == 2 ? ks * (g.n_kc >> 1) : 0;       // staged activations start at the unit's K slice
        float add = 0.f;
        if (tid < 128 && ob < nb && n < g.n_valid) {
            if (g.bias) add = __ldg(g.bias + n);
            if (g.res_in) add += __ldcg(g.res_in + (long)ob * g.ld_out + n);
        }
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c0 = 0; c0 < n_blocks; c0 += 8) {
            mbar_wait(&sm.full[ring.slot], ring.phase);
            const int blk = c0 + warp;
            if (blk < n_blocks) {
                const uint4* ap = reinterpret_cast<const uint4*>(sm.ring + (size_t)ring.slot * MG_SLOT + warp * 1024) + lane;
                const uint4 lo = ap[0], hi = ap[32];
                const uint4 xb = *reinterpret_cast<const uint4*>(xrow + blk * 32);
                mg_mma(acc, lo.x, hi.x, lo.y, hi.y, xb.x, xb.y);
                mg_mma(acc, lo.z, hi.z, lo.w, hi.w, xb.z, xb.w);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[ring.slot]);
            ring.advance();
        }
        (void)kb0;
        float* r = sm.red + warp * 128;
        r[gq * 8 + tq * 2] = acc[0]; r[gq * 8 + tq * 2 + 1] = acc[1];
        r[(gq + 8) * 8 + tq * 2] = acc[2]; r[(gq + 8) * 8 + tq * 2 + 1] = acc[3];
        consumer_sync();
        if (tid < 128 && ob < nb && n < g.n_valid) {
            float v = add;
#pragma unroll
            for (int w = 0; w < 8; ++w) v += sm.red[w * 128 + orow * 8 + ob];
            if (g.epi == EPI_F32) __stcg(g.out_f32 + (long)ob * g.ld_out + n, v);
            else if (g.epi == EPI_GELU_BF16) g.out_bf16[(long)ob * g.ld_out + n] = __float2bfloat16(gelu_erf(v));
            else __stcg(g.out_f32 + ((long)ks * 8 + ob) * g.ld_out + n, v);
        }
        consumer_sync();                              // red is reused by the next unit
    });
}

// ---- prologues: build the bf16 activation rows in shared memory ------------------------------------------------
enum { PRO_EMBED = 0, PRO_COMBINE = 1, PRO_PLAIN = 2 };

// warp b normalises row b.  x = (embed | xb_in (+ bias + part0 + part1)); optionally stored to xb_out by CTA 0.
// All global loads of a row are issued back to back (one L2 round trip) before anything is stored; the LayerNorm
// weights are fetched by all 256 threads into shared memory in the same round trip.
__device__ __noinline__ void prologue_ln(const MegaSmem sm, const MegaArgs& a, const MegaModel& M, int mode, const float* __restrict__ xb_in,
                                            float* __restrict__ xb_out, const float* __restrict__ cbias, const float* __restrict__ ln_g,
                                            const float* __restrict__ ln_b, int pos, int warp, int lane, bool store) {
    const int d = M.d, nv = d >> 7, tid = warp * 32 + lane;
    float4* sgb = reinterpret_cast<float4*>(sm.red);               // [2][d / 4]: gamma, beta (red|sp|sq are contiguous, >= 13 KB)
    float4 gb[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int q = tid + i * MG_CONSUMERS;
        if (q < d / 2) gb[i] = __ldg(reinterpret_cast<const float4*>(q < d / 4 ? ln_g : ln_b) + (q < d / 4 ? q : q - d / 4));
    }
    float4 v[12];
    if (warp < a.nb) {
        if (mode == PRO_EMBED) {
            if (a.x_in) {
#pragma unroll
                for (int i = 0; i < 12; ++i) if (i < nv) v[i] = __ldcg(reinterpret_cast<const float4*>(a.x_in + (long)warp * d) + lane + 32 * i);
            } else {
                const int tok = a.tokens[warp * DEC_TOK_LD + pos];
                uint2 e[12]; float4 p[12];
#pragma unroll
                for (int i = 0; i < 12; ++i) if (i < nv) {
                    e[i] = __ldg(reinterpret_cast<const uint2*>(M.tok_emb + (long)tok * d) + lane + 32 * i);
                    p[i] = __ldg(reinterpret_cast<const float4*>(M.pos_emb + (long)pos * d) + lane + 32 * i);
                }
#pragma unroll
                for (int i = 0; i < 12; ++i) if (i < nv)
                    v[i] = make_float4(bf16lo(e[i].x) + p[i].x, bf16hi(e[i].x) + p[i].y, bf16lo(e[i].y) + p[i].z, bf16hi(e[i].y) + p[i].w);
            }
        } else if (mode == PRO_PLAIN) {
#pragma unroll
            for (int i = 0; i < 12; ++i) if (i < nv) v[i] = __ldcg(reinterpret_cast<const float4*>(xb_in + (long)warp * d) + lane + 32 * i);
        } else {                                                       // x + bias + (part0 + part1), three batches of four
#pragma unroll
            for (int h = 0; h < 3; ++h) {
                float4 x[4], bb[4], p0[4], p1[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) { const int i = h * 4 + j; if (i < nv) {
                    const int c4 = lane + 32 * i;
                    x[j] = __ldcg(reinterpret_cast<const float4*>(xb_in + (long)warp * d) + c4);
                    bb[j] = __ldg(reinterpret_cast<const float4*>(cbias) + c4);
                    p0[j] = __ldcg(reinterpret_cast<const float4*>(a.part_m2 + (long)warp * d) + c4);
                    p1[j] = __ldcg(reinterpret_cast<const float4*>(a.part_m2 + (long)(8 + warp) * d) + c4);
                } }
#pragma unroll
                for (int j = 0; j < 4; ++j) { const int i = h * 4 + j; if (i < nv) {
                    v[i].x = x[j].x + bb[j].x + (p0[j].x + p1[j].x); v[i].y = x[j].y + bb[j].y + (p0[j].y + p1[j].y);
                    v[i].z = x[j].z + bb[j].z + (p0[j].z + p1[j].z); v[i].w = x[j].w + bb[j].w + (p0[j].w + p1[j].w);
                } }
            }
        }
        if (store && xb_out) {
#pragma unroll
            for (int i = 0; i < 12; ++i) if (i < nv) __stcg(reinterpret_cast<float4*>(xb_out + (long)warp * d) + lane + 32 * i, v[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) { const int q = tid + i * MG_CONSUMERS; if (q < d / 2) sgb[q] = gb[i]; }
    float mean = 0.f, rstd = 0.f;
    if (warp < a.nb) {
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 12; ++i) if (i < nv) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        mean = warp_sum(sum) / d;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < 12; ++i) if (i < nv) {
            v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
            sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
        }
        rstd = rsqrtf(warp_sum(sq) / d + 1e-5f);
    }
    consumer_sync();                                                   // gamma / beta are in shared memory
    if (warp < a.nb) {
        bf16* row = sm.xs + (long)warp * sm.ldx;
#pragma unroll
        for (int i = 0; i < 12; ++i) if (i < nv) {
            const int c4 = lane + 32 * i;
            const float4 ga = sgb[c4], be = sgb[d / 4 + c4];
            uint2 pk;
            pk.x = pack_bf16(v[i].x * rstd * ga.x + be.x, v[i].y * rstd * ga.y + be.y);
            pk.y = pack_bf16(v[i].z * rstd * ga.z + be.z, v[i].w * rstd * ga.w + be.w);
            *reinterpret_cast<uint2*>(row + c4 * 4) = pk;
        }
    }
    consumer_sync();
}

// copy bf16 rows [nb][n] (row stride ld, starting at column k0) into xs
__device__ __noinline__ void prologue_copy(const MegaSmem sm, const bf16* src, long ld, int k0, int n, int nb, int tid) {
    const int per_row = n >> 3;
    for (int i = tid; i < nb * per_row; i += MG_CONSUMERS) {
        const int b = i / per_row, c = (i % per_row) * 8;
        *reinterpret_cast<uint4*>(sm.xs + (long)b * sm.ldx + c) = __ldcg(reinterpret_cast<const uint4*>(src + (long)b * ld + k0 + c));
    }
    consumer_sync();
}

// ---- grid barrier (consumers only; the producer is data independent) ---------------------------------------------------
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
// timeline marks of one CTA (MEGA_DBG_CTA): after each prologue / unit loop / grid barrier
#define MEGA_DBG_CTA 100
__device__ __forceinline__ void dbg_mark(unsigned long long* dbg, int tid) {
    if (dbg && blockIdx.x == MEGA_DBG_CTA % gridDim.x && tid == 0) { const unsigned long long n = dbg[0]; if (n < 2000) { dbg[1 + n] = gtimer(); dbg[0] = n + 1; } }
}
// bar[0] counts arrivals monotonically within a launch (barrier k completes at k * nctas); the last CTA to leave the
// kernel re-arms both words, so every launch starts from zero without host help (graph replays keep their arguments).
__device__ __forceinline__ void grid_sync(unsigned* bar, unsigned& k, int nctas, int tid) {
    consumer_sync();                                   // orders this CTA's stores before thread 0's release (cumulativity)
    if (tid == 0) {
        ++k;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        const unsigned target = k * (unsigned)nctas;
        unsigned spins = 0;
        while (ld_acquire(bar) < target) {
            if (++spins > (1u << 26)) { printf("b200: grid barrier %u timeout (cta %d)\n", k, blockIdx.x); __trap(); }
        }
    }
    consumer_sync();
}
__device__ __forceinline__ void grid_leave(unsigned* bar, int nctas, int tid) {
    if (tid == 0) {
        const unsigned old = atomicAdd(&bar[1], 1u);
        if (old == (unsigned)nctas - 1) { bar[0] = 0; bar[1] = 0; __threadfence(); }
    }
}

// =================================================================================================================
__global__ void __launch_bounds__(MG_THREADS, 1) decoder_mega_kernel(const MegaArgs a) {
    extern __shared__ __align__(128) uint8_t mg_raw[];
    const MegaModel& M = *a.model;
    const int d = M.d, H = M.H, nctas = gridDim.x, cta = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (a.st && a.st->done) return;                    // uniform: the decode already finished (graph replays past the end)

    MegaSmem sm;
    sm.ring = mg_raw;
    sm.ldx = a.xs_cols + MG_XS_PAD;
    sm.xs = reinterpret_cast<bf16*>(mg_raw + (size_t)MG_NSLOTS * MG_SLOT);
    sm.red = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sm.xs) + (size_t)8 * sm.ldx * 2);
    sm.sp = sm.red + 8 * 128;                          // [8][MG_SPLIT_KEYS] scores / self-attention scores [456]
    sm.sq = sm.sp + 8 * MG_SPLIT_KEYS;                 // [8][64]
    sm.stat = sm.sq + 8 * 64;                          // 64 floats
    sm.full = reinterpret_cast<uint64_t*>(sm.stat + 64);
    sm.empty = sm.full + MG_NSLOTS;

    if (tid == 0) {
        for (int s = 0; s < MG_NSLOTS; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 8); }
        fence_barrier_init();
    }
    __syncthreads();

    dbg_mark(a.dbg, tid);
    const int pos = a.st ? a.st->pos : a.text_offset;  // text_offset of this step
    const int n_kc_d = d >> 5, n_kc_4d = d >> 3;
    const int n_splits = (CROSS_KEYS_PAD / 16 + MG_SPLIT_TILES - 1) / MG_SPLIT_TILES;
    const int n_ktiles = CROSS_KEYS_PAD / 16, n_vkc = CROSS_KEYS_PAD / 32;
    const long head_elems = (long)64 * CROSS_KEYS_PAD;

    if (warp == 8) {
        // =========================================== producer ===========================================
        if (lane != 0) return;
        Ring ring{0, 0, a.n_slots};
        auto push = [&](const bf16* src, int n_blocks) {
            for (int c0 = 0; c0 < n_blocks; c0 += 8) {
                const int nb8 = min(8, n_blocks - c0);
                mbar_wait(&sm.empty[ring.slot], ring.phase ^ 1);
                mbar_expect_tx(&sm.full[ring.slot], (uint32_t)nb8 * 1024);
                bulk_g2s(sm.ring + (size_t)ring.slot * MG_SLOT, src + (long)c0 * 512, (uint32_t)nb8 * 1024, &sm.full[ring.slot]);
                ring.advance();
            }
        };
        auto gemv = [&](const bf16* w, int n_tiles, int n_kc, int ks, int rot) {
            for_gemv_units(w, n_tiles, n_kc, ks, (cta + rot) % nctas, nctas, [&](int, int, const bf16* src, int nblk) { push(src, nblk); });
        };
        for (int l = 0; l < M.Ld; ++l) {
            const MegaLayer& L = M.layers[l];
            gemv(L.qkv, 3 * d / 16, n_kc_d, 1, 0);
            gemv(L.attn_out, d / 16, n_kc_d, 1, 0);
            gemv(L.cross_q, d / 16, n_kc_d, 1, nctas / 2);
            {   // cross-attention unit (head, split): K tiles then the four V^T dim tiles of the split
                const int u = (cta + nctas / 4) % nctas;
                if (u < H * n_splits) {
                    const int h = u / n_splits, s = u % n_splits;
                    const bf16* kf = a.ckv_frag + (long)(l * 2) * H * head_elems + h * head_elems;
                    const bf16* vf = a.ckv_frag + (long)(l * 2 + 1) * H * head_elems + h * head_elems;
                    const int t0 = s * MG_SPLIT_TILES, nt = min(MG_SPLIT_TILES, n_ktiles - t0);
                    push(kf + (long)t0 * 1024, nt * 2);
                    const int kc0 = s * (MG_SPLIT_TILES / 2), nkc = min(MG_SPLIT_TILES / 2, n_vkc - kc0);
                    for (int dt = 0; dt < 4; ++dt) push(vf + ((long)dt * n_vkc + kc0) * 512, nkc);
                }
            }
            gemv(L.cross_out, d / 16, n_kc_d, 1, 0);
            gemv(L.mlp1, 4 * d / 16, n_kc_d, 1, 0);
            gemv(L.mlp2, d / 16, n_kc_4d, 2, 0);
        }
        gemv(M.tok_emb_frag, M.n_tiles_vocab, n_kc_d, 1, 0);
        return;
    }

    // =============================================== consumers ===============================================
    Ring ring{0, 0, a.n_slots};
    unsigned bar_target = 0;                            // barriers passed so far in this launch
    int xp = 0;                                         // xb[xp] holds the current residual stream

    for (int l = 0; l < M.Ld; ++l) {
        const MegaLayer& L = M.layers[l];
        // ---------------- stage 0: LN1 + fused q|k|v ----------------
        if (l == 0) {
            prologue_ln(sm, a, M, PRO_EMBED, nullptr, a.xb[0], nullptr, L.ln1_w, L.ln1_b, pos, warp, lane, cta == 0); dbg_mark(a.dbg, tid);
            xp = 0;
        } else {
            prologue_ln(sm, a, M, PRO_COMBINE, a.xb[xp], a.xb[xp ^ 1], M.layers[l - 1].mlp2_b, L.ln1_w, L.ln1_b, pos, warp, lane, cta == 0); dbg_mark(a.dbg, tid);
            xp ^= 1;
        }
        {
            GemvStage g{L.qkv, 3 * d / 16, n_kc_d, 1, cta, EPI_F32, L.qkv_b, nullptr, a.part_qkv, nullptr, 3L * d, 3 * d};
            stage_gemv(sm, ring, g, a.nb, nctas, warp, lane);
        }
        dbg_mark(a.dbg, tid); grid_sync(a.barrier, bar_target, nctas, tid); dbg_mark(a.dbg, tid);

        // ---------------- stage 1: self-attention, unit = (beam, head) ----------------
        for (int u = cta; u < a.nb * H; u += nctas) {
            const int b = u / H, h = u % H;
            consumer_sync();
            const float* row = a.part_qkv + (long)b * 3 * d + h * 64;
            float* ss = sm.sp;                          // [<= 449] scores
            float* sknew = sm.sq; float* svnew = sm.sq + 64; float* sqv = sm.sq + 128; float* sred = sm.red;     // sred: [8][64]
            if (tid < 64) {
                sqv[tid] = __ldcg(row + tid);
                const bf16 kb = __float2bfloat16(__ldcg(row + d + tid)), vb = __float2bfloat16(__ldcg(row + 2 * d + tid));
                sknew[tid] = __bfloat162float(kb); svnew[tid] = __bfloat162float(vb);
                const long off = ((long)b * 448 + pos) * d + h * 64 + tid;     // the new row lives in physical slot b
                L.cache_k[off] = kb; L.cache_v[off] = vb;
            }
            if (h == 0 && tid == 0) a.table[b * 448 + pos] = b;
            consumer_sync();
            const float q0 = sqv[2 * lane], q1 = sqv[2 * lane + 1];
            const int* tab = a.table + b * 448;
            for (int j = warp; j < pos; j += 8) {
                const int slot = tab[j];
                const uint32_t kk = __ldcg(reinterpret_cast<const uint32_t*>(L.cache_k + ((long)slot * 448 + j) * d + h * 64 + 2 * lane));
                const float s = warp_sum(q0 * bf16lo(kk) + q1 * bf16hi(kk));
                if (lane == 0) ss[j] = s + (a.mask ? a.mask[j] : 0.f);
            }
            if (warp == 0) {
                const float s = warp_sum(q0 * sknew[2 * lane] + q1 * sknew[2 * lane + 1]);
                if (lane == 0) ss[pos] = s + (a.mask ? a.mask[448] : 0.f);
            }
            consumer_sync();
            float m = -INFINITY;
            for (int j = tid; j <= pos; j += MG_CONSUMERS) m = fmaxf(m, ss[j]);
            m = warp_max(m);
            if (lane == 0) sm.stat[warp] = m;
            consumer_sync();
            m = sm.stat[0];
#pragma unroll
            for (int w = 1; w < 8; ++w) m = fmaxf(m, sm.stat[w]);
            float lsum = 0.f;
            for (int j = tid; j <= pos; j += MG_CONSUMERS) { const float p = __expf(ss[j] - m); ss[j] = p; lsum += p; }
            lsum = warp_sum(lsum);
            if (lane == 0) sm.stat[8 + warp] = lsum;
            consumer_sync();
            lsum = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) lsum += sm.stat[8 + w];
            float o0 = 0.f, o1 = 0.f;
            for (int j = warp; j < pos; j += 8) {
                const int slot = tab[j];
                const uint32_t vv = __ldcg(reinterpret_cast<const uint32_t*>(L.cache_v + ((long)slot * 448 + j) * d + h * 64 + 2 * lane));
                const float p = ss[j];
                o0 = fmaf(p, bf16lo(vv), o0); o1 = fmaf(p, bf16hi(vv), o1);
            }
            if (warp == 0) { const float p = ss[pos]; o0 = fmaf(p, svnew[2 * lane], o0); o1 = fmaf(p, svnew[2 * lane + 1], o1); }
            sred[warp * 64 + 2 * lane] = o0; sred[warp * 64 + 2 * lane + 1] = o1;
            consumer_sync();
            if (tid < 64) {
                float o = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) o += sred[w * 64 + tid];
                a.attn[(long)b * d + h * 64 + tid] = __float2bfloat16(o / lsum);
            }
        }
        dbg_mark(a.dbg, tid); grid_sync(a.barrier, bar_target, nctas, tid); dbg_mark(a.dbg, tid);

        // ---------------- stage 2: attention out-projection + residual ----------------
        prologue_copy(sm, a.attn, d, 0, d, a.nb, tid); dbg_mark(a.dbg, tid);
        {
            GemvStage g{L.attn_out, d / 16, n_kc_d, 1, cta, EPI_F32, L.attn_out_b, a.xb[xp], a.xb[xp ^ 1], nullptr, (long)d, d};
            stage_gemv(sm, ring, g, a.nb, nctas, warp, lane);
        }
        xp ^= 1;
        dbg_mark(a.dbg, tid); grid_sync(a.barrier, bar_target, nctas, tid); dbg_mark(a.dbg, tid);

        // ---------------- stage 3: LN2 + cross query ----------------
        prologue_ln(sm, a, M, PRO_PLAIN, a.xb[xp], nullptr, nullptr, L.ln2_w, L.ln2_b, pos, warp, lane, false); dbg_mark(a.dbg, tid);
        {
            GemvStage g{L.cross_q, d / 16, n_kc_d, 1, (cta + nctas / 2) % nctas, EPI_F32, L.cross_q_b, nullptr, a.part_q, nullptr, (long)d, d};
            stage_gemv(sm, ring, g, a.nb, nctas, warp, lane);
        }
        dbg_mark(a.dbg, tid); grid_sync(a.barrier, bar_target, nctas, tid); dbg_mark(a.dbg, tid);

        // ---------------- stage 4: cross-attention, unit = (head, key split); K/V of a head are read once for all beams ----------------
        {
            const int u = (cta + nctas / 4) % nctas;
            if (u < H * n_splits) {
                const int h = u / n_splits, s = u % n_splits;
                const int t0 = s * MG_SPLIT_TILES, nt = min(MG_SPLIT_TILES, n_ktiles - t0);
                const int key0 = t0 * 16, nkeys = min(nt * 16, N_AUDIO_CTX - key0);     // valid (unpadded) keys of the split
                // q of this head -> xs rows (bf16, 64 columns)
                for (int i = tid; i < a.nb * 64; i += MG_CONSUMERS) {
                    const int b = i >> 6, c = i & 63;
                    sm.xs[(long)b * sm.ldx + c] = __float2bfloat16(__ldcg(a.part_q + (long)b * d + h * 64 + c));
                }
                consumer_sync();
                // scores: key tile tt (2 blocks) -> warp (tt % 8); slots carry 4 tiles each
                const int gq = lane >> 2, tq = lane & 3;
                for (int c0 = 0; c0 < nt * 2; c0 += 8) {
                    mbar_wait(&sm.full[ring.slot], ring.phase);
                    const int tile_in_slot = warp >> 1, kc = warp & 1, tt = (c0 >> 1) + tile_in_slot;
                    float acc[4] = {0.f, 0.f, 0.f, 0.f};
                    if (tt < nt) {
                        const uint4* ap = reinterpret_cast<const uint4*>(sm.ring + (size_t)ring.slot * MG_SLOT + warp * 1024) + lane;
                        const uint4 lo = ap[0], hi = ap[32];
                        const uint4 xb = *reinterpret_cast<const uint4*>(sm.xs + (long)gq * sm.ldx + kc * 32 + tq * 8);
                        mg_mma(acc, lo.x, hi.x, lo.y, hi.y, xb.x, xb.y);
                        mg_mma(acc, lo.z, hi.z, lo.w, hi.w, xb.z, xb.w);
                    }
                    // the two warps of a tile (kc = 0 / 1) add their halves through shared memory
                    float* r = sm.red + warp * 128;
                    r[gq * 8 + tq * 2] = acc[0]; r[gq * 8 + tq * 2 + 1] = acc[1];
                    r[(gq + 8) * 8 + tq * 2] = acc[2]; r[(gq + 8) * 8 + tq * 2 + 1] = acc[3];
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.empty[ring.slot]);
                    ring.advance();
                    consumer_sync();
                    if (tid < 128) {                    // 4 tiles x 16 keys x 8 beams = 512 sums, 4 per thread
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int idx = tid * 4 + e, tl = idx >> 7, rem = idx & 127, key = rem >> 3, b = rem & 7;
                            const int kloc = ((c0 >> 1) + tl) * 16 + key;
                            if ((c0 >> 1) + tl < nt) sm.sp[b * MG_SPLIT_KEYS + kloc] = sm.red[(tl * 2) * 128 + rem] + sm.red[(tl * 2 + 1) * 128 + rem];
                        }
                    }
                    consumer_sync();
                }
                // partial softmax of beam `warp` over the split's valid keys; p (bf16) becomes the B operand of P V
                if (warp < a.nb) {
                    float m = -INFINITY;
                    for (int j = lane; j < nkeys; j += 32) m = fmaxf(m, sm.sp[warp * MG_SPLIT_KEYS + j]);
                    m = warp_max(m);
                    float lsum = 0.f;
                    for (int j = lane; j < MG_SPLIT_KEYS; j += 32) {
                        float p = 0.f;
                        if (j < nkeys) { p = __expf(sm.sp[warp * MG_SPLIT_KEYS + j] - m); lsum += p; }
                        sm.xs[(long)warp * sm.ldx + j] = __float2bfloat16(p);
                    }
                    lsum = warp_sum(lsum);
                    if (lane == 0) { sm.stat[warp] = m; sm.stat[8 + warp] = lsum; }
                }
                consumer_sync();
                // o[dim][beam] += V^T[dim][key] p[key][beam]: slot dt holds the split's 7 key blocks of dim tile dt
                const int kc0 = s * (MG_SPLIT_TILES / 2), nkc = min(MG_SPLIT_TILES / 2, n_vkc - kc0);
                float* part = a.ca_part + ((long)h * n_splits + s) * 8 * 66;
                for (int dt = 0; dt < 4; ++dt) {
                    mbar_wait(&sm.full[ring.slot], ring.phase);
                    float acc[4] = {0.f, 0.f, 0.f, 0.f};
                    if (warp < nkc) {
                        const uint4* ap = reinterpret_cast<const uint4*>(sm.ring + (size_t)ring.slot * MG_SLOT + warp * 1024) + lane;
                        const uint4 lo = ap[0], hi = ap[32];
                        const uint4 xb = *reinterpret_cast<const uint4*>(sm.xs + (long)gq * sm.ldx + warp * 32 + tq * 8);
                        mg_mma(acc, lo.x, hi.x, lo.y, hi.y, xb.x, xb.y);
                        mg_mma(acc, lo.z, hi.z, lo.w, hi.w, xb.z, xb.w);
                    }
                    float* r = sm.red + warp * 128;
                    r[gq * 8 + tq * 2] = acc[0]; r[gq * 8 + tq * 2 + 1] = acc[1];
                    r[(gq + 8) * 8 + tq * 2] = acc[2]; r[(gq + 8) * 8 + tq * 2 + 1] = acc[3];
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.empty[ring.slot]);
                    ring.advance();
                    consumer_sync();
                    if (tid < 128) {
                        const int dim = tid >> 3, b = tid & 7;
                        float o = 0.f;
#pragma unroll
                        for (int w = 0; w < 8; ++w) o += sm.red[w * 128 + tid];
                        if (b < a.nb) __stcg(part + b * 66 + 2 + dt * 16 + dim, o);
                    }
                    consumer_sync();
                }
                if (tid < a.nb) { __stcg(part + tid * 66, sm.stat[tid]); __stcg(part + tid * 66 + 1, sm.stat[8 + tid]); }
                // the last split of a head to finish merges the partials in split order
                __threadfence();
                consumer_sync();
                int* s_last = reinterpret_cast<int*>(sm.stat + 32);
                if (tid == 0) *s_last = (atomicAdd(&a.ca_counters[h], 1) == n_splits - 1);
                consumer_sync();
                if (*s_last) {
                    __threadfence();
                    const float* ph = a.ca_part + (long)h * n_splits * 8 * 66;
                    for (int e = tid; e < a.nb * 64; e += MG_CONSUMERS) {
                        const int b = e >> 6, c = e & 63;
                        float mm = -INFINITY;
                        for (int q = 0; q < n_splits; ++q) mm = fmaxf(mm, __ldcg(ph + (q * 8 + b) * 66));
                        float ll = 0.f, oo = 0.f;
                        for (int q = 0; q < n_splits; ++q) {
                            const float w = __expf(__ldcg(ph + (q * 8 + b) * 66) - mm);
                            ll = fmaf(__ldcg(ph + (q * 8 + b) * 66 + 1), w, ll);
                            oo = fmaf(__ldcg(ph + (q * 8 + b) * 66 + 2 + c), w, oo);
                        }
                        a.attn[(long)b * d + h * 64 + c] = __float2bfloat16(oo / ll);
                    }
                    if (tid == 0) a.ca_counters[h] = 0;
                }
            }
        }
        dbg_mark(a.dbg, tid); grid_sync(a.barrier, bar_target, nctas, tid); dbg_mark(a.dbg, tid);

        // ---------------- stage 5: cross-attention out-projection + residual ----------------
        prologue_copy(sm, a.attn, d, 0, d, a.nb, tid); dbg_mark(a.dbg, tid);
        {
            GemvStage g{L.cross_out, d / 16, n_kc_d, 1, cta, EPI_F32, L.cross_out_b, a.xb[xp], a.xb[xp ^ 1], nullptr, (long)d, d};
            stage_gemv(sm, ring, g, a.nb, nctas, warp, lane);
        }
        xp ^= 1;
        dbg_mark(a.dbg, tid); grid_sync(a.barrier, bar_target, nctas, tid); dbg_mark(a.dbg, tid);

        // ---------------- stage 6: LN3 + MLP up-projection + GELU ----------------
        prologue_ln(sm, a, M, PRO_PLAIN, a.xb[xp], nullptr, nullptr, L.ln3_w, L.ln3_b, pos, warp, lane, false); dbg_mark(a.dbg, tid);
        {
            GemvStage g{L.mlp1, 4 * d / 16, n_kc_d, 1, cta, EPI_GELU_BF16, L.mlp1_b, nullptr, nullptr, a.hid, 4L * d, 4 * d};
            stage_gemv(sm, ring, g, a.nb, nctas, warp, lane);
        }
        dbg_mark(a.dbg, tid); grid_sync(a.barrier, bar_target, nctas, tid); dbg_mark(a.dbg, tid);

        // ---------------- stage 7: MLP down-projection, K split in two; raw partial sums ----------------
        {
            const int ks = cta & 1, g = n_kc_4d >> 1;
            prologue_copy(sm, a.hid, 4L * d, ks * g * 32, g * 32, a.nb, tid); dbg_mark(a.dbg, tid);
            GemvStage gs{L.mlp2, d / 16, n_kc_4d, 2, cta, EPI_PARTIAL, nullptr, nullptr, a.part_m2, nullptr, (long)d, d};
            stage_gemv(sm, ring, gs, a.nb, nctas, warp, lane);
        }
        dbg_mark(a.dbg, tid); grid_sync(a.barrier, bar_target, nctas, tid); dbg_mark(a.dbg, tid);
    }

    // ---------------- final LN + tied vocabulary projection ----------------
    prologue_ln(sm, a, M, PRO_COMBINE, a.xb[xp], nullptr, M.layers[M.Ld - 1].mlp2_b, M.ln_w, M.ln_b, pos, warp, lane, false); dbg_mark(a.dbg, tid);
    {
        GemvStage g{M.tok_emb_frag, M.n_tiles_vocab, n_kc_d, 1, cta, EPI_F32, nullptr, nullptr, a.logits, nullptr, a.ld_logits, M.V};
        stage_gemv(sm, ring, g, a.nb, nctas, warp, lane);
    }
    dbg_mark(a.dbg, tid);
    if (!a.do_sampling) { grid_leave(a.barrier, nctas, tid); return; }
    dbg_mark(a.dbg, tid); grid_sync(a.barrier, bar_target, nctas, tid); dbg_mark(a.dbg, tid);

    // ---------------- sampling: logit filters + partial log-softmax / top-k per (chunk, beam) ----------------
    SampleArgs sa;
    sa.logits = a.logits; sa.ld_logits = a.ld_logits; sa.tokens = a.tokens; sa.st = a.st; sa.spec = a.spec; sa.nb = a.nb; sa.k = a.k;
    sa.part = a.sp; sa.cand_lp = a.cand_lp; sa.cand_tok = a.cand_tok;
    for (int u = cta; u < SAMPLE_CHUNKS * a.nb; u += nctas) sample_partial_body(sa, u % SAMPLE_CHUNKS, u / SAMPLE_CHUNKS, tid, ConsumerSync());
    dbg_mark(a.dbg, tid); grid_sync(a.barrier, bar_target, nctas, tid); dbg_mark(a.dbg, tid);

    // ---------------- merge + greedy / beam update (CTA 0) ----------------
    if (cta == 0) {
        BeamUpdateArgs ba;
        ba.part = a.sp; ba.timestamp_begin = a.spec.timestamp_begin; ba.update = 1; ba.cand_lp = a.cand_lp; ba.cand_tok = a.cand_tok;
        ba.nb = a.nb; ba.k = a.k; ba.tokens = a.tokens; ba.table = a.table; ba.fin_tokens = a.fin_tokens; ba.st = a.st;
        ba.eot = a.spec.eot; ba.n_text_ctx = N_TEXT_CTX;
        beam_update_body(ba, reinterpret_cast<int*>(sm.ring), tid, ConsumerSync());
    }
    grid_leave(a.barrier, nctas, tid);
}

size_t mega_smem_bytes(int xs_cols) {
    return (size_t)MG_NSLOTS * MG_SLOT + (size_t)8 * (xs_cols + MG_XS_PAD) * 2 + (8 * 128 + 8 * MG_SPLIT_KEYS + 8 * 64 + 64) * 4 +
           2 * MG_NSLOTS * 8 + 128;
}

bool mega_launch(const MegaArgs& a, int n_ctas, cudaStream_t s) {
    const size_t smem = mega_smem_bytes(a.xs_cols);
    static size_t attr = 0;
    if (smem > attr) {
        if (cudaFuncSetAttribute(decoder_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            record_error("decoder_mega: %zu bytes of shared memory unavailable", smem);
            return false;
        }
        attr = smem;
    }
    MegaArgs args = a;
    void* params[] = {(void*)&args};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)decoder_mega_kernel, dim3(n_ctas), dim3(MG_THREADS), params, smem, s);
    ++g_launch_count;
    if (e != cudaSuccess) { cudaGetLastError(); record_error("decoder_mega launch: %s", cudaGetErrorString(e)); return false; }
    return true;
}

}  // namespace b200
