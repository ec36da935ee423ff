// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   warp 0 (1 thread)  : TMA producer   - cp.async.bulk.tensor into a STAGES-deep smem ring
//   warp 1 (1 thread)  : MMA issuer     - tcgen05.mma.cta_group::1.kind::f16, 128 x BLOCK_N x 16,
//                                          accumulators in TMEM, double buffered (2 x BLOCK_N columns)
//   warps 2..9         : epilogue       - tcgen05.ld -> bias / GELU / residual -> global; two warps per TMEM
//                                          lane quadrant take alternate 32-column chunks (the epilogue, not the MMA,
//                                          bounded the tile time with four warps)
//
// The epilogue of tile i overlaps the main loop of tile i+1 through the tmem_full / tmem_empty
// mbarrier pair.  Operands are bf16, K-major, 128-byte swizzled (TMA SWIZZLE_128B <-> UMMA
// LayoutType 2), accumulation is fp32.
#include "gemm.cuh"

#include <stdlib.h>

#include <map>
#include <tuple>
#include <vector>

namespace b200 {

static constexpr int BLOCK_M = 128;
static constexpr int BLOCK_K = 64;
static constexpr int GEMM_THREADS = 320;             // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quadrant)

struct GemmKernelArgs {
    int num_a_maps, kblocks_per_map, num_kb;
    int rows_per_batch, m_tiles_per_batch, num_m_tiles, num_n_tiles;
    int N;
    const float* bias;
    int gelu;
    const float* add;
    int add_rows;
    long ld_add;
    void* C;
    int c_fp32;
    long ldc;
    int c_batch_rows, c_row0;
    int c_split;
    long c_split_stride;
    bf16* C2; long c2_batch_stride; int c2_heads, c2_keys;
    int vec_ok;                 // output / residual rows are 16-byte aligned: 128-bit epilogue accesses allowed
    int dbg_skip;               // experiments (B200_GEMM_SKIP): 1 = epilogue without global stores, 2 = no epilogue at all
    unsigned long long* dbg;    // optional clock64 marks of CTA 0 (b200TestGemmTimeline): [tile][16]
};

// Epilogue of one 128-row x BLOCK_N accumulator tile: the calling warp owns TMEM lanes [quad * 32, +32) (tmem_acc already
// points at them) and the 64-column units half, half + 2, ...; thread <-> output row t in the TMEM layout.
//
// Stores are coalesced through a 4 KB transposition buffer per warp.  In the TMEM layout a warp store instruction writes 16
// bytes of each of 32 rows; the LSU retires such an instruction piece by piece (~4 cycles per 16-byte piece, measured with
// b200TestGemmTimeline: 15 000 cycles to store one 128 x 256 bf16 tile, twice the K = 1280 main loop), and the epilogue, not
// the MMAs, paced the GEMMs.  After the transposition lane l handles 16 bytes of the rows 4 i + (l >> 3), so an instruction
// writes (and, for the fp32 residual, reads) four whole 128-byte row segments.
//   bf16 outputs: a 64-column unit at once (two tcgen05.ld in flight, 128 bytes per row);
//   fp32 outputs: the unit's two 32-column chunks one after the other (128 bytes per row each).
// Chunks that are not fully inside N, or that also feed the fragment-major second output, keep the direct path.
template <int BLOCK_N, int INFLIGHT>
__device__ __forceinline__ void epilogue_tile(const GemmKernelArgs& g, uint32_t tmem_acc, int half, int lane, int n_blk, int b, int t,
                                              bool row_ok, long c_row, const float* add_row, float4* stage, unsigned long long* mk) {
    if (g.dbg_skip == 2) return;
    constexpr int NUNIT = BLOCK_N / 128;                // 64-column units of this warp
    constexpr int NF = 1;
    const int sub = lane >> 3, ch = lane & 7;
#pragma unroll 1
    for (int k = 0; k < NUNIT; ++k) {
    const int u64 = half + 2 * k;
    const int nU = n_blk * BLOCK_N + u64 * 64;
    if (!g.c_fp32 && !g.C2 && g.vec_ok && nU + 64 <= g.N && g.dbg_skip == 0) {
        // ---- bf16, coalesced: the whole 64-column unit ----
        uint32_t ra[32], rb[32];
        tmem_ld_32x32(tmem_acc + u64 * 64, ra);
        tmem_ld_32x32(tmem_acc + u64 * 64 + 32, rb);
        const float bva = g.bias ? __ldg(g.bias + nU + lane) : 0.f, bvb = g.bias ? __ldg(g.bias + nU + 32 + lane) : 0.f;
        if (mk) mk[2 + 6 * k] = clock64();
        tmem_ld_wait();
        if (mk) mk[3 + 6 * k] = clock64();
        uint4* st16 = reinterpret_cast<uint4*>(stage);          // [32 rows][8 x 16 bytes], 16-byte index ^= row & 7
        if (g.bias) {
            // bias: lane i loaded column nU + i (+ 32); every thread needs all 64.  Broadcast through the staging buffer with
            // 128-bit shared loads (16 per unit) - as 64 shuffles per unit the bias add alone cost 10 % of the QKV GEMM
            float* sb = reinterpret_cast<float*>(stage);
            sb[lane] = bva; sb[32 + lane] = bvb;
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 ba = reinterpret_cast<const float4*>(sb)[i], bb = reinterpret_cast<const float4*>(sb)[8 + i];
                ra[4 * i] = __float_as_uint(__uint_as_float(ra[4 * i]) + ba.x); ra[4 * i + 1] = __float_as_uint(__uint_as_float(ra[4 * i + 1]) + ba.y);
                ra[4 * i + 2] = __float_as_uint(__uint_as_float(ra[4 * i + 2]) + ba.z); ra[4 * i + 3] = __float_as_uint(__uint_as_float(ra[4 * i + 3]) + ba.w);
                rb[4 * i] = __float_as_uint(__uint_as_float(rb[4 * i]) + bb.x); rb[4 * i + 1] = __float_as_uint(__uint_as_float(rb[4 * i + 1]) + bb.y);
                rb[4 * i + 2] = __float_as_uint(__uint_as_float(rb[4 * i + 2]) + bb.z); rb[4 * i + 3] = __float_as_uint(__uint_as_float(rb[4 * i + 3]) + bb.w);
            }
            __syncwarp();                               // every lane has read the bias before the buffer is overwritten
        }
        if (g.gelu == 1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                ra[i] = __float_as_uint(gelu_erf_fit(__uint_as_float(ra[i])));
                rb[i] = __float_as_uint(gelu_erf_fit(__uint_as_float(rb[i])));
            }
        } else if (g.gelu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                ra[i] = __float_as_uint(gelu_erf(__uint_as_float(ra[i])));
                rb[i] = __float_as_uint(gelu_erf(__uint_as_float(rb[i])));
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint4 w;
            w.x = pack_bf16(__uint_as_float(ra[8 * i]), __uint_as_float(ra[8 * i + 1])); w.y = pack_bf16(__uint_as_float(ra[8 * i + 2]), __uint_as_float(ra[8 * i + 3]));
            w.z = pack_bf16(__uint_as_float(ra[8 * i + 4]), __uint_as_float(ra[8 * i + 5])); w.w = pack_bf16(__uint_as_float(ra[8 * i + 6]), __uint_as_float(ra[8 * i + 7]));
            st16[lane * 8 + (i ^ (lane & 7))] = w;
            w.x = pack_bf16(__uint_as_float(rb[8 * i]), __uint_as_float(rb[8 * i + 1])); w.y = pack_bf16(__uint_as_float(rb[8 * i + 2]), __uint_as_float(rb[8 * i + 3]));
            w.z = pack_bf16(__uint_as_float(rb[8 * i + 4]), __uint_as_float(rb[8 * i + 5])); w.w = pack_bf16(__uint_as_float(rb[8 * i + 6]), __uint_as_float(rb[8 * i + 7]));
            st16[lane * 8 + ((4 + i) ^ (lane & 7))] = w;
        }
        __syncwarp();
        const long col = (g.c_split ? (long)(nU >> 6) * g.c_split_stride : (long)nU) + 8 * ch;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int rr = 4 * i + sub;
            const uint4 x = st16[rr * 8 + (ch ^ (rr & 7))];
            if (t - lane + rr < g.rows_per_batch)
                *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(g.C) + (c_row - lane + rr) * g.ldc + col) = x;
        }
        __syncwarp();                                   // the next unit overwrites the staging buffer
        if (mk) mk[4 + 6 * k] = clock64();
        continue;
    }
#pragma unroll 1
    for (int s0 = 0; s0 < 2; ++s0) {
    const int cidx = 2 * u64 + s0;                      // 32-column chunk of the tile
    uint32_t r[NF][32];
#pragma unroll
    for (int u = 0; u < NF; ++u) tmem_ld_32x32(tmem_acc + cidx * 32, r[u]);
    // the bias and residual loads of the chunk are issued BEFORE the wait for the TMEM load (their L2 latency was on the
    // epilogue's critical path four times per tile)
    float bvs[NF]; float4 qs[NF][8];
#pragma unroll
    for (int u = 0; u < NF; ++u) {
        const int n0 = n_blk * BLOCK_N + cidx * 32;
        bvs[u] = (g.bias && n0 + lane < g.N) ? __ldg(g.bias + n0 + lane) : 0.f;   // lane i holds column n0 + i; broadcast by shuffle below
        const bool tp = g.vec_ok && n0 + 32 <= g.N && !g.C2 && g.c_fp32;
        if (add_row && tp) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float* ar = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(add_row), 4 * i + sub));
                qs[u][i] = *reinterpret_cast<const float4*>(ar + n0 + 4 * ch);
            }
        } else if (add_row && row_ok && g.vec_ok && n0 + 32 <= g.N) {
#pragma unroll
            for (int i = 0; i < 8; ++i) qs[u][i] = *reinterpret_cast<const float4*>(add_row + n0 + 4 * i);
        }
    }
    if (mk) mk[2 + 6 * k + 3 * s0] = clock64();
    tmem_ld_wait();
    if (mk) mk[3 + 6 * k + 3 * s0] = clock64();
#pragma unroll
    for (int u = 0; u < NF; ++u) {
        const int n0 = n_blk * BLOCK_N + cidx * 32;
        if (n0 >= g.N) continue;                        // warp uniform
        const bool full = g.vec_ok && n0 + 32 <= g.N;
        const float bv = bvs[u];
        float4 (&q)[8] = qs[u];
        // the (kernel-uniform) epilogue kind is tested OUTSIDE the element loops: as per-element branches it kept ptxas from
        // interleaving the 32 independent elements (46 instructions per element at 0.35 IPC in the MLP1 GEMM)
        float v[32];
        if (g.gelu == 1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = gelu_erf_fit(__uint_as_float(r[u][i]) + __shfl_sync(0xffffffffu, bv, i));
        } else if (g.gelu) {                            // B200_GELU_EXACT=1: erff()
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = gelu_erf(__uint_as_float(r[u][i]) + __shfl_sync(0xffffffffu, bv, i));
        } else if (g.bias) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[u][i]) + __shfl_sync(0xffffffffu, bv, i);
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[u][i]);
        }
        if (g.dbg_skip == 1) { float acc = 0.f; for (int i = 0; i < 32; ++i) acc += v[i]; if (acc == 1.2345e-30f) reinterpret_cast<float*>(g.C)[0] = acc; continue; }
        if (full && !g.C2 && g.c_fp32) {                // coalesced path (warp uniform)
#pragma unroll
            for (int i = 0; i < 8; ++i) stage[lane * 8 + (i ^ (lane & 7))] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            __syncwarp();
            const long col = (g.c_split ? (long)(n0 >> 6) * g.c_split_stride + (n0 & 63) : (long)n0) + 4 * ch;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = 4 * i + sub;
                float4 x = stage[rr * 8 + (ch ^ (rr & 7))];
                if (add_row) { x.x += q[i].x; x.y += q[i].y; x.z += q[i].z; x.w += q[i].w; }
                if (t - lane + rr < g.rows_per_batch) {
                    const long off = (c_row - lane + rr) * g.ldc + col;
                    *reinterpret_cast<float4*>(reinterpret_cast<float*>(g.C) + off) = x;
                }
            }
            __syncwarp();                               // the next chunk overwrites the staging buffer
            continue;
        }
        if (!row_ok) continue;                          // rows past the batch: nothing to add or store (after the shuffles)
        if (add_row) {
            if (full) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { v[4 * i] += q[i].x; v[4 * i + 1] += q[i].y; v[4 * i + 2] += q[i].z; v[4 * i + 3] += q[i].w; }
            } else {
                for (int i = 0; i < 32 && n0 + i < g.N; ++i) v[i] += add_row[n0 + i];
            }
        }
        if (g.C2) {                                     // fragment-major copy of cross K / V^T (see gemm.cuh)
            const int grp = n0 >> 6, c0 = n0 & 63, j = t;
            bf16* base = g.C2 + (long)b * g.c2_batch_stride + (long)grp * 64 * g.c2_keys;
            if (((grp / g.c2_heads) & 1) == 0) {
                bf16* dst = base + ((((j >> 4) * 2 + (c0 >> 5)) * 2 + ((j & 15) >> 3)) * 256) + (j & 7) * 32;
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    uint4 w;
                    w.x = pack_bf16(v[i], v[i + 1]); w.y = pack_bf16(v[i + 2], v[i + 3]);
                    w.z = pack_bf16(v[i + 4], v[i + 5]); w.w = pack_bf16(v[i + 6], v[i + 7]);
                    *reinterpret_cast<uint4*>(dst + i) = w;
                }
            } else {
                const int n_kc = g.c2_keys >> 5;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int c = c0 + i;
                    base[((((c >> 4) * n_kc + (j >> 5)) * 2 + ((c & 15) >> 3)) * 256) + ((c & 7) * 4 + ((j & 31) >> 3)) * 8 + (j & 7)] =
                        __float2bfloat16(v[i]);
                }
            }
        }
        // split mode: 64-column groups (heads) are c_split_stride elements apart
        const long col_off = g.c_split ? (long)(n0 >> 6) * g.c_split_stride + (n0 & 63) : (long)n0;
        if (g.c_fp32) {
            float* out = reinterpret_cast<float*>(g.C) + c_row * g.ldc + col_off;
            if (full) {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<float4*>(out + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
                for (int i = 0; i < 32 && n0 + i < g.N; ++i) out[i] = v[i];
            }
        } else {
            bf16* out = reinterpret_cast<bf16*>(g.C) + c_row * g.ldc + col_off;
            if (full) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    uint4 w;
                    w.x = pack_bf16(v[i], v[i + 1]); w.y = pack_bf16(v[i + 2], v[i + 3]);
                    w.z = pack_bf16(v[i + 4], v[i + 5]); w.w = pack_bf16(v[i + 6], v[i + 7]);
                    *reinterpret_cast<uint4*>(out + i) = w;
                }
            } else {
                for (int i = 0; i < 32 && n0 + i < g.N; ++i) out[i] = __float2bfloat16(v[i]);
            }
        }
    }
    if (mk) mk[4 + 6 * k + 3 * s0] = clock64();
    }
    }
}

template <int BLOCK_N, int STAGES, int INFLIGHT>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                    const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB,
                    const GemmKernelArgs g) {
    constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;
    constexpr uint32_t B_BYTES = BLOCK_N * BLOCK_K * 2;
    constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;            // power of two: 256 or 512

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = g.num_m_tiles * g.num_n_tiles;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapA0);
        tma_prefetch_desc(&mapB);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 8); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (threadIdx.x == 0) {
        // ------------------------------ TMA producer ------------------------------
        int s = 0; uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int n_blk = tile % g.num_n_tiles, m_tile = tile / g.num_n_tiles;
            const int b = m_tile / g.m_tiles_per_batch, t0 = (m_tile % g.m_tiles_per_batch) * BLOCK_M;
            for (int kb = 0; kb < g.num_kb; ++kb) {
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                uint8_t* sa = smem + s * STAGE_BYTES;
                const int mi = kb / g.kblocks_per_map, c0 = (kb % g.kblocks_per_map) * BLOCK_K;
                const CUtensorMap* am = mi == 0 ? &mapA0 : (mi == 1 ? &mapA1 : &mapA2);
                tma_load_3d(sa, am, &full_bar[s], c0, t0, b);
                tma_load_2d(sa + A_BYTES, &mapB, &full_bar[s], kb * BLOCK_K, n_blk * BLOCK_N);
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (threadIdx.x == 32) {
        // ------------------------------ MMA issuer ------------------------------
        constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N);
        int s = 0; uint32_t ph = 0; int a = 0; uint32_t aph = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait(&tmem_empty[a], aph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + a * BLOCK_N;
            for (int kb = 0; kb < g.num_kb; ++kb) {
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
                const uint64_t da = umma_desc_k128(sa), db = umma_desc_k128(sa + A_BYTES);
#pragma unroll
                for (int k = 0; k < BLOCK_K / 16; ++k)      // +32 B per K=16 slice inside the swizzle atom
                    umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                umma_commit(&empty_bar[s]);                 // frees the smem stage when the MMAs retire
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
            umma_commit(&tmem_full[a]);                     // accumulator ready for the epilogue
            if (g.dbg && blockIdx.x == 0) g.dbg[(size_t)(tile / gridDim.x) * 16 + 15] = clock64();
            if (++a == 2) { a = 0; aph ^= 1; }
        }
    } else if (warp >= 2) {
        // ------------------------------ epilogue ------------------------------
        const int quad = warp & 3;                          // TMEM lane quadrant this warp may access
        const int half = (warp - 2) >> 2;                   // the two warps of a quadrant take alternate 32-column chunks
        float4* stage = reinterpret_cast<float4*>(smem + STAGES * STAGE_BYTES + 256) + (warp - 2) * 256;   // 4 KB transposition buffer
        int a = 0; uint32_t aph = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int n_blk = tile % g.num_n_tiles, m_tile = tile / g.num_n_tiles;
            const int b = m_tile / g.m_tiles_per_batch, t0 = (m_tile % g.m_tiles_per_batch) * BLOCK_M;
            const int t = t0 + quad * 32 + lane;
            const bool row_ok = t < g.rows_per_batch;
            const long c_row = (long)b * g.c_batch_rows + g.c_row0 + t;
            const long m_flat = (long)b * g.rows_per_batch + t;
            const float* add_row = g.add ? g.add + (m_flat % g.add_rows) * g.ld_add : nullptr;
            unsigned long long* mk = (g.dbg && blockIdx.x == 0 && warp == 2 && lane == 0) ? g.dbg + (size_t)(tile / gridDim.x) * 16 : nullptr;
            if (mk) mk[0] = clock64();
            mbar_wait(&tmem_full[a], aph);
            if (mk) mk[1] = clock64();
            tc_fence_after();
            epilogue_tile<BLOCK_N, INFLIGHT>(g, tmem_base + ((uint32_t)(quad * 32) << 16) + a * BLOCK_N, half, lane, n_blk, b, t, row_ok, c_row, add_row, stage, mk);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[a]);
            if (mk) mk[14] = clock64();
            if (++a == 2) { a = 0; aph ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ---------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256 tile.  Each CTA stages 128 rows
// of A and 128 of the tile's 256 B rows per k-block (32 KB instead of the 48 KB of the 128 x 256 single-CTA tile, so the
// ring is PAIR_STAGES deep), the leader's MMA thread issues tcgen05.mma.cta_group::2 with M = 256, and each CTA's TMEM ends
// up with its own 128 accumulator rows x 256 columns.  Per SM and per MMA the shared-memory traffic (operand reads + TMA
// writes) drops from 192 to 128 bytes per clock - the single-CTA kernel sat on that port limit at ~65 % of the tensor rate.
//
//   full[s]       leader only, 1 arrival + 64 KB of transactions: both CTAs' TMA loads count on the leader's barrier
//   empty[s]      both CTAs, signalled by the leader's multicast tcgen05.commit
//   tmem_full[a]  both CTAs, multicast commit; tmem_empty[a] leader only, 16 arrivals (8 epilogue warps of each CTA)
// ---------------------------------------------------------------------------------------------------------------------
static constexpr int PAIR_N = 256, PAIR_STAGES = 6;

template <int INFLIGHT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                         const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB,
                         const GemmKernelArgs g) {
    constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;            // 128 rows of A
    constexpr uint32_t B_BYTES = (PAIR_N / 2) * BLOCK_K * 2;       // this CTA's half of the tile's B rows
    constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr uint32_t TMEM_COLS = 2 * PAIR_N;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + PAIR_STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + PAIR_STAGES;
    uint64_t* tmem_full = empty_bar + PAIR_STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int num_tiles = g.num_m_tiles * g.num_n_tiles;           // m tiles are 256 rows here

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapA0);
        tma_prefetch_desc(&mapB);
        for (int s = 0; s < PAIR_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 16); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                            // the peer's barriers exist before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (threadIdx.x == 0) {
        // ------------------------------ TMA producer (both CTAs) ------------------------------
        int s = 0; uint32_t ph = 0;
        for (int tile = pair; tile < num_tiles; tile += num_pairs) {
            const int n_blk = tile % g.num_n_tiles, m_tile = tile / g.num_n_tiles;
            const int b = m_tile / g.m_tiles_per_batch, t0 = (m_tile % g.m_tiles_per_batch) * 256 + (int)rank * BLOCK_M;
            for (int kb = 0; kb < g.num_kb; ++kb) {
                mbar_wait(&empty_bar[s], ph ^ 1);
                if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
                const uint32_t lead_full = mapa_u32(smem_u32(&full_bar[s]), 0);
                uint8_t* sa = smem + s * STAGE_BYTES;
                const int mi = kb / g.kblocks_per_map, c0 = (kb % g.kblocks_per_map) * BLOCK_K;
                const CUtensorMap* am = mi == 0 ? &mapA0 : (mi == 1 ? &mapA1 : &mapA2);
                tma_load_3d_pair(sa, am, lead_full, c0, t0, b);
                tma_load_2d_pair(sa + A_BYTES, &mapB, lead_full, kb * BLOCK_K, n_blk * PAIR_N + (int)rank * (PAIR_N / 2));
                if (++s == PAIR_STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (threadIdx.x == 32 && rank == 0) {
        // ------------------------------ MMA issuer (leader CTA) ------------------------------
        constexpr uint32_t idesc = umma_idesc_bf16(256, PAIR_N);
        int s = 0; uint32_t ph = 0; int a = 0; uint32_t aph = 0;
        for (int tile = pair; tile < num_tiles; tile += num_pairs) {
            mbar_wait(&tmem_empty[a], aph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + a * PAIR_N;
            for (int kb = 0; kb < g.num_kb; ++kb) {
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
                const uint64_t da = umma_desc_k128(sa), db = umma_desc_k128(sa + A_BYTES);
#pragma unroll
                for (int k = 0; k < BLOCK_K / 16; ++k)
                    umma_bf16_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                umma_commit_pair(&empty_bar[s], 3);
                if (++s == PAIR_STAGES) { s = 0; ph ^= 1; }
            }
            umma_commit_pair(&tmem_full[a], 3);
            if (++a == 2) { a = 0; aph ^= 1; }
        }
    } else if (warp >= 2) {
        // ------------------------------ epilogue (both CTAs: own 128 rows x 256 columns) ------------------------------
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        float4* stage = reinterpret_cast<float4*>(smem + PAIR_STAGES * STAGE_BYTES + 256) + (warp - 2) * 256;
        int a = 0; uint32_t aph = 0;
        for (int tile = pair; tile < num_tiles; tile += num_pairs) {
            const int n_blk = tile % g.num_n_tiles, m_tile = tile / g.num_n_tiles;
            const int b = m_tile / g.m_tiles_per_batch, t0 = (m_tile % g.m_tiles_per_batch) * 256 + (int)rank * BLOCK_M;
            const int t = t0 + quad * 32 + lane;
            const bool row_ok = t < g.rows_per_batch;
            const long c_row = (long)b * g.c_batch_rows + g.c_row0 + t;
            const long m_flat = (long)b * g.rows_per_batch + t;
            const float* add_row = g.add ? g.add + (m_flat % g.add_rows) * g.ld_add : nullptr;
            mbar_wait(&tmem_full[a], aph);
            tc_fence_after();
            epilogue_tile<PAIR_N, INFLIGHT>(g, tmem_base + ((uint32_t)(quad * 32) << 16) + a * PAIR_N, half, lane, n_blk, b, t, row_ok, c_row, add_row, stage, nullptr);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[a]), 0));
            if (++a == 2) { a = 0; aph ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                            // both CTAs are done with both TMEMs and shared memories
    if (warp == 1) { tc_fence_after(); tmem_dealloc_pair(tmem_base, TMEM_COLS); }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
            record_error("cuTensorMapEncodeTiled entry point unavailable");
            return nullptr;
        }
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// bf16 tensor map with a (64 x box_rows [x 1]) box and 128-byte swizzle, zero OOB fill.
static bool make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[3] = {dims[0], dims[1], rank == 3 ? dims[2] : 1};
    cuuint64_t gstr[2] = {strides_bytes[0], rank == 3 ? strides_bytes[1] : 0};
    cuuint32_t box[3] = {BLOCK_K, box_rows, 1};
    cuuint32_t est[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, box, est,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        record_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu strides %llu %llu base %p", (int)r, rank,
                     (unsigned long long)gdim[0], (unsigned long long)gdim[1], (unsigned long long)gdim[2],
                     (unsigned long long)gstr[0], (unsigned long long)gstr[1], base);
        return false;
    }
    return true;
}

typedef std::tuple<const void*, int, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, uint32_t> MapKey;
static std::map<MapKey, CUtensorMap> g_map_cache;

static const CUtensorMap* cached_map(const void* base, int rank, const uint64_t* dims, const uint64_t* str, uint32_t box_rows) {
    MapKey key(base, rank, dims[0], dims[1], rank == 3 ? dims[2] : 1, str[0], rank == 3 ? str[1] : 0, box_rows);
    auto it = g_map_cache.find(key);
    if (it != g_map_cache.end()) return &it->second;
    CUtensorMap m;
    if (!make_map(&m, base, rank, dims, str, box_rows)) return nullptr;
    return &g_map_cache.emplace(key, m).first->second;
}

void gemm_clear_map_cache() { g_map_cache.clear(); }

static int g_num_sms = 0;
unsigned long long* g_gemm_dbg = nullptr;    // tests: clock marks of CTA 0 (b200TestGemmTimeline)

// BLOCK_N == 0 selects the CTA-pair kernel (256 x 256 tiles)
template <int BLOCK_N, int STAGES, int INFLIGHT>
static void launch(const GemmParams& p, cudaStream_t stream) {
    constexpr bool PAIR = BLOCK_N == 0;
    constexpr int TILE_M = PAIR ? 256 : BLOCK_M, TILE_N = PAIR ? PAIR_N : BLOCK_N, B_BOX = PAIR ? PAIR_N / 2 : BLOCK_N;
    const CUtensorMap* am[3] = {nullptr, nullptr, nullptr};
    for (int i = 0; i < p.num_a_maps; ++i) {
        uint64_t dims[3] = {(uint64_t)p.a_inner, (uint64_t)p.rows_per_batch, (uint64_t)p.batch};
        uint64_t str[2] = {(uint64_t)p.a_row_stride * 2, (uint64_t)p.a_batch_stride * 2};
        am[i] = cached_map(p.A[i], 3, dims, str, BLOCK_M);
        if (!am[i]) return;
    }
    for (int i = p.num_a_maps; i < 3; ++i) am[i] = am[0];
    uint64_t bdims[2] = {(uint64_t)p.K, (uint64_t)p.N};
    uint64_t bstr[1] = {(uint64_t)p.ldb * 2};
    const CUtensorMap* bm = cached_map(p.B, 2, bdims, bstr, B_BOX);
    if (!bm) return;

    GemmKernelArgs g;
    g.num_a_maps = p.num_a_maps; g.kblocks_per_map = p.kblocks_per_map; g.num_kb = p.K / BLOCK_K;
    g.rows_per_batch = p.rows_per_batch; g.m_tiles_per_batch = cdiv(p.rows_per_batch, TILE_M);
    g.num_m_tiles = g.m_tiles_per_batch * p.batch; g.num_n_tiles = cdiv(p.N, TILE_N);
    g.N = p.N; g.bias = p.bias; g.gelu = p.gelu; g.add = p.add; g.add_rows = p.add_rows > 0 ? p.add_rows : 1;
    g.ld_add = p.ld_add; g.C = p.C; g.c_fp32 = p.c_fp32; g.ldc = p.ldc; g.c_batch_rows = p.c_batch_rows; g.c_row0 = p.c_row0;
    g.c_split = p.c_split; g.c_split_stride = p.c_split_stride;
    g.C2 = p.C2; g.c2_batch_stride = p.c2_batch_stride; g.c2_heads = p.c2_heads > 0 ? p.c2_heads : 1; g.c2_keys = p.c2_keys;
    const long c_elem = p.c_fp32 ? 4 : 2;
    g.vec_ok = ((p.ldc * c_elem) % 16 == 0) && (((uintptr_t)p.C) % 16 == 0) && ((p.c_split_stride * c_elem) % 16 == 0) &&
               (!p.add || (((p.ld_add * 4) % 16 == 0) && (((uintptr_t)p.add) % 16 == 0)));

    { static int ex = -1; if (ex < 0) { const char* e = getenv("B200_GELU_EXACT"); ex = e ? atoi(e) : 0; } if (g.gelu && ex) g.gelu = 2; }
    g.dbg = g_gemm_dbg;
    { static int skip = -1; if (skip < 0) { const char* e = getenv("B200_GEMM_SKIP"); skip = e ? atoi(e) : 0; } g.dbg_skip = skip; }
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int tiles = g.num_m_tiles * g.num_n_tiles;
    static bool attr_set = false;
    if constexpr (PAIR) {
        constexpr size_t smem = PAIR_STAGES * (BLOCK_M * BLOCK_K * 2 + (PAIR_N / 2) * BLOCK_K * 2) + 1024 + 256 + 8 * 4096;
        if (!attr_set) {
            B200_CHECK(cudaFuncSetAttribute(gemm_tcgen05_pair_kernel<INFLIGHT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
        const int pairs = tiles < g_num_sms / 2 ? tiles : g_num_sms / 2;
        gemm_tcgen05_pair_kernel<INFLIGHT><<<2 * pairs, GEMM_THREADS, smem, stream>>>(*am[0], *am[1], *am[2], *bm, g);
    } else {
        constexpr size_t smem = STAGES * (BLOCK_M * BLOCK_K * 2 + TILE_N * BLOCK_K * 2) + 1024 + 256 + 8 * 4096;
        if (!attr_set) {
            B200_CHECK(cudaFuncSetAttribute(gemm_tcgen05_kernel<TILE_N, STAGES, INFLIGHT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
        const int grid = tiles < g_num_sms ? tiles : g_num_sms;
        gemm_tcgen05_kernel<TILE_N, STAGES, INFLIGHT><<<grid, GEMM_THREADS, smem, stream>>>(*am[0], *am[1], *am[2], *bm, g);
    }
    B200_LAUNCH_CHECK();
}

int g_gemm_force = -1;       // tests: 0 = automatic, 1 = 128x128, 2 = 128x256, 3 = CTA pair 256x256

void gemm_tcgen05(const GemmParams& p, cudaStream_t stream) {
    if (p.K % BLOCK_K != 0) { record_error("gemm_tcgen05: K=%d is not a multiple of 64", p.K); return; }
    if (g_gemm_force < 0) { const char* e = getenv("B200_GEMM_TILE"); g_gemm_force = e ? atoi(e) : 0; }
    // 128x256 single-CTA tiles keep the smem operand read rate under the 128 B/clk port limit; 128x128 when the problem would
    // not give every SM a tile.  The CTA-pair kernel (256x256) pays from about four windows per batch (M >= 6000: the 8-window
    // encoder runs 20.15 instead of 21.07 ms with it); on the headline's M = 3000 it measured 1-10 % slower than 128x256.
    // B200_GEMM_TILE=1/2/3 forces a configuration.
    const long tiles256 = (long)p.batch * cdiv(p.rows_per_batch, BLOCK_M) * cdiv(p.N, 256);
    const long tiles_pair = (long)p.batch * cdiv(p.rows_per_batch, 256) * cdiv(p.N, 256);
    int sel = g_gemm_force;
    if (sel == 0) sel = ((long)p.batch * p.rows_per_batch >= 6000 && tiles_pair >= 148 && p.N >= 256) ? 3 : (tiles256 >= 120 && p.N >= 256) ? 2 : 1;
    if (sel == 3) launch<0, PAIR_STAGES, 1>(p, stream);
    else if (sel == 2) launch<256, 4, 1>(p, stream);
    else launch<128, 6, 1>(p, stream);
}

GemmParams gemm_plain(const bf16* A, const bf16* B, void* C, int M, int N, int K) {
    GemmParams p{};
    p.A[0] = A; p.num_a_maps = 1; p.kblocks_per_map = K / BLOCK_K; p.a_inner = K; p.a_row_stride = K;
    p.a_batch_stride = (long)M * K; p.rows_per_batch = M; p.batch = 1;
    p.B = B; p.ldb = K; p.N = N; p.K = K;
    p.C = C; p.ldc = N; p.c_batch_rows = 0; p.c_row0 = 0; p.add_rows = 1;
    return p;
}

// ---------------------------------------------------------------------------------------------------
// SIMT checker: one thread per output element, fp32 accumulation in K order.  Tests only.
// ---------------------------------------------------------------------------------------------------
__global__ void gemm_simt_kernel(const bf16* __restrict__ A, const bf16* __restrict__ B, const float* __restrict__ bias,
                                 void* C, int M, int N, int K, int c_fp32, int gelu) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
    if (n >= N || m >= M) return;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc += __bfloat162float(A[(long)m * K + k]) * __bfloat162float(B[(long)n * K + k]);
    if (bias) acc += bias[n];
    if (gelu) acc = gelu_erf(acc);
    if (c_fp32) reinterpret_cast<float*>(C)[(long)m * N + n] = acc;
    else reinterpret_cast<bf16*>(C)[(long)m * N + n] = __float2bfloat16(acc);
}

void gemm_simt(const bf16* A, const bf16* B, const float* bias, void* C, int M, int N, int K, int c_fp32, int gelu,
               cudaStream_t stream) {
    dim3 grid(cdiv(N, 128), M);
    gemm_simt_kernel<<<grid, 128, 0, stream>>>(A, B, bias, C, M, N, K, c_fp32, gelu);
    B200_LAUNCH_CHECK();
}

}  // namespace b200
