// tcgen05 flash attention for the audio encoder (whisper/encoder.py:36-59): per head
// softmax(Q K^T) V over 1500 frames, head dim 64, no mask.  Q/K/V are read in place from the fused
// [token][3d] QKV activation through 4-D TMA maps (dim, token, head, window).
//
//   CTA  = 128 queries of one (head, window); 2 CTAs per SM so one CTA's softmax overlaps the other's MMAs
//   warp 0 : TMA producer   Q once, then K_j / V_j (64 keys) through a 5-stage ring
//   warp 1 : MMA issuer     S_j = Q K_j^T   -> TMEM S[j & 1]   (128 x 64 fp32)
//                           O  += P_j V_j   -> TMEM O          (128 x 64 fp32, accumulated over the key blocks)
//   warps 2-5 : softmax     one query row per thread: tcgen05.ld S -> running max / sum in the log2 domain ->
//                           P_j (bf16 pairs) into TMEM with tcgen05.st: the A operand of the second MMA comes from TMEM, so
//                           a block moves 48 KB through shared memory (Q, K, V reads + the TMA writes) instead of 80
//
// O accumulates in TMEM across the key blocks (PV_j with accumulate = 1), so the softmax warps never wait for a PV MMA: the
// chain of a block is  S ready -> tcgen05.ld -> max -> exp2 -> P -> arrive.  The running maximum baked into O and l is only
// raised when some row of the warp outgrows it by more than 2^8 (lazy rescale: P stays <= 256, exact in the end because O
// and l share the factor); then the warp waits for PV_{j-1}, multiplies its 32 O rows in TMEM (tcgen05.ld / st) and goes on.
// (An earlier version kept O in registers and folded every block's PV result: the wait -> ld -> fold chain through the MMA
// warp, not the exp2, set the block time - with all arithmetic removed it still took 60 of 64 us.)
//
// V_j is consumed straight from its [key][dim] rows as an MN-major B operand.  Per (128 x 64) block the
// tensor pipe needs 256 cycles and the 8192 exp2 need 512 MUFU cycles, so the kernel is exp-bound by design.
#include "ops.cuh"

#include <map>
#include <tuple>

namespace b200 {

constexpr int FA_BM = 128, FA_BN = 64, FA_STAGES = 5, FA_THREADS = 192;      // 16 + 5 x 16 KB = 97 KB per CTA: two CTAs fit one SM
constexpr uint32_t FA_Q_BYTES = FA_BM * 64 * 2, FA_KV_BYTES = FA_BN * 64 * 2;
constexpr uint32_t FA_SMEM = FA_Q_BYTES + FA_STAGES * 2 * FA_KV_BYTES + 256;                // 16 + 80 KB + barriers
constexpr uint32_t FA_TMEM_COLS = 256;                     // S0 [0,64) S1 [64,128) O [128,192) P0 [192,224) P1 [224,256) (bf16 pairs)
constexpr uint32_t FA_P_COL = 192;

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

__device__ __forceinline__ float fmax3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

struct FaArgs {
    int n_q, n_k, n_kb;
    unsigned long long* dbg;      // optional clock64 marks of CTA (0,0,0): [block][8] (softmax warp 2: 0-4, MMA thread: 5-7)
    bf16* O; long ldo, o_head_stride, o_batch_stride;
};

__global__ void __launch_bounds__(FA_THREADS, 2)
flash_attn_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                     const __grid_constant__ CUtensorMap mapV, const FaArgs a) {
    extern __shared__ uint8_t fa_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fa_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + FA_Q_BYTES;                                  // [stage][64 keys][64 dims]
    uint8_t* sV = sK + FA_STAGES * FA_KV_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + FA_STAGES * FA_KV_BYTES);
    uint64_t* q_full = bars;                 // 1
    uint64_t* kv_full = bars + 1;            // STAGES
    uint64_t* kv_empty = kv_full + FA_STAGES;
    uint64_t* s_full = kv_empty + FA_STAGES; // 2
    uint64_t* s_empty = s_full + 2;          // 2
    uint64_t* p_full = s_empty + 2;          // 2
    uint64_t* p_empty = p_full + 2;          // 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * FA_BM, h = blockIdx.y, b = blockIdx.z;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapQ); tma_prefetch_desc(&mapK); tma_prefetch_desc(&mapV);
        mbar_init(q_full, 1);
        for (int s = 0; s < FA_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 4); mbar_init(&p_full[s], 4); mbar_init(&p_empty[s], 1); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, FA_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (threadIdx.x == 0) {
        // ------------------------------ TMA producer ------------------------------
        mbar_expect_tx(q_full, FA_Q_BYTES);
        tma_load_4d(sQ, &mapQ, q_full, 0, q0, h, b);
        int s = 0; uint32_t ph = 0;
        for (int j = 0; j < a.n_kb; ++j) {
            while (!mbar_try_wait(&kv_empty[s], ph ^ 1)) __nanosleep(128);   // the producer shares a scheduler with a softmax warp: do not spin
            mbar_expect_tx(&kv_full[s], 2 * FA_KV_BYTES);
            tma_load_4d(sK + s * FA_KV_BYTES, &mapK, &kv_full[s], 0, j * FA_BN, h, b);
            tma_load_4d(sV + s * FA_KV_BYTES, &mapV, &kv_full[s], 0, j * FA_BN, h, b);
            if (++s == FA_STAGES) { s = 0; ph ^= 1; }
        }
    } else if (threadIdx.x == 32) {
        // ------------------------------ MMA issuer ------------------------------
        // Per block this one thread performs four mbarrier waits, eight tcgen05.mma and three commits of 60-250 cycles each
        // (tools/probe_attn.py prints the per-block timeline of CTA 0).  Splitting QK and PV between two issuing threads was
        // measured slower: as a seventh warp it costs the softmax warps 40 registers (spills), as a second lane of this warp
        // the two loops serialise.
        constexpr uint32_t idesc_qk = umma_idesc_bf16(FA_BM, FA_BN, 0, 0);        // A = Q (K-major), B = K_j (K-major)
        constexpr uint32_t idesc_pv = umma_idesc_bf16(FA_BM, 64, 0, 1);           // A = P (TMEM), B = V_j (MN-major)
        const uint64_t dq = umma_desc_k128(smem_u32(sQ));
        mbar_wait(q_full, 0);
        auto issue_qk = [&](int j) {
            const int s = j % FA_STAGES, sb = j & 1;
            const bool mk = a.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && j > 0;
            mbar_wait(&kv_full[s], (j / FA_STAGES) & 1);
            if (mk) a.dbg[(j - 1) * 16 + 8] = clock64();
            mbar_wait(&s_empty[sb], ((j >> 1) & 1) ^ 1);
            if (mk) a.dbg[(j - 1) * 16 + 9] = clock64();
            tc_fence_after();
            const uint64_t dk = umma_desc_k128(smem_u32(sK + s * FA_KV_BYTES));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem + sb * FA_BN, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
            if (mk) a.dbg[(j - 1) * 16 + 10] = clock64();
            umma_commit(&s_full[sb]);
        };
        issue_qk(0);
        for (int j = 0; j < a.n_kb; ++j) {
            const bool mk = a.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
            if (mk) a.dbg[j * 16 + 5] = clock64();
            if (j + 1 < a.n_kb) issue_qk(j + 1);                   // keep the tensor pipe busy while softmax_j runs
            if (mk) a.dbg[j * 16 + 6] = clock64();
            const int s = j % FA_STAGES, sb = j & 1;
            mbar_wait(&p_full[sb], (j >> 1) & 1);
            if (mk) a.dbg[j * 16 + 7] = clock64();
            tc_fence_after();
            const uint64_t dv = umma_desc_mn128(smem_u32(sV + s * FA_KV_BYTES), 8192);
#pragma unroll
            for (int k = 0; k < 4; ++k)                            // 16 keys per MMA: 8 TMEM columns of P (bf16 pairs), +2 x 1024 B in V
                umma_bf16_ts(tmem + 2 * FA_BN, tmem + FA_P_COL + sb * 32 + k * 8, dv + 128 * k, idesc_pv, (j | k) != 0);
            if (mk) a.dbg[j * 16 + 11] = clock64();
            umma_commit(&kv_empty[s]);
            umma_commit(&p_empty[sb]);
            if (mk) a.dbg[j * 16 + 12] = clock64();
        }
    } else if (warp >= 2) {
        // ------------------------------ softmax / output ------------------------------
        const int quad = warp & 3;
        const int row = quad * 32 + lane;                          // TMEM lane == query row of the tile
        const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
        const float LOG2E = 1.4426950408889634f;
        // m: the row maximum (raw scores) that O and l are currently relative to; l: running sum of 2^((s - m) log2 e)
        float m = -INFINITY, l = 0.f;
        for (int j = 0; j < a.n_kb; ++j) {
            const int sb = j & 1;
            const bool mk = a.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && warp == 2 && lane == 0;
            if (mk) a.dbg[j * 16 + 0] = clock64();
            mbar_wait(&s_full[sb], (j >> 1) & 1);
            if (mk) a.dbg[j * 16 + 1] = clock64();
            tc_fence_after();
            uint32_t r0[32], r1[32];
            tmem_ld_32x32(lane_addr + sb * FA_BN, r0);
            tmem_ld_32x32(lane_addr + sb * FA_BN + 32, r1);
            tmem_ld_wait();
            if (mk) a.dbg[j * 16 + 2] = clock64();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[sb]);
            const int nvalid = a.n_k - j * FA_BN;                  // keys past n_k were zero-filled by TMA
            if (nvalid < FA_BN) {                                  // last block only (uniform)
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (i >= nvalid) r0[i] = 0xff800000u;          // -inf
                    if (i + 32 >= nvalid) r1[i] = 0xff800000u;
                }
            }
            float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four independent chains
#pragma unroll
            for (int i = 0; i < 32; ++i) mx4[i & 3] = fmax3(mx4[i & 3], __uint_as_float(r0[i]), __uint_as_float(r1[i]));
            const float mx = fmax3(fmaxf(mx4[0], mx4[1]), mx4[2], mx4[3]);
            if (j == 0) {
                m = mx;
            } else if (__any_sync(0xffffffffu, (mx - m) * LOG2E > 8.f)) {
                // lazy rescale (warp uniform): raise every row of the warp to its exact maximum
                const float mnew = fmaxf(m, mx);
                const float alpha = ex2_approx((m - mnew) * LOG2E);
                m = mnew;
                l *= alpha;
                mbar_wait(&p_empty[(j - 1) & 1], ((j - 1) >> 1) & 1);   // PV_{j-1} (and with it PV_0 .. PV_{j-2}) has landed in O
                tc_fence_after();
                uint32_t q0r[32], q1r[32];
                tmem_ld_32x32(lane_addr + 2 * FA_BN, q0r);
                tmem_ld_32x32(lane_addr + 2 * FA_BN + 32, q1r);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    q0r[c] = __float_as_uint(__uint_as_float(q0r[c]) * alpha);
                    q1r[c] = __float_as_uint(__uint_as_float(q1r[c]) * alpha);
                }
                tmem_st_32x32(lane_addr + 2 * FA_BN, q0r);
                tmem_st_32x32(lane_addr + 2 * FA_BN + 32, q1r);
                tmem_st_wait();
                tc_fence_before();                                 // ordered before PV_j through the p_full arrive below
            }
            const float m2 = m * LOG2E;
            float sum4[4] = {0.f, 0.f, 0.f, 0.f};
            mbar_wait(&p_empty[sb], ((j >> 1) & 1) ^ 1);           // PV_{j-2} no longer reads this P buffer
            if (mk) a.dbg[j * 16 + 3] = clock64();
            uint32_t pk[32];                                       // P_j as bf16 pairs: the A operand of PV_j, read from TMEM
#pragma unroll
            for (int i = 0; i < 16; ++i) {                         // exp2 and pack pair by pair: a score register dies as its pair is packed
                const float pa = ex2_approx(fmaf(__uint_as_float(r0[2 * i]), LOG2E, -m2));
                const float pb = ex2_approx(fmaf(__uint_as_float(r0[2 * i + 1]), LOG2E, -m2));
                const float pc = ex2_approx(fmaf(__uint_as_float(r1[2 * i]), LOG2E, -m2));
                const float pd = ex2_approx(fmaf(__uint_as_float(r1[2 * i + 1]), LOG2E, -m2));
                sum4[i & 3] += (pa + pb) + (pc + pd);
                pk[i] = pack_bf16(pa, pb);
                pk[16 + i] = pack_bf16(pc, pd);
            }
            tmem_st_32x32(lane_addr + FA_P_COL + sb * 32, pk);
            tmem_st_wait();
            l += (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[sb]);
            if (mk) a.dbg[j * 16 + 4] = clock64();
        }
        {
            uint32_t r0[32], r1[32];
            mbar_wait(&p_empty[(a.n_kb - 1) & 1], ((a.n_kb - 1) >> 1) & 1);   // the last PV has landed
            tc_fence_after();
            tmem_ld_32x32(lane_addr + 2 * FA_BN, r0);
            tmem_ld_32x32(lane_addr + 2 * FA_BN + 32, r1);
            tmem_ld_wait();
            const float inv = 1.f / l;
            if (q0 + row < a.n_q) {
                bf16* op = a.O + (long)b * a.o_batch_stride + (long)h * a.o_head_stride + (long)(q0 + row) * a.ldo;
#pragma unroll
                for (int c = 0; c < 32; c += 8) {
                    uint4 u, w;
                    u.x = pack_bf16(__uint_as_float(r0[c]) * inv, __uint_as_float(r0[c + 1]) * inv);
                    u.y = pack_bf16(__uint_as_float(r0[c + 2]) * inv, __uint_as_float(r0[c + 3]) * inv);
                    u.z = pack_bf16(__uint_as_float(r0[c + 4]) * inv, __uint_as_float(r0[c + 5]) * inv);
                    u.w = pack_bf16(__uint_as_float(r0[c + 6]) * inv, __uint_as_float(r0[c + 7]) * inv);
                    w.x = pack_bf16(__uint_as_float(r1[c]) * inv, __uint_as_float(r1[c + 1]) * inv);
                    w.y = pack_bf16(__uint_as_float(r1[c + 2]) * inv, __uint_as_float(r1[c + 3]) * inv);
                    w.z = pack_bf16(__uint_as_float(r1[c + 4]) * inv, __uint_as_float(r1[c + 5]) * inv);
                    w.w = pack_bf16(__uint_as_float(r1[c + 6]) * inv, __uint_as_float(r1[c + 7]) * inv);
                    *reinterpret_cast<uint4*>(op + c) = u;
                    *reinterpret_cast<uint4*>(op + 32 + c) = w;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, FA_TMEM_COLS); }
}

// ---- host: 4-D tensor maps (dim 64, tokens, heads, windows), cached ---------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef std::tuple<const void*, long, long, long, int, int, int, int> FaKey;
static std::map<FaKey, CUtensorMap> g_fa_maps;
void attention_clear_map_cache() { g_fa_maps.clear(); }

static const CUtensorMap* fa_map(const bf16* base, long ld, long head_stride, long batch_stride, int rows, int heads, int batch,
                                 int box_rows) {
    FaKey key(base, ld, head_stride, batch_stride, rows, heads, batch, box_rows);
    auto it = g_fa_maps.find(key);
    if (it != g_fa_maps.end()) return &it->second;
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) {
            record_error("cuTensorMapEncodeTiled entry point unavailable"); return nullptr;
        }
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    cuuint64_t gdim[4] = {64, (cuuint64_t)rows, (cuuint64_t)heads, (cuuint64_t)batch};
    cuuint64_t gstr[3] = {(cuuint64_t)ld * 2, (cuuint64_t)head_stride * 2, (cuuint64_t)(batch > 1 ? batch_stride : (long)rows * ld) * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1}, est[4] = {1, 1, 1, 1};
    CUtensorMap m;
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(base), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { record_error("attention: cuTensorMapEncodeTiled failed (%d) ld %ld hs %ld bs %ld", (int)r, ld, head_stride, batch_stride); return nullptr; }
    return &g_fa_maps.emplace(key, m).first->second;
}

unsigned long long* g_fa_dbg = nullptr;      // tests: clock marks of CTA (0,0,0) (b200TestAttentionTimeline)

void attention_tc(const AttnParams& p, cudaStream_t s) {
    if (p.mask || p.qk_dump) { attention_simt(p, s); return; }     // masks / QK dumps only exist on the 256-row prefill path
    const CUtensorMap* mq = fa_map(p.Q, p.ldq, p.q_head_stride, p.q_batch_stride, p.n_q, p.n_head, p.batch, FA_BM);
    const CUtensorMap* mk = fa_map(p.K, p.ldk, p.k_head_stride, p.k_batch_stride, p.n_k, p.n_head, p.batch, FA_BN);
    const CUtensorMap* mv = fa_map(p.V, p.ldv, p.v_head_stride, p.v_batch_stride, p.n_k, p.n_head, p.batch, FA_BN);
    if (!mq || !mk || !mv) return;
    FaArgs a;
    a.n_q = p.n_q; a.n_k = p.n_k; a.n_kb = cdiv(p.n_k, FA_BN);
    a.dbg = g_fa_dbg;
    a.O = p.O; a.ldo = p.ldo; a.o_head_stride = p.o_head_stride; a.o_batch_stride = p.o_batch_stride;
    constexpr size_t smem = FA_SMEM + 1024 + 256;
    static bool attr = false;
    if (!attr) { B200_CHECK(cudaFuncSetAttribute(flash_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    dim3 grid(cdiv(p.n_q, FA_BM), p.n_head, p.batch);
    flash_attn_tc_kernel<<<grid, FA_THREADS, smem, s>>>(*mq, *mk, *mv, a);
    B200_LAUNCH_CHECK();
}

}  // namespace b200
