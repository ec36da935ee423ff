// tcgen05 flash attention for the audio encoder (whisper/encoder.py:36-59): per head
// softmax(Q K^T) V over 1500 frames, head dim 64, no mask.  Q/K/V are read in place from the fused
// [token][3d] QKV activation through 4-D TMA maps (dim, token, head, window).
//
//   CTA  = 128 queries of one (head, window); 2 CTAs per SM so one CTA's softmax overlaps the other's MMAs
//   warp 0 : TMA producer   Q once, then K_j / V_j (64 keys) through a 3-stage ring
//   warp 1 : MMA issuer     S_j = Q K_j^T   -> TMEM S[j & 1]   (128 x 64 fp32)
//                           O_j = P_j V_j   -> TMEM Oblk       (128 x 64 fp32, fresh each block)
//   warps 2-5 : softmax     one query row per thread: tcgen05.ld S -> online max/sum in the log2 domain ->
//                           P_j (bf16) into 128B-swizzled smem as the A operand of the second MMA ->
//                           o = (o + Oblk_{j-1}) * alpha_j in registers (no TMEM read-modify-write)
//
// V_j is consumed straight from its [key][dim] rows as an MN-major B operand.  Per (128 x 64) block the
// tensor pipe needs 256 cycles and the 8192 exp2 need 512 MUFU cycles, so the kernel is exp-bound by design.
#include "ops.cuh"

#include <map>
#include <tuple>

namespace b200 {

constexpr int FA_BM = 128, FA_BN = 64, FA_STAGES = 3, FA_THREADS = 192;      // 3 stages: 97 KB per CTA, two CTAs fit one SM
constexpr uint32_t FA_Q_BYTES = FA_BM * 64 * 2, FA_KV_BYTES = FA_BN * 64 * 2, FA_P_BYTES = FA_BM * FA_BN * 2;
constexpr uint32_t FA_SMEM = FA_Q_BYTES + FA_STAGES * 2 * FA_KV_BYTES + 2 * FA_P_BYTES;     // 16 + 48 + 32 KB
constexpr uint32_t FA_TMEM_COLS = 256;                     // S0 [0,64) S1 [64,128) Oblk [128,192)

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

__device__ __forceinline__ float fmax3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float ex2_approx(float x) { float d; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(x)); return d; }

struct FaArgs {
    int n_q, n_k, n_kb;
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
    uint8_t* sP = sV + FA_STAGES * FA_KV_BYTES;                     // [2][128 rows][64 keys]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * FA_P_BYTES);
    uint64_t* q_full = bars;                 // 1
    uint64_t* kv_full = bars + 1;            // STAGES
    uint64_t* kv_empty = kv_full + FA_STAGES;
    uint64_t* s_full = kv_empty + FA_STAGES; // 2
    uint64_t* s_empty = s_full + 2;          // 2
    uint64_t* p_full = s_empty + 2;          // 2
    uint64_t* p_empty = p_full + 2;          // 2
    uint64_t* o_full = p_empty + 2;          // 1
    uint64_t* o_empty = o_full + 1;          // 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * FA_BM, h = blockIdx.y, b = blockIdx.z;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapQ); tma_prefetch_desc(&mapK); tma_prefetch_desc(&mapV);
        mbar_init(q_full, 1);
        for (int s = 0; s < FA_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 4); mbar_init(&p_full[s], 4); mbar_init(&p_empty[s], 1); }
        mbar_init(o_full, 1); mbar_init(o_empty, 4);
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
            mbar_wait(&kv_empty[s], ph ^ 1);
            mbar_expect_tx(&kv_full[s], 2 * FA_KV_BYTES);
            tma_load_4d(sK + s * FA_KV_BYTES, &mapK, &kv_full[s], 0, j * FA_BN, h, b);
            tma_load_4d(sV + s * FA_KV_BYTES, &mapV, &kv_full[s], 0, j * FA_BN, h, b);
            if (++s == FA_STAGES) { s = 0; ph ^= 1; }
        }
    } else if (threadIdx.x == 32) {
        // ------------------------------ MMA issuer ------------------------------
        constexpr uint32_t idesc_qk = umma_idesc_bf16(FA_BM, FA_BN, 0, 0);        // A = Q (K-major), B = K_j (K-major)
        constexpr uint32_t idesc_pv = umma_idesc_bf16(FA_BM, 64, 0, 1);           // A = P (K-major), B = V_j (MN-major)
        const uint64_t dq = umma_desc_k128(smem_u32(sQ));
        mbar_wait(q_full, 0);
        auto issue_qk = [&](int j) {
            const int s = j % FA_STAGES, sb = j & 1;
            mbar_wait(&kv_full[s], (j / FA_STAGES) & 1);
            mbar_wait(&s_empty[sb], ((j >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint64_t dk = umma_desc_k128(smem_u32(sK + s * FA_KV_BYTES));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem + sb * FA_BN, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
            umma_commit(&s_full[sb]);
        };
        issue_qk(0);
        for (int j = 0; j < a.n_kb; ++j) {
            if (j + 1 < a.n_kb) issue_qk(j + 1);                   // keep the tensor pipe busy while softmax_j runs
            const int s = j % FA_STAGES, sb = j & 1;
            mbar_wait(&p_full[sb], (j >> 1) & 1);
            mbar_wait(o_empty, (j & 1) ^ 1);
            tc_fence_after();
            const uint64_t dp = umma_desc_k128(smem_u32(sP + sb * FA_P_BYTES));
            const uint64_t dv = umma_desc_mn128(smem_u32(sV + s * FA_KV_BYTES), 8192);
#pragma unroll
            for (int k = 0; k < 4; ++k)                            // 16 keys per MMA: +32 B in P rows, +2 x 1024 B in V
                umma_bf16(tmem + 2 * FA_BN, dp + 2 * k, dv + 128 * k, idesc_pv, k != 0);
            umma_commit(o_full);
            umma_commit(&kv_empty[s]);
            umma_commit(&p_empty[sb]);
        }
    } else if (warp >= 2) {
        // ------------------------------ softmax / output ------------------------------
        const int quad = warp & 3;
        const int row = quad * 32 + lane;                          // TMEM lane == query row of the tile
        const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
        const float LOG2E = 1.4426950408889634f;
        float o[64];
#pragma unroll
        for (int c = 0; c < 64; ++c) o[c] = 0.f;
        // m: running row maximum of the raw scores; l: running sum of 2^((s - m) log2 e).  o holds the output relative to the
        // maximum of the block BEFORE the newest folded one, so that folding a block is one FFMA per element:
        // o <- o * alpha_prev + Oblk.
        float m = -INFINITY, l = 0.f, alpha_prev = 0.f;
        uint8_t* prow = sP + row * 128;
        for (int j = 0; j < a.n_kb; ++j) {
            const int sb = j & 1;
            mbar_wait(&s_full[sb], (j >> 1) & 1);
            tc_fence_after();
            uint32_t r0[32], r1[32];
            tmem_ld_32x32(lane_addr + sb * FA_BN, r0);
            tmem_ld_32x32(lane_addr + sb * FA_BN + 32, r1);
            tmem_ld_wait();
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
            float mx = m;
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmax3(mx, __uint_as_float(r0[i]), __uint_as_float(r1[i]));
            const float alpha = ex2_approx((m - mx) * LOG2E);      // 0 on the first block (m = -inf)
            m = mx;
            const float m2 = mx * LOG2E;
            float sum = 0.f;
            mbar_wait(&p_empty[sb], ((j >> 1) & 1) ^ 1);           // PV_{j-2} no longer reads this P buffer
            uint8_t* pb = prow + sb * FA_P_BYTES;
#pragma unroll
            for (int c = 0; c < 8; ++c) {                          // 8 chunks of 8 keys (16 B), 128B swizzle: chunk ^= row & 7
                float p[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int i = c * 8 + e;
                    p[e] = ex2_approx(fmaf(__uint_as_float(i < 32 ? r0[i] : r1[i - 32]), LOG2E, -m2));
                    sum += p[e];
                }
                uint4 u;
                u.x = pack_bf16(p[0], p[1]); u.y = pack_bf16(p[2], p[3]); u.z = pack_bf16(p[4], p[5]); u.w = pack_bf16(p[6], p[7]);
                *reinterpret_cast<uint4*>(pb + ((c ^ (row & 7)) << 4)) = u;
            }
            l = l * alpha + sum;
            fence_proxy_async();                                   // generic-proxy smem writes -> visible to the UMMA (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[sb]);
            if (j > 0) {                                           // fold in Oblk_{j-1}
                mbar_wait(o_full, (j - 1) & 1);
                tc_fence_after();
                tmem_ld_32x32(lane_addr + 2 * FA_BN, r0);
                tmem_ld_32x32(lane_addr + 2 * FA_BN + 32, r1);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(o_empty);
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    o[c] = fmaf(o[c], alpha_prev, __uint_as_float(r0[c]));
                    o[c + 32] = fmaf(o[c + 32], alpha_prev, __uint_as_float(r1[c]));
                }
            }
            alpha_prev = alpha;
        }
        {
            uint32_t r0[32], r1[32];
            mbar_wait(o_full, (a.n_kb - 1) & 1);
            tc_fence_after();
            tmem_ld_32x32(lane_addr + 2 * FA_BN, r0);
            tmem_ld_32x32(lane_addr + 2 * FA_BN + 32, r1);
            tmem_ld_wait();
            const float inv = 1.f / l;
            if (q0 + row < a.n_q) {
                bf16* op = a.O + (long)b * a.o_batch_stride + (long)h * a.o_head_stride + (long)(q0 + row) * a.ldo;
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    o[c] = fmaf(o[c], alpha_prev, __uint_as_float(r0[c])) * inv;
                    o[c + 32] = fmaf(o[c + 32], alpha_prev, __uint_as_float(r1[c])) * inv;
                }
#pragma unroll
                for (int c = 0; c < 64; c += 8) {
                    uint4 u;
                    u.x = pack_bf16(o[c], o[c + 1]); u.y = pack_bf16(o[c + 2], o[c + 3]);
                    u.z = pack_bf16(o[c + 4], o[c + 5]); u.w = pack_bf16(o[c + 6], o[c + 7]);
                    *reinterpret_cast<uint4*>(op + c) = u;
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

void attention_tc(const AttnParams& p, cudaStream_t s) {
    if (p.mask || p.qk_dump) { attention_simt(p, s); return; }     // masks / QK dumps only exist on the 256-row prefill path
    const CUtensorMap* mq = fa_map(p.Q, p.ldq, p.q_head_stride, p.q_batch_stride, p.n_q, p.n_head, p.batch, FA_BM);
    const CUtensorMap* mk = fa_map(p.K, p.ldk, p.k_head_stride, p.k_batch_stride, p.n_k, p.n_head, p.batch, FA_BN);
    const CUtensorMap* mv = fa_map(p.V, p.ldv, p.v_head_stride, p.v_batch_stride, p.n_k, p.n_head, p.batch, FA_BN);
    if (!mq || !mk || !mv) return;
    FaArgs a;
    a.n_q = p.n_q; a.n_k = p.n_k; a.n_kb = cdiv(p.n_k, FA_BN);
    a.O = p.O; a.ldo = p.ldo; a.o_head_stride = p.o_head_stride; a.o_batch_stride = p.o_batch_stride;
    constexpr size_t smem = FA_SMEM + 1024 + 256;
    static bool attr = false;
    if (!attr) { B200_CHECK(cudaFuncSetAttribute(flash_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    dim3 grid(cdiv(p.n_q, FA_BM), p.n_head, p.batch);
    flash_attn_tc_kernel<<<grid, FA_THREADS, smem, s>>>(*mq, *mk, *mv, a);
    B200_LAUNCH_CHECK();
}

}  // namespace b200
