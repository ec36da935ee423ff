// decoder1 as ONE persistent kernel per token step (whisper/decoder.py:241-257, 261-327; decoding.py:707-737).
//
// The step is a chain of ~38 small dependent stages (per layer: LN+QKV, self-attention, out-proj, LN+cross-q,
// cross-attention, merge of its key splits, cross-out, LN+MLP1, MLP2; then LN+vocabulary) over 360 MB of weights and
// caches that are each read exactly once.  One CTA per SM stays resident and
//
//   * warp 4 (producer) walks the CTA's static byte schedule - its share of every stage's weights and of the
//     window's cross K/V, all stored fragment-major by the exporter so a share is one contiguous run - and streams
//     it with cp.async.bulk into a ring of 40 KB slots (one 16-row weight tile x K <= 1280 per slot), running as far
//     ahead of the math as the ring allows: HBM stays busy across stage boundaries;
//   * warps 0-3 (consumers) wait for a slot, feed it to mma.m16n8k16 (weights = A fragments straight from the
//     slot, the <= 8 beams = the N dimension, activations = B fragments from a bf16 copy in shared memory),
//     reduce across warps and apply the stage epilogue.
//
// What shaped the code (tools/bench_barrier.cu, bench_bcast.cu, bench_mma.cu, step_timeline.py):
//   * Stages hand their activations to each other WITHOUT grid barriers.  A software barrier costs 1.5 us idle and
//     2.2-3.2 us while HBM is saturated (almost all of it the release/acquire fences), times 36 stages.  Instead every
//     cross-CTA activation is an "LL" word: a 64-bit {payload, epoch} pair written with one single-copy-atomic
//     st.relaxed.gpu.b64 and polled by its readers until the epoch matches (epoch = launch sequence number * 64 +
//     layer + 1, so a value left by an earlier layer or step never matches).  No fence is needed because the flag
//     travels inside the same word as the data.  A buffer is rewritten one layer (9 stages) later, and a CTA can only
//     be two stages ahead of the slowest CTA (its own output is needed downstream), so a reader never sees a future
//     value either.
//   * Readers first wait on one sentinel word per producer unit and only then read everything: 148 CTAs polling whole
//     buffers flood the L2 slices that hold them and slow the producers' stores by 10x.
//   * 160 threads per CTA (255 registers; 288 threads would cap the kernel at 168 and spill).  With one warp per
//     scheduler nothing hides latency, and with 226 KB of shared memory there is no L1: a dependent global access is an
//     L2 round trip and every instruction on the critical path shows.  Hence: no local memory at all, every stage body
//     inlined exactly once (the kernel walks a table of stages), loads issued in batches before any store, no integer
//     division in per-item code, one mbarrier wait per 40 KB tile instead of per 8 KB.
//
// The kernel ends with the logits; logit filters, top-k and beam update follow as two small kernels (sampling.cu) in
// the same CUDA graph.
#include "decoder_mega.cuh"

#include <stdlib.h>

#include "sampling_dev.cuh"

namespace b200 {

constexpr int MG_CONSUMERS = 128, MG_THREADS = 160, MG_CWARPS = 4;
constexpr int MG_SLOT_BLOCKS = 40, MG_SLOT = MG_SLOT_BLOCKS * 1024, MG_MAX_SLOTS = 4;
constexpr int MG_XS_PAD = 32;
constexpr int MG_SPLIT_TILES = 14;                    // 16-key tiles per cross-attention split (224 keys)
constexpr int MG_SPLIT_KEYS = MG_SPLIT_TILES * 16;
constexpr int MG_N_SPLITS = (CROSS_KEYS_PAD / 16 + MG_SPLIT_TILES - 1) / MG_SPLIT_TILES;
constexpr unsigned MG_SPIN_LIMIT = 1u << 22;

__device__ __forceinline__ void mg_mma(float (&d)[4], const uint4& lo, const uint4& hi, const uint4& xb) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(lo.x), "r"(hi.x), "r"(lo.y), "r"(hi.y), "r"(xb.x), "r"(xb.y));
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(lo.z), "r"(hi.z), "r"(lo.w), "r"(hi.w), "r"(xb.z), "r"(xb.w));
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
struct ConsumerSync { __device__ __forceinline__ void operator()() const { consumer_sync(); } };

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// A protocol bug traps (-> launch error) instead of hanging the GPU box.
__device__ __noinline__ void mg_timeout(int where) {
    printf("b200: decoder step kernel: wait %d timed out (cta %d thread %d)\n", where, blockIdx.x, threadIdx.x);
    __trap();
}
__device__ __forceinline__ void mg_wait(uint64_t* bar, uint32_t parity) {         // compact bounded mbarrier wait
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) if (++spins > (1u << 24)) mg_timeout(0);
}

// ---- LL words -----------------------------------------------------------------------------------------------------
typedef unsigned long long u64;
__device__ __forceinline__ void ll_store(uint2* p, uint32_t payload, uint32_t epoch) {
    const u64 v = ((u64)epoch << 32) | payload;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 ll_load1(const uint2* p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ll_load2(const uint2* p, u64& a, u64& b) {         // p is 16-byte aligned
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ bool ll_ok(u64 v, uint32_t epoch) { return (uint32_t)(v >> 32) == epoch; }
__device__ __forceinline__ void ll_backoff(unsigned& spins, int where) {
    if (++spins > MG_SPIN_LIMIT) mg_timeout(where);
    if (spins > 48) __nanosleep(64);                  // the first polls spin: __nanosleep's granularity is coarser than an L2 round trip
}
__device__ __forceinline__ uint32_t ll_wait1(const uint2* p, uint32_t epoch, int where) {
    u64 v;
    unsigned spins = 0;
    while (!ll_ok(v = ll_load1(p), epoch)) ll_backoff(spins, where);
    return (uint32_t)v;
}
// wait for word `stride - 1` of every `stride`-word group of the last row (all_rows: of every row) of an LL matrix
__device__ __forceinline__ void ll_wait_sentinels(const uint2* buf, uint32_t epoch, int row_words, int stride, bool all_rows, int nb, int tid,
                                                  int where) {
    const int per_row = row_words / stride, n_sent = all_rows ? nb * per_row : per_row;
    for (int k = tid; k < n_sent; k += MG_CONSUMERS) {
        const int r = all_rows ? k / per_row : nb - 1, c = (all_rows ? k - r * per_row : k) * stride + stride - 1;
        ll_wait1(buf + (long)r * row_words + c, epoch, where);
    }
    consumer_sync();
}
// Five 16-byte LL loads issued from ONE asm statement: ptxas cannot interleave their consumers between them or recycle
// their destination registers, so all five are in flight together.  (Written as C++ the same loads were serialised in
// groups of three - the kernel sits at the 255-register limit and ptxas schedules for register count, not for latency.)
struct LL5 { u64 w[5][2]; };
__device__ __forceinline__ void ll_load2x5(LL5& r, const uint2* p0, const uint2* p1, const uint2* p2, const uint2* p3, const uint2* p4) {
    asm volatile(
        "ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%10];\n\t"
        "ld.relaxed.gpu.global.v2.u64 {%2, %3}, [%11];\n\t"
        "ld.relaxed.gpu.global.v2.u64 {%4, %5}, [%12];\n\t"
        "ld.relaxed.gpu.global.v2.u64 {%6, %7}, [%13];\n\t"
        "ld.relaxed.gpu.global.v2.u64 {%8, %9}, [%14];"
        : "=l"(r.w[0][0]), "=l"(r.w[0][1]), "=l"(r.w[1][0]), "=l"(r.w[1][1]), "=l"(r.w[2][0]), "=l"(r.w[2][1]), "=l"(r.w[3][0]), "=l"(r.w[3][1]),
          "=l"(r.w[4][0]), "=l"(r.w[4][1])
        : "l"(p0), "l"(p1), "l"(p2), "l"(p3), "l"(p4) : "memory");
}
// five 8-byte LL loads from one asm statement (see ll_load2x5)
struct L5 { u64 w[5]; };
__device__ __forceinline__ void ll_load1x5(L5& r, const uint2* p0, const uint2* p1, const uint2* p2, const uint2* p3, const uint2* p4) {
    asm volatile(
        "ld.relaxed.gpu.global.u64 %0, [%5];\n\t"
        "ld.relaxed.gpu.global.u64 %1, [%6];\n\t"
        "ld.relaxed.gpu.global.u64 %2, [%7];\n\t"
        "ld.relaxed.gpu.global.u64 %3, [%8];\n\t"
        "ld.relaxed.gpu.global.u64 %4, [%9];"
        : "=l"(r.w[0]), "=l"(r.w[1]), "=l"(r.w[2]), "=l"(r.w[3]), "=l"(r.w[4])
        : "l"(p0), "l"(p1), "l"(p2), "l"(p3), "l"(p4) : "memory");
}
// Reads an LL matrix of n_rows x row_items 16-byte items (2 words each) and hands every item to f(row, c, word0, word1).
// A thread owns items c = tid + 128 i (i < 5) of every row; two rows (10 loads) are in flight per round, and nothing is
// stored before the round's loads have all returned.  Items beyond row_items / n_rows re-read item 0 and are dropped.
constexpr int MG_IPR = 5, MG_LN_ROWS = 3;                            // items per row and thread (row_items <= 640)
template <class F>
__device__ __forceinline__ void ll_read_rows(const uint2* __restrict__ src, uint32_t epoch, int n_rows, int row_items, int tid, int where, F f) {
    bool v[MG_IPR];
    long off[MG_IPR];
#pragma unroll
    for (int i = 0; i < MG_IPR; ++i) { v[i] = tid + i * MG_CONSUMERS < row_items; off[i] = v[i] ? 2L * (tid + i * MG_CONSUMERS) : 0L; }
    for (int r0 = 0; r0 < n_rows; r0 += 2) {
        const bool two = r0 + 1 < n_rows;
        const uint2* pa = src + 2L * r0 * row_items;
        const uint2* pb = two ? pa + 2L * row_items : pa;
        LL5 ra, rb;
        unsigned spins = 0;
        bool ok;
        do {
            ll_load2x5(ra, pa + off[0], pa + off[1], pa + off[2], pa + off[3], pa + off[4]);
            ll_load2x5(rb, pb + off[0], pb + off[1], pb + off[2], pb + off[3], pb + off[4]);
            ok = true;
#pragma unroll
            for (int i = 0; i < MG_IPR; ++i) ok = ok & ll_ok(ra.w[i][0], epoch) & ll_ok(ra.w[i][1], epoch) & ll_ok(rb.w[i][0], epoch) & ll_ok(rb.w[i][1], epoch);
            if (!ok) ll_backoff(spins, where);
        } while (!ok);
#pragma unroll
        for (int i = 0; i < MG_IPR; ++i) if (v[i]) f(r0, tid + i * MG_CONSUMERS, (uint32_t)ra.w[i][0], (uint32_t)ra.w[i][1]);
        if (two) {
#pragma unroll
            for (int i = 0; i < MG_IPR; ++i) if (v[i]) f(r0 + 1, tid + i * MG_CONSUMERS, (uint32_t)rb.w[i][0], (uint32_t)rb.w[i][1]);
        }
    }
}

// The same for short rows (row_items <= 384: three items per row and thread, the 64-column attention outputs): all the rows of
// up to five beams in ONE round of loads (15 x 16 bytes in flight per thread) instead of two rows per round.
struct LL3 { u64 w[3][2]; };
__device__ __forceinline__ void ll_load2x3(LL3& r, const uint2* p0, const uint2* p1, const uint2* p2) {
    asm volatile(
        "ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%6];\n\t"
        "ld.relaxed.gpu.global.v2.u64 {%2, %3}, [%7];\n\t"
        "ld.relaxed.gpu.global.v2.u64 {%4, %5}, [%8];"
        : "=l"(r.w[0][0]), "=l"(r.w[0][1]), "=l"(r.w[1][0]), "=l"(r.w[1][1]), "=l"(r.w[2][0]), "=l"(r.w[2][1])
        : "l"(p0), "l"(p1), "l"(p2) : "memory");
}
template <class F>
__device__ __forceinline__ void ll_read_rows3(const uint2* __restrict__ src, uint32_t epoch, int n_rows, int row_items, int tid, int where, F f) {
    bool v[3];
    long off[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { v[i] = tid + i * MG_CONSUMERS < row_items; off[i] = v[i] ? 2L * (tid + i * MG_CONSUMERS) : 0L; }
    for (int r0 = 0; r0 < n_rows; r0 += 5) {
        LL3 r[5];
        const uint2* p[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) p[q] = src + 2L * min(r0 + q, n_rows - 1) * row_items;
        unsigned spins = 0;
        bool ok;
        do {
#pragma unroll
            for (int q = 0; q < 5; ++q) ll_load2x3(r[q], p[q] + off[0], p[q] + off[1], p[q] + off[2]);
            ok = true;
#pragma unroll
            for (int q = 0; q < 5; ++q)
#pragma unroll
                for (int i = 0; i < 3; ++i) ok = ok & ll_ok(r[q].w[i][0], epoch) & ll_ok(r[q].w[i][1], epoch);
            if (!ok) ll_backoff(spins, where);
        } while (!ok);
#pragma unroll
        for (int q = 0; q < 5; ++q)
            if (r0 + q < n_rows) {
#pragma unroll
                for (int i = 0; i < 3; ++i) if (v[i]) f(r0 + q, tid + i * MG_CONSUMERS, (uint32_t)r[q].w[i][0], (uint32_t)r[q].w[i][1]);
            }
    }
}

// The model descriptor lives in constant memory: its pointers are read at every stage.
__constant__ MegaModel c_model;

// ---- the CTA's position in the slot ring; producer and consumers advance identical copies -----------------
struct Ring {
    int slot; uint32_t phase;
    int n;
    __device__ __forceinline__ void advance() { if (++slot == n) { slot = 0; phase ^= 1; } }
};

// ---- static schedule: the tiles [u0, u1) of a GEMV stage this CTA owns ----------------------------------------------
__device__ __forceinline__ void gemv_range(int n_tiles, int vcta, int nctas, int& u0, int& u1) {
    u0 = vcta * n_tiles / nctas; u1 = (vcta + 1) * n_tiles / nctas;
}

struct MegaSmem {
    uint8_t* ring; bf16* xs; float* red; float* sp; float* sq; uint64_t* full; uint64_t* empty; float* stat;
    int ldx;
};

__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
// stage timeline: thread 0 of every CTA stores %globaltimer into buf[cta * MEGA_DBG_LD + mark index]
struct Dbg {
    unsigned long long* buf; bool on;
    __device__ __forceinline__ void mark(int idx) { if (on && idx < MEGA_DBG_LD) buf[idx] = gtimer(); }
    __device__ __forceinline__ void cyc(int idx) { if (on && idx < MEGA_DBG_LD) buf[idx] = (unsigned long long)clock64(); }
};

// ---- consumer: one GEMV stage -------------------------------------------------------------------------------------
// A unit = 16 output rows x all beams; its K/32 weight blocks arrive <= 40 per slot and warp w multiplies blocks
// w, w + 4, ... of the slot into two independent accumulators.  Every thread owns one (beam, row) output: its additive
// term (bias, residual) is fetched BEFORE the MMA loop so that the latency hides behind the weight stream.
enum { EPI_LL_F32 = 0, EPI_LL_GELU_BF16 = 1, EPI_LOGITS = 2 };
enum { RES_NONE = 0, RES_LL = 1, RES_EMBED = 2, RES_XIN = 3 };
struct GemvStage {
    int n_tiles, n_kc, vcta;
    int epi;
    const float* bias;          // [N] or nullptr
    int res_mode; const uint2* res_ll; uint32_t res_epoch;      // residual: LL fp32 [nb][ld_out]
    uint2* out_ll; float* out_f32; long ld_out;
    uint2* out_llb;             // EPI_LL_F32: the same outputs again as bf16x2 LL words (LayerNorm input of the next stage) or nullptr
    int n_valid;                // outputs >= n_valid are not stored
    uint32_t epoch;
};

__device__ __forceinline__ void stage_gemv(const MegaSmem& sm, Ring& ring, int& red_buf, const GemvStage& g, const MegaArgs& a,
                                           int pos, int nctas, int warp, int lane) {
    const MegaModel& M = c_model;
    const int gq = lane >> 2, tq = lane & 3;
    const int tid = warp * 32 + lane, ob = tid >> 4, orow = tid & 15;
    const bf16* xrow = sm.xs + (long)gq * sm.ldx + tq * 8;
    int u0, u1;
    gemv_range(g.n_tiles, g.vcta, nctas, u0, u1);
    for (int t = u0; t < u1; ++t) {
        const int n = t * 16 + orow;
        const bool owner = ob < a.nb && n < g.n_valid;
        float add = 0.f;
        if (owner) {
            if (g.bias) add = __ldg(g.bias + n);
            if (g.res_mode == RES_LL) add += __uint_as_float(ll_wait1(g.res_ll + (long)ob * g.ld_out + n, g.res_epoch, 1));
            else if (g.res_mode == RES_EMBED)
                add += __bfloat162float(M.tok_emb[(long)a.tokens[ob * DEC_TOK_LD + pos] * M.d + n]) + __ldg(M.pos_emb + (long)pos * M.d + n);
            else if (g.res_mode == RES_XIN) add += __ldg(a.x_in + (long)ob * M.d + n);
        }
        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        for (int c0 = 0; c0 < g.n_kc; c0 += MG_SLOT_BLOCKS) {
            const int nblk = min(MG_SLOT_BLOCKS, g.n_kc - c0);
            const uint4* sl = reinterpret_cast<const uint4*>(sm.ring + (size_t)ring.slot * MG_SLOT) + lane;
            const bf16* xk = xrow + c0 * 32;
            mg_wait(&sm.full[ring.slot], ring.phase);
#pragma unroll
            for (int q0 = 0; q0 < MG_SLOT_BLOCKS / MG_CWARPS; q0 += 5) {          // two batches of five blocks per warp
                uint4 lo[5], hi[5], xb[5];
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    const int blk = warp + MG_CWARPS * (q0 + q);
                    if (blk < nblk) { lo[q] = sl[blk * 64]; hi[q] = sl[blk * 64 + 32]; xb[q] = *reinterpret_cast<const uint4*>(xk + blk * 32); }
                }
#pragma unroll
                for (int q = 0; q < 5; ++q)
                    if (warp + MG_CWARPS * (q0 + q) < nblk) mg_mma(acc[q & 1], lo[q], hi[q], xb[q]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[ring.slot]);
            ring.advance();
        }
        float* r = sm.red + red_buf * 512 + warp * 128;
        r[gq * 8 + tq * 2] = acc[0][0] + acc[1][0]; r[gq * 8 + tq * 2 + 1] = acc[0][1] + acc[1][1];
        r[(gq + 8) * 8 + tq * 2] = acc[0][2] + acc[1][2]; r[(gq + 8) * 8 + tq * 2 + 1] = acc[0][3] + acc[1][3];
        consumer_sync();
        {
            const float* rr = sm.red + red_buf * 512 + orow * 8 + ob;
            float v = add + ((rr[0] + rr[128]) + (rr[256] + rr[384]));
            if (g.epi == EPI_LL_GELU_BF16) {
                v = gelu_erf(v);
                const float nxt = __shfl_down_sync(0xffffffffu, v, 1);          // lanes are (ob, orow): orow + 1 is the next lane
                if (owner && !(orow & 1)) ll_store(g.out_ll + ((long)ob * g.ld_out + n) / 2, pack_bf16(v, nxt), g.epoch);
            } else if (g.epi == EPI_LL_F32) {
                if (owner) ll_store(g.out_ll + (long)ob * g.ld_out + n, __float_as_uint(v), g.epoch);
                if (g.out_llb) {                                      // (stage uniform)
                    const float nxt = __shfl_down_sync(0xffffffffu, v, 1);
                    if (owner && !(orow & 1)) ll_store(g.out_llb + ((long)ob * g.ld_out + n) / 2, pack_bf16(v, nxt), g.epoch);
                }
            } else if (owner) {
                __stcg(g.out_f32 + (long)ob * g.ld_out + n, v);
            }
        }
        red_buf ^= 1;                                 // the next unit reduces through the other buffer: one barrier per unit
    }
}

// ---- prologues: build the bf16 activation rows in shared memory ------------------------------------------------
enum { PRO_EMBED = 0, PRO_LL = 1 };

// five 8-byte shared-memory loads, 1 KB apart (a thread's five items of a row), from one asm statement (see ll_load2x5)
__device__ __forceinline__ void lds5(float2 (&v)[5], uint32_t saddr) {
    asm volatile(
        "ld.shared.v2.f32 {%0, %1}, [%10];\n\t"
        "ld.shared.v2.f32 {%2, %3}, [%10+1024];\n\t"
        "ld.shared.v2.f32 {%4, %5}, [%10+2048];\n\t"
        "ld.shared.v2.f32 {%6, %7}, [%10+3072];\n\t"
        "ld.shared.v2.f32 {%8, %9}, [%10+4096];"
        : "=f"(v[0].x), "=f"(v[0].y), "=f"(v[1].x), "=f"(v[1].y), "=f"(v[2].x), "=f"(v[2].y), "=f"(v[3].x), "=f"(v[3].y), "=f"(v[4].x), "=f"(v[4].y)
        : "r"(saddr) : "memory");
}
__device__ __forceinline__ float warp_sum2(float& a, float& b) {           // two interleaved butterfly sums
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    return a;
}
// LayerNorm of the residual stream.  Every thread owns items c = tid + 128 i (2 columns each) of every row: it fetches
// them - token + position embeddings (or x_in), or an LL buffer, ten loads in flight - parks them as fp32 in the unused
// tail of the xs row, and adds them into the row's sum and sum of squares (per warp, then across warps through shared
// memory).  After one barrier it normalises its own items into the bf16 row; gamma | beta arrive through the ring.
__device__ __forceinline__ void prologue_ln(const MegaSmem& sm, Ring& ring, const MegaArgs& a, int mode, const uint2* __restrict__ x_ll, uint32_t epoch,
                                            int pos, int warp, int lane, Dbg& dbg, int db) {
    const MegaModel& M = c_model;
    const int d = M.d, tid = warp * 32 + lane, row_items = d >> 1;         // 2 columns per 16-byte item
    float* part = sm.red;                                                  // [4 warps][8 rows][sum, sum of squares]
    consumer_sync();                                                       // the previous stage is done with xs / red / sp / sq
    dbg.cyc(db + 1);
    bool v[MG_IPR];
    int cc[MG_IPR];
#pragma unroll
    for (int i = 0; i < MG_IPR; ++i) { v[i] = tid + i * MG_CONSUMERS < row_items; cc[i] = v[i] ? tid + i * MG_CONSUMERS : 0; }   // out of range: re-read item 0
    auto row_done = [&](int row, float s1, float s2) {
        warp_sum2(s1, s2);
        if (lane == 0) { part[(warp * 8 + row) * 2] = s1; part[(warp * 8 + row) * 2 + 1] = s2; }
    };
    if (mode == PRO_EMBED) {
        int* stok = reinterpret_cast<int*>(sm.stat);
        if (!a.x_in) {
            if (tid < a.nb) stok[tid] = a.tokens[tid * DEC_TOK_LD + pos];
            consumer_sync();
        }
        if (a.x_in) {                                                      // reference ABI: embedded rows from the caller, one row per round
#pragma unroll 1
            for (int b = 0; b < a.nb; ++b) {
                float2 xv[MG_IPR];
                const float2* xp = reinterpret_cast<const float2*>(a.x_in + (long)b * d);
                asm volatile(
                    "ld.global.nc.v2.f32 {%0, %1}, [%10];\n\t"
                    "ld.global.nc.v2.f32 {%2, %3}, [%11];\n\t"
                    "ld.global.nc.v2.f32 {%4, %5}, [%12];\n\t"
                    "ld.global.nc.v2.f32 {%6, %7}, [%13];\n\t"
                    "ld.global.nc.v2.f32 {%8, %9}, [%14];"
                    : "=f"(xv[0].x), "=f"(xv[0].y), "=f"(xv[1].x), "=f"(xv[1].y), "=f"(xv[2].x), "=f"(xv[2].y), "=f"(xv[3].x), "=f"(xv[3].y), "=f"(xv[4].x), "=f"(xv[4].y)
                    : "l"(xp + cc[0]), "l"(xp + cc[1]), "l"(xp + cc[2]), "l"(xp + cc[3]), "l"(xp + cc[4]) : "memory");
                float2* dst = reinterpret_cast<float2*>(sm.xs + (long)b * sm.ldx + d);
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int i = 0; i < MG_IPR; ++i)
                    if (v[i]) {
                        dst[cc[i]] = xv[i];
                        s1 += xv[i].x + xv[i].y; s2 = fmaf(xv[i].x, xv[i].x, fmaf(xv[i].y, xv[i].y, s2));
                    }
                row_done(b, s1, s2);
            }
        } else {
            // the position row is shared by the beams (one load), the beams' token rows are fetched five at a time: one round
            // of loads instead of one per beam (this prologue opens every step: nothing hides its latency)
            float2 xv[MG_IPR];
            const float2* xp = reinterpret_cast<const float2*>(M.pos_emb + (long)pos * d);
            asm volatile(
                "ld.global.nc.v2.f32 {%0, %1}, [%10];\n\t"
                "ld.global.nc.v2.f32 {%2, %3}, [%11];\n\t"
                "ld.global.nc.v2.f32 {%4, %5}, [%12];\n\t"
                "ld.global.nc.v2.f32 {%6, %7}, [%13];\n\t"
                "ld.global.nc.v2.f32 {%8, %9}, [%14];"
                : "=f"(xv[0].x), "=f"(xv[0].y), "=f"(xv[1].x), "=f"(xv[1].y), "=f"(xv[2].x), "=f"(xv[2].y), "=f"(xv[3].x), "=f"(xv[3].y), "=f"(xv[4].x), "=f"(xv[4].y)
                : "l"(xp + cc[0]), "l"(xp + cc[1]), "l"(xp + cc[2]), "l"(xp + cc[3]), "l"(xp + cc[4]) : "memory");
#pragma unroll 1
            for (int b0 = 0; b0 < a.nb; b0 += 5) {
                uint32_t tv[5][MG_IPR];
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    const uint32_t* tp = reinterpret_cast<const uint32_t*>(M.tok_emb + (long)stok[min(b0 + q, a.nb - 1)] * d);
                    asm volatile(
                        "ld.global.nc.u32 %0, [%5];\n\t"
                        "ld.global.nc.u32 %1, [%6];\n\t"
                        "ld.global.nc.u32 %2, [%7];\n\t"
                        "ld.global.nc.u32 %3, [%8];\n\t"
                        "ld.global.nc.u32 %4, [%9];"
                        : "=r"(tv[q][0]), "=r"(tv[q][1]), "=r"(tv[q][2]), "=r"(tv[q][3]), "=r"(tv[q][4])
                        : "l"(tp + cc[0]), "l"(tp + cc[1]), "l"(tp + cc[2]), "l"(tp + cc[3]), "l"(tp + cc[4]) : "memory");
                }
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    if (b0 + q < a.nb) {                                   // warp uniform
                        float2* dst = reinterpret_cast<float2*>(sm.xs + (long)(b0 + q) * sm.ldx + d);
                        float s1 = 0.f, s2 = 0.f;
#pragma unroll
                        for (int i = 0; i < MG_IPR; ++i)
                            if (v[i]) {
                                const float x = xv[i].x + bf16lo(tv[q][i]), y = xv[i].y + bf16hi(tv[q][i]);
                                dst[cc[i]] = make_float2(x, y);
                                s1 += x + y; s2 = fmaf(x, x, fmaf(y, y, s2));
                            }
                        row_done(b0 + q, s1, s2);
                    }
                }
            }
        }
    } else {
        // x_ll holds bf16x2 LL words here (one word = one 2-column item): half the bytes of the fp32 residual words, and all
        // rows of up to five beams fit one round of loads (25 x 8 bytes in flight per thread) instead of two
        ll_wait_sentinels(x_ll, epoch, d >> 1, 8, false, a.nb, tid, 12);   // written by 16-column GEMV tiles (8 words)
        dbg.cyc(db + 2);
#pragma unroll 1
        for (int r0 = 0; r0 < a.nb; r0 += 5) {
            L5 rr[5];
            const uint2* pr[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) pr[q] = x_ll + (long)min(r0 + q, a.nb - 1) * row_items;     // rows past nb re-read the last row
            unsigned spins = 0;
            bool ok;
            do {
#pragma unroll
                for (int q = 0; q < 5; ++q) ll_load1x5(rr[q], pr[q] + cc[0], pr[q] + cc[1], pr[q] + cc[2], pr[q] + cc[3], pr[q] + cc[4]);
                ok = true;
#pragma unroll
                for (int q = 0; q < 5; ++q)
#pragma unroll
                    for (int i = 0; i < MG_IPR; ++i) ok = ok & ll_ok(rr[q].w[i], epoch);
                if (!ok) ll_backoff(spins, 2);
            } while (!ok);
            // the five rows' sums go through the butterfly together (rows past nb repeat the last row: no branches in the chain)
            float s1[5], s2[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                float2* dst = reinterpret_cast<float2*>(sm.xs + (long)min(r0 + q, a.nb - 1) * sm.ldx + d);
                s1[q] = 0.f; s2[q] = 0.f;
#pragma unroll
                for (int i = 0; i < MG_IPR; ++i)
                    if (v[i]) {
                        const float x = bf16lo((uint32_t)rr[q].w[i]), y = bf16hi((uint32_t)rr[q].w[i]);
                        dst[cc[i]] = make_float2(x, y);
                        s1[q] += x + y; s2[q] = fmaf(x, x, fmaf(y, y, s2[q]));
                    }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int q = 0; q < 5; ++q) { s1[q] += __shfl_xor_sync(0xffffffffu, s1[q], o); s2[q] += __shfl_xor_sync(0xffffffffu, s2[q], o); }
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < 5; ++q)
                    if (r0 + q < a.nb) { part[(warp * 8 + r0 + q) * 2] = s1[q]; part[(warp * 8 + r0 + q) * 2 + 1] = s2[q]; }
            }
        }
    }
    dbg.cyc(db + 11);
    mg_wait(&sm.full[ring.slot], ring.phase);                              // gamma | beta: one slot ahead of the stage's tiles
    consumer_sync();
    dbg.cyc(db + 12);
    {
        float2 ga[MG_IPR], be[MG_IPR];
        const uint32_t gaddr = smem_u32(sm.ring + (size_t)ring.slot * MG_SLOT) + tid * 8;
        lds5(ga, gaddr); lds5(be, gaddr + d * 4);
        const float inv_d = 1.f / d;
        // rows in groups of five (one group for beam size 5): the staged items and the row statistics of a group are all fetched
        // before any arithmetic
#pragma unroll 1
        for (int r0 = 0; r0 < a.nb; r0 += 5) {
            float2 xv[5][MG_IPR];
            float mean[5], rstd[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const int r = min(r0 + q, a.nb - 1);                    // (skipping the rows past nb here made ptxas spill)
                lds5(xv[q], smem_u32(sm.xs + (long)r * sm.ldx + d) + tid * 8);
                const float2 p0 = *reinterpret_cast<const float2*>(part + r * 2), p1 = *reinterpret_cast<const float2*>(part + (8 + r) * 2);
                const float2 p2 = *reinterpret_cast<const float2*>(part + (16 + r) * 2), p3 = *reinterpret_cast<const float2*>(part + (24 + r) * 2);
                mean[q] = ((p0.x + p1.x) + (p2.x + p3.x)) * inv_d;
                rstd[q] = rsqrtf(fmaxf(((p0.y + p1.y) + (p2.y + p3.y)) * inv_d - mean[q] * mean[q], 0.f) + 1e-5f);
            }
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                if (r0 + q < a.nb) {
                    uint32_t* row = reinterpret_cast<uint32_t*>(sm.xs + (long)(r0 + q) * sm.ldx);
#pragma unroll
                    for (int i = 0; i < MG_IPR; ++i)
                        if (v[i]) row[cc[i]] = pack_bf16((xv[q][i].x - mean[q]) * rstd[q] * ga[i].x + be[i].x, (xv[q][i].y - mean[q]) * rstd[q] * ga[i].y + be[i].y);
                }
            }
        }
    }
    __syncwarp();
    dbg.cyc(db + 16);
    if (lane == 0) mbar_arrive(&sm.empty[ring.slot]);
    ring.advance();
    consumer_sync();
    dbg.cyc(db + 17);
}

// poll a bf16x2 LL matrix [n_rows][row_items x 16 bytes] and copy it into xs: row -> xs row (row >> row_shift), column
// offset (row & mask) * row_items * 4
__device__ __forceinline__ void prologue_copy(const MegaSmem& sm, const uint2* __restrict__ src, uint32_t epoch, int n_rows, int row_items,
                                              int row_shift, int nb, int tid, int sent_stride, bool sent_all_rows) {
    consumer_sync();
    ll_wait_sentinels(src, epoch, (row_items * 2) << row_shift, sent_stride, sent_all_rows, nb, tid, 13);
    const int sub_mask = (1 << row_shift) - 1;
    auto put = [&](int row, int c, uint32_t w0, uint32_t w1) {
        *reinterpret_cast<uint2*>(sm.xs + (long)(row >> row_shift) * sm.ldx + ((row & sub_mask) * row_items + c) * 4) = make_uint2(w0, w1);
    };
    if (row_items <= 3 * MG_CONSUMERS) ll_read_rows3(src, epoch, n_rows, row_items, tid, 3, put);
    else ll_read_rows(src, epoch, n_rows, row_items, tid, 3, put);
    consumer_sync();
}

// ---- self-attention of the new token, unit = (beam, head) ------------------------------------------------------------------
// The cached K | V rows of the first `cap` positions are copied into shared memory (256 bytes per key, over the idle xs
// rows) BEFORE the unit waits for this step's q | k | v, so only the arithmetic is left on the critical path: thread j
// owns keys j, j + 128, ... and reads its K row in 16-byte chunks rotated by j (conflict free without padding).
// Positions >= cap (long segments only) are read from global memory on demand.
__device__ __forceinline__ void stage_self_attn(const MegaSmem& sm, const MegaArgs& a, bf16* cache_k, bf16* cache_v, uint32_t ep, int pos,
                                                int cta, int nctas, int rot_sa, int warp, int lane) {
    const MegaModel& M = c_model;
    const int d = M.d, H = M.H, tid = warp * 32 + lane, cap = a.sa_cap;
    for (int u = (cta + rot_sa) % nctas; u < a.nb * H; u += nctas) {
        const int b = u / H, h = u - b * H;
        float* ss = sm.sp;                          // [<= 449] scores
        float* sqv = sm.sq;                         // q fp32 [64]
        bf16* sknew = reinterpret_cast<bf16*>(sm.sq + 64); bf16* svnew = sknew + 64;       // new K | V row when pos >= cap
        bf16* kv = sm.xs;                           // [cap][K 64 | V 64] bf16
        consumer_sync();
        const int* tab = a.table + b * 448;
        const int n_sm = min(pos, cap);
#pragma unroll 1
        for (int j = tid; j < n_sm; j += MG_CONSUMERS) {
            const long row = ((long)tab[j] * 448 + j) * d + h * 64;
            const uint4* kp = reinterpret_cast<const uint4*>(cache_k + row);
            const uint4* vp = reinterpret_cast<const uint4*>(cache_v + row);
#pragma unroll
            for (int i = 0; i < 8; ++i) { cp_async16(kv + (long)j * 128 + i * 8, kp + i); cp_async16(kv + (long)j * 128 + 64 + i * 8, vp + i); }
        }
        u64 rq[2];                                  // words tid and 128 + tid of the unit's 192: both polls in flight together
        {
            const uint2* p0 = a.ll_qkv + (long)b * 3 * d + (tid >> 6) * d + h * 64 + (tid & 63);
            const uint2* p1 = a.ll_qkv + (long)b * 3 * d + 2 * d + h * 64 + (tid & 63);
            unsigned spins = 0;
            for (;;) {
                rq[0] = ll_load1(p0); rq[1] = tid < 64 ? ll_load1(p1) : rq[0];
                if (ll_ok(rq[0], ep) && ll_ok(rq[1], ep)) break;
                ll_backoff(spins, 4);
            }
        }
        bf16* knew = pos < cap ? kv + (long)pos * 128 : sknew;
        bf16* vnew = pos < cap ? kv + (long)pos * 128 + 64 : svnew;
        const long off = ((long)b * 448 + pos) * d + h * 64 + (tid & 63);          // the new row lives in physical slot b
        if (tid < 64) {
            sqv[tid] = __uint_as_float((uint32_t)rq[0]);
            const bf16 vb = __float2bfloat16(__uint_as_float((uint32_t)rq[1]));
            vnew[tid] = vb; cache_v[off] = vb;
        } else {
            const bf16 kb = __float2bfloat16(__uint_as_float((uint32_t)rq[0]));
            knew[tid & 63] = kb; cache_k[off] = kb;
        }
        if (h == 0 && tid == 0) a.table[b * 448 + pos] = b;
        cp_async_wait_all();
        consumer_sync();
        float m = -INFINITY;
#pragma unroll 1
        for (int j = tid; j <= pos; j += MG_CONSUMERS) {
            const bf16* kp = j < cap ? kv + (long)j * 128 : (j == pos ? sknew : cache_k + ((long)tab[j] * 448 + j) * d + h * 64);
            const int rot = j < cap ? j : 0;
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int ch = (i + rot) & 7;
                const uint4 kvv = *reinterpret_cast<const uint4*>(kp + ch * 8);
                const float4 qa = reinterpret_cast<const float4*>(sqv)[2 * ch], qb = reinterpret_cast<const float4*>(sqv)[2 * ch + 1];
                s = fmaf(qa.x, bf16lo(kvv.x), s); s = fmaf(qa.y, bf16hi(kvv.x), s); s = fmaf(qa.z, bf16lo(kvv.y), s); s = fmaf(qa.w, bf16hi(kvv.y), s);
                s = fmaf(qb.x, bf16lo(kvv.z), s); s = fmaf(qb.y, bf16hi(kvv.z), s); s = fmaf(qb.z, bf16lo(kvv.w), s); s = fmaf(qb.w, bf16hi(kvv.w), s);
            }
            s += a.mask ? a.mask[j == pos ? 448 : j] : 0.f;
            ss[j] = s;
            m = fmaxf(m, s);
        }
        m = warp_max(m);
        if (lane == 0) sm.stat[warp] = m;
        consumer_sync();
        m = fmaxf(fmaxf(sm.stat[0], sm.stat[1]), fmaxf(sm.stat[2], sm.stat[3]));
        float lsum = 0.f;
#pragma unroll 1
        for (int j = tid; j <= pos; j += MG_CONSUMERS) { const float p = __expf(ss[j] - m); ss[j] = p; lsum += p; }   // own entries only
        lsum = warp_sum(lsum);
        if (lane == 0) sm.stat[8 + warp] = lsum;
        consumer_sync();
        lsum = (sm.stat[8] + sm.stat[9]) + (sm.stat[10] + sm.stat[11]);
        float o0 = 0.f, o1 = 0.f;
#pragma unroll 4
        for (int j = warp; j <= pos; j += MG_CWARPS) {
            const bf16* vp = j < cap ? kv + (long)j * 128 + 64 : (j == pos ? svnew : cache_v + ((long)tab[j] * 448 + j) * d + h * 64);
            const uint32_t vv = *reinterpret_cast<const uint32_t*>(vp + 2 * lane);
            const float p = ss[j];
            o0 = fmaf(p, bf16lo(vv), o0); o1 = fmaf(p, bf16hi(vv), o1);
        }
        float* sred = sm.red;                       // [4][64]
        sred[warp * 64 + 2 * lane] = o0; sred[warp * 64 + 2 * lane + 1] = o1;
        consumer_sync();
        if (tid < 32) {
            const float e0 = (sred[2 * tid] + sred[64 + 2 * tid]) + (sred[128 + 2 * tid] + sred[192 + 2 * tid]);
            const float e1 = (sred[2 * tid + 1] + sred[64 + 2 * tid + 1]) + (sred[128 + 2 * tid + 1] + sred[192 + 2 * tid + 1]);
            ll_store(a.ll_att + ((long)b * d + h * 64) / 2 + tid, pack_bf16(e0 / lsum, e1 / lsum), ep);
        }
    }
}

// ---- cross-attention, unit = (head, key split): K/V of a head are read once for all beams; the CTA of a head's last
// (shortest) split then merges the head's partials in split order --------------------------------------------------------
__device__ __forceinline__ void stage_cross_attn(const MegaSmem& sm, Ring& ring, const MegaArgs& a, uint32_t ep, int cta, int nctas, int rot_ca,
                                                 int warp, int lane) {
    const MegaModel& M = c_model;
    const int d = M.d, H = M.H, tid = warp * 32 + lane, gq = lane >> 2, tq = lane & 3;
    constexpr int n_ktiles = CROSS_KEYS_PAD / 16, n_vkc = CROSS_KEYS_PAD / 32;
    for (int u = (cta + rot_ca) % nctas; u < H * MG_N_SPLITS; u += nctas) {
        const int h = u / MG_N_SPLITS, s = u - h * MG_N_SPLITS;
        const int t0 = s * MG_SPLIT_TILES, nt = min(MG_SPLIT_TILES, n_ktiles - t0);
        const int nkeys = min(nt * 16, N_AUDIO_CTX - t0 * 16);                  // valid (unpadded) keys of the split
        consumer_sync();
        {   // q of this head -> xs rows (bf16, 64 columns): items tid and tid + 128 of the nb * 32 two-word items
            u64 rq[2][2];
            const bool has0 = tid < a.nb * 32, has1 = tid + 128 < a.nb * 32;
            const uint2* p0 = a.ll_q + (long)(tid >> 5) * d + h * 64 + (tid & 31) * 2;
            const uint2* p1 = p0 + 4L * d;
            unsigned spins = 0;
            for (;;) {
                bool ok = true;
                if (has0) ll_load2(p0, rq[0][0], rq[0][1]);
                if (has1) ll_load2(p1, rq[1][0], rq[1][1]);
                if (has0) ok = ll_ok(rq[0][0], ep) && ll_ok(rq[0][1], ep);
                if (has1) ok = ok && ll_ok(rq[1][0], ep) && ll_ok(rq[1][1], ep);
                if (ok) break;
                ll_backoff(spins, 5);
            }
            uint32_t* dst = reinterpret_cast<uint32_t*>(sm.xs + (long)(tid >> 5) * sm.ldx) + (tid & 31);
            if (has0) *dst = pack_bf16(__uint_as_float((uint32_t)rq[0][0]), __uint_as_float((uint32_t)rq[0][1]));
            if (has1) dst[2 * sm.ldx] = pack_bf16(__uint_as_float((uint32_t)rq[1][0]), __uint_as_float((uint32_t)rq[1][1]));   // 4 rows further
        }
        consumer_sync();
        // scores: one slot carries the split's key tiles (2 blocks each); warp w takes tiles w, w + 4, ...
        mg_wait(&sm.full[ring.slot], ring.phase);
        {
            const uint4* sl = reinterpret_cast<const uint4*>(sm.ring + (size_t)ring.slot * MG_SLOT) + lane;
            const uint4 xb0 = *reinterpret_cast<const uint4*>(sm.xs + (long)gq * sm.ldx + tq * 8);
            const uint4 xb1 = *reinterpret_cast<const uint4*>(sm.xs + (long)gq * sm.ldx + 32 + tq * 8);
#pragma unroll
            for (int q = 0; q < (MG_SPLIT_TILES + MG_CWARPS - 1) / MG_CWARPS; ++q) {
                const int tt = warp + MG_CWARPS * q;
                if (tt < nt) {
                    float acc[4] = {0.f, 0.f, 0.f, 0.f};
                    mg_mma(acc, sl[tt * 128], sl[tt * 128 + 32], xb0);
                    mg_mma(acc, sl[tt * 128 + 64], sl[tt * 128 + 96], xb1);
                    float* spb = sm.sp + tt * 16;
                    spb[(2 * tq) * MG_SPLIT_KEYS + gq] = acc[0]; spb[(2 * tq + 1) * MG_SPLIT_KEYS + gq] = acc[1];
                    spb[(2 * tq) * MG_SPLIT_KEYS + gq + 8] = acc[2]; spb[(2 * tq + 1) * MG_SPLIT_KEYS + gq + 8] = acc[3];
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[ring.slot]);
        ring.advance();
        consumer_sync();
        // partial softmax per beam over the split's valid keys; p (bf16) becomes the B operand of P V
        for (int b = warp; b < a.nb; b += MG_CWARPS) {
            float sv[MG_SPLIT_KEYS / 32];
            float m = -INFINITY;
#pragma unroll
            for (int i = 0; i < MG_SPLIT_KEYS / 32; ++i) { const int j = lane + 32 * i; sv[i] = j < nkeys ? sm.sp[b * MG_SPLIT_KEYS + j] : -INFINITY; m = fmaxf(m, sv[i]); }
            m = warp_max(m);
            float lsum = 0.f;
#pragma unroll
            for (int i = 0; i < MG_SPLIT_KEYS / 32; ++i) {
                const float p = __expf(sv[i] - m);                              // exp(-inf) = 0 for the padded keys
                lsum += p;
                sm.xs[(long)b * sm.ldx + lane + 32 * i] = __float2bfloat16(p);
            }
            lsum = warp_sum(lsum);
            if (lane == 0) { sm.stat[b] = m; sm.stat[8 + b] = lsum; }
        }
        consumer_sync();
        // o[dim][beam] = V^T[dim][key] p[key][beam]: the slot holds the split's key blocks of the four dim tiles; warp dt owns tile dt
        const int nkc = min(MG_SPLIT_TILES / 2, n_vkc - s * (MG_SPLIT_TILES / 2));
        uint2* part = a.ll_cap + ((long)h * MG_N_SPLITS + s) * 8 * 66;
        mg_wait(&sm.full[ring.slot], ring.phase);
        {
            const uint4* sl = reinterpret_cast<const uint4*>(sm.ring + (size_t)ring.slot * MG_SLOT + (size_t)warp * nkc * 1024) + lane;
            float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int kc = 0; kc < MG_SPLIT_TILES / 2; ++kc)
                if (kc < nkc) mg_mma(acc[kc & 1], sl[kc * 64], sl[kc * 64 + 32], *reinterpret_cast<const uint4*>(sm.xs + (long)gq * sm.ldx + kc * 32 + tq * 8));
            const int b0 = 2 * tq, dim = warp * 16 + gq;
            if (b0 < a.nb) {
                ll_store(part + b0 * 66 + 2 + dim, __float_as_uint(acc[0][0] + acc[1][0]), ep);
                ll_store(part + b0 * 66 + 2 + dim + 8, __float_as_uint(acc[0][2] + acc[1][2]), ep);
            }
            if (b0 + 1 < a.nb) {
                ll_store(part + (b0 + 1) * 66 + 2 + dim, __float_as_uint(acc[0][1] + acc[1][1]), ep);
                ll_store(part + (b0 + 1) * 66 + 2 + dim + 8, __float_as_uint(acc[0][3] + acc[1][3]), ep);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[ring.slot]);
        ring.advance();
        if (tid < a.nb) { ll_store(part + tid * 66, __float_as_uint(sm.stat[tid]), ep); ll_store(part + tid * 66 + 1, __float_as_uint(sm.stat[8 + tid]), ep); }
        if (s == MG_N_SPLITS - 1) {
            const uint2* ph = a.ll_cap + (long)h * MG_N_SPLITS * 8 * 66;
#pragma unroll 1
            for (int e = tid; e < a.nb * 64; e += MG_CONSUMERS) {
                const int b = e >> 6, c = e & 63;
                u64 rm[MG_N_SPLITS], rl[MG_N_SPLITS], ro[MG_N_SPLITS];         // all loads of a round are in flight together
                unsigned spins = 0;
                bool ok;
                do {
                    ok = true;
#pragma unroll
                    for (int q = 0; q < MG_N_SPLITS; ++q) { const uint2* pq = ph + (q * 8 + b) * 66; ll_load2(pq, rm[q], rl[q]); ro[q] = ll_load1(pq + 2 + c); }
#pragma unroll
                    for (int q = 0; q < MG_N_SPLITS; ++q) ok = ok && ll_ok(rm[q], ep) && ll_ok(rl[q], ep) && ll_ok(ro[q], ep);
                    if (!ok) ll_backoff(spins, 6);
                } while (!ok);
                float mm = -INFINITY;
#pragma unroll
                for (int q = 0; q < MG_N_SPLITS; ++q) mm = fmaxf(mm, __uint_as_float((uint32_t)rm[q]));
                float ll = 0.f, oo = 0.f;
#pragma unroll
                for (int q = 0; q < MG_N_SPLITS; ++q) {
                    const float w = __expf(__uint_as_float((uint32_t)rm[q]) - mm);
                    ll = fmaf(__uint_as_float((uint32_t)rl[q]), w, ll); oo = fmaf(__uint_as_float((uint32_t)ro[q]), w, oo);
                }
                const float o = oo / ll;
                const float nxt = __shfl_down_sync(0xffffffffu, o, 1);
                if (!(c & 1)) ll_store(a.ll_catt + ((long)b * d + h * 64 + c) / 2, pack_bf16(o, nxt), ep);
            }
        }
    }
}

// ---- grid barrier (consumers only) for the sampling tail: arrivals count monotonically within a launch (barrier k completes at
// k * nctas); the last CTA to leave the kernel re-arms the word --------------------------------------------------------------
__device__ __forceinline__ void grid_sync(unsigned* bar, unsigned& k, int nctas, int tid) {
    consumer_sync();                                   // orders this CTA's stores before thread 0's release (cumulativity)
    if (tid == 0) {
        ++k;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        const unsigned target = k * (unsigned)nctas;
        unsigned spins = 0, v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if (++spins > MG_SPIN_LIMIT) mg_timeout(100 + (int)k);
        } while (v < target);
    }
    consumer_sync();
}

// ---- the stage table: stage `it` of the step (8 per layer + the vocabulary projection) ----------------------------------------
enum { ST_QKV = 0, ST_SA, ST_OUT, ST_CQ, ST_CA, ST_CO, ST_M1, ST_M2, ST_VOCAB };
struct StageDesc {                     // what the producer needs: the weight matrix of a GEMV stage and who owns which tile
    const bf16* w; int n_tiles, n_kc, vcta;
    const float *ln_g, *ln_b;          // LayerNorm in front of the stage (gamma | beta ride in one slot) or nullptr
};
__device__ __forceinline__ int stage_rot(int st, int nctas) {
    // stages with fewer units than CTAs start at different CTAs so that every CTA streams about the same bytes per layer
    return st == ST_CQ ? nctas / 2 : st == ST_CA ? nctas / 4 : st == ST_CO ? (3 * nctas) / 4 : st == ST_M2 ? nctas / 3 : st == ST_SA ? nctas / 8 : 0;
}
__device__ __forceinline__ StageDesc stage_desc(int st, int l, int cta, int nctas) {
    const MegaModel& M = c_model;
    const MegaLayer& L = M.layers[l];
    const int d = M.d;
    StageDesc s{nullptr, d / 16, d >> 5, (cta + stage_rot(st, nctas)) % nctas, nullptr, nullptr};
    switch (st) {
    case ST_QKV: s.w = L.qkv; s.n_tiles = 3 * d / 16; s.ln_g = L.ln1_w; s.ln_b = L.ln1_b; break;
    case ST_OUT: s.w = L.attn_out; break;
    case ST_CQ: s.w = L.cross_q; s.ln_g = L.ln2_w; s.ln_b = L.ln2_b; break;
    case ST_CO: s.w = L.cross_out; break;
    case ST_M1: s.w = L.mlp1; s.n_tiles = 4 * d / 16; s.ln_g = L.ln3_w; s.ln_b = L.ln3_b; break;
    case ST_M2: s.w = L.mlp2; s.n_kc = d >> 3; break;
    case ST_VOCAB: s.w = M.tok_emb_frag; s.n_tiles = M.n_tiles_vocab; s.ln_g = M.ln_w; s.ln_b = M.ln_b; break;
    default: break;
    }
    return s;
}

// =================================================================================================================
__global__ void __launch_bounds__(MG_THREADS, 1) decoder_mega_kernel(const __grid_constant__ MegaArgs a) {
    extern __shared__ __align__(128) uint8_t mg_raw[];
    const MegaModel& M = c_model;
    const int d = M.d, H = M.H, nctas = gridDim.x, cta = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (a.d_done && *a.d_done) return;                 // uniform: the decode already finished (graph replays past the end)

    // xs comes first: the MMA B operand always reads 8 rows, rows >= xs_rows alias what follows (harmless garbage in the
    // output columns of beams that do not exist)
    MegaSmem sm;
    sm.ldx = a.xs_cols + MG_XS_PAD;
    sm.xs = reinterpret_cast<bf16*>(mg_raw);
    sm.red = reinterpret_cast<float*>(mg_raw + (size_t)a.xs_rows * sm.ldx * 2);    // [2][4][128]
    sm.sp = sm.red + 2 * MG_CWARPS * 128;              // [8][MG_SPLIT_KEYS] cross scores / self-attention scores [<= 449]
    sm.sq = sm.sp + 8 * MG_SPLIT_KEYS;                 // [8][64]
    sm.stat = sm.sq + 8 * 64;                          // 64 floats
    sm.full = reinterpret_cast<uint64_t*>(sm.stat + 64);
    sm.empty = sm.full + MG_MAX_SLOTS;
    sm.ring = mg_raw + a.ring_offset;

    if (tid == 0) {
        for (int s = 0; s < MG_MAX_SLOTS; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], MG_CWARPS); }
        fence_barrier_init();
    }
    const unsigned seq = *a.seq;                       // written by the previous launch
    __syncthreads();

    const int pos = a.d_pos ? *a.d_pos : a.text_offset;    // text_offset of this step
    const int n_stages = M.Ld * 8 + (a.no_vocab ? 0 : 1);

    if (warp == MG_CWARPS) {
        // =========================================== producer ===========================================
        if (lane != 0) return;
        if (a.dbg_delay) { const long long t0 = clock64(); while (clock64() - t0 < a.dbg_delay) {} }     // experiment: hold the weight stream back
        Ring ring{0, 0, a.n_slots};
        constexpr int n_ktiles = CROSS_KEYS_PAD / 16, n_vkc = CROSS_KEYS_PAD / 32;
        const long head_elems = (long)64 * CROSS_KEYS_PAD;
        for (int it = 0; it < n_stages; ++it) {
            const int l = it >> 3, st = it == M.Ld * 8 ? ST_VOCAB : (it & 7);
            if (st == ST_SA) continue;
            if (st == ST_CA) {
                for (int u = (cta + stage_rot(ST_CA, nctas)) % nctas; u < H * MG_N_SPLITS; u += nctas) {
                    // cross-attention unit (head, split): one slot of K tiles, one slot with the four V^T dim tiles
                    const int h = u / MG_N_SPLITS, s = u - h * MG_N_SPLITS;
                    const bf16* kf = a.ckv_frag + (long)(l * 2) * H * head_elems + h * head_elems;
                    const bf16* vf = a.ckv_frag + (long)(l * 2 + 1) * H * head_elems + h * head_elems;
                    const int t0 = s * MG_SPLIT_TILES, nt = min(MG_SPLIT_TILES, n_ktiles - t0);
                    const int kc0 = s * (MG_SPLIT_TILES / 2), nkc = min(MG_SPLIT_TILES / 2, n_vkc - kc0);
                    mg_wait(&sm.empty[ring.slot], ring.phase ^ 1);
                    mbar_expect_tx(&sm.full[ring.slot], (uint32_t)nt * 2048);
                    bulk_g2s(sm.ring + (size_t)ring.slot * MG_SLOT, kf + (long)t0 * 1024, (uint32_t)nt * 2048, &sm.full[ring.slot]);
                    ring.advance();
                    mg_wait(&sm.empty[ring.slot], ring.phase ^ 1);
                    mbar_expect_tx(&sm.full[ring.slot], (uint32_t)nkc * 4096);
#pragma unroll 1
                    for (int dt = 0; dt < 4; ++dt)
                        bulk_g2s(sm.ring + (size_t)ring.slot * MG_SLOT + (size_t)dt * nkc * 1024, vf + ((long)dt * n_vkc + kc0) * 512, (uint32_t)nkc * 1024,
                                 &sm.full[ring.slot]);
                    ring.advance();
                }
                continue;
            }
            const StageDesc sd = stage_desc(st, l < M.Ld ? l : 0, cta, nctas);
            int u0, u1;
            gemv_range(sd.n_tiles, sd.vcta, nctas, u0, u1);
            if (sd.ln_g && u1 > u0) {
                mg_wait(&sm.empty[ring.slot], ring.phase ^ 1);
                mbar_expect_tx(&sm.full[ring.slot], (uint32_t)d * 8);
                bulk_g2s(sm.ring + (size_t)ring.slot * MG_SLOT, sd.ln_g, (uint32_t)d * 4, &sm.full[ring.slot]);
                bulk_g2s(sm.ring + (size_t)ring.slot * MG_SLOT + (size_t)d * 4, sd.ln_b, (uint32_t)d * 4, &sm.full[ring.slot]);
                ring.advance();
            }
#pragma unroll 1
            for (int t = u0; t < u1; ++t)
#pragma unroll 1
                for (int c0 = 0; c0 < sd.n_kc; c0 += MG_SLOT_BLOCKS) {
                    const uint32_t bytes = (uint32_t)min(MG_SLOT_BLOCKS, sd.n_kc - c0) * 1024;
                    mg_wait(&sm.empty[ring.slot], ring.phase ^ 1);
                    mbar_expect_tx(&sm.full[ring.slot], bytes);
                    bulk_g2s(sm.ring + (size_t)ring.slot * MG_SLOT, sd.w + ((long)t * sd.n_kc + c0) * 512, bytes, &sm.full[ring.slot]);
                    ring.advance();
                }
        }
        return;
    }

    // =============================================== consumers ===============================================
    Ring ring{0, 0, a.n_slots};
    int red_buf = 0;
    Dbg dbg{a.dbg + (size_t)cta * MEGA_DBG_LD, a.dbg != nullptr && tid == 0};
    dbg.mark(0);

    for (int it = 0; it < n_stages; ++it) {
        const int l = it >> 3, st = it == M.Ld * 8 ? ST_VOCAB : (it & 7);
        const MegaLayer& L = M.layers[l < M.Ld ? l : 0];
        const uint32_t ep = seq * 64u + (uint32_t)l + 1u, ep_prev = ep - 1u;      // ep_prev: x3 of the layer below
        if (st == ST_SA) {
            dbg.mark(2 * it + 1);
            stage_self_attn(sm, a, a.mkv + (long)(2 * l) * a.kv_stride, a.mkv + (long)(2 * l + 1) * a.kv_stride, ep, pos, cta, nctas, stage_rot(ST_SA, nctas), warp, lane);
            dbg.mark(2 * it + 2);
            continue;
        }
        if (st == ST_CA) {
            dbg.mark(2 * it + 1);
            stage_cross_attn(sm, ring, a, ep, cta, nctas, stage_rot(ST_CA, nctas), warp, lane);
            dbg.mark(2 * it + 2);
            continue;
        }
        const StageDesc sd = stage_desc(st, l < M.Ld ? l : 0, cta, nctas);
        GemvStage g{sd.n_tiles, sd.n_kc, sd.vcta, EPI_LL_F32, nullptr, RES_NONE, nullptr, ep, nullptr, nullptr, (long)d, nullptr, d, ep};
        const uint2* pro_src = nullptr; uint32_t pro_ep = ep;
        switch (st) {
        case ST_QKV: pro_src = a.ll_x3b; pro_ep = ep_prev; g.bias = L.qkv_b; g.out_ll = a.ll_qkv; g.ld_out = 3L * d; g.n_valid = 3 * d; break;
        case ST_OUT:
            pro_src = a.ll_att; g.bias = L.attn_out_b; g.out_ll = a.ll_x1; g.out_llb = a.ll_x1b;
            g.res_mode = l == 0 ? (a.x_in ? RES_XIN : RES_EMBED) : RES_LL; g.res_ll = a.ll_x3; g.res_epoch = ep_prev;
            break;
        case ST_CQ: pro_src = a.ll_x1b; g.bias = L.cross_q_b; g.out_ll = a.ll_q; break;
        case ST_CO: pro_src = a.ll_catt; g.bias = L.cross_out_b; g.out_ll = a.ll_x2; g.out_llb = a.ll_x2b; g.res_mode = RES_LL; g.res_ll = a.ll_x1; break;
        case ST_M1: pro_src = a.ll_x2b; g.epi = EPI_LL_GELU_BF16; g.bias = L.mlp1_b; g.out_ll = a.ll_hid; g.ld_out = 4L * d; g.n_valid = 4 * d; break;
        case ST_M2: pro_src = a.ll_hid; g.bias = L.mlp2_b; g.out_ll = a.ll_x3; g.out_llb = a.ll_x3b; g.res_mode = RES_LL; g.res_ll = a.ll_x2; break;
        default:    pro_src = a.ll_x3b; pro_ep = seq * 64u + (uint32_t)M.Ld; g.epi = EPI_LOGITS; g.out_f32 = a.logits; g.ld_out = a.ld_logits; g.n_valid = M.V; break;
        }
        int u0, u1;
        gemv_range(g.n_tiles, g.vcta, nctas, u0, u1);
        if (u1 > u0) {                                 // a CTA without a tile in this stage does not read its input at all
            if (sd.ln_g) prologue_ln(sm, ring, a, it == 0 ? PRO_EMBED : PRO_LL, pro_src, pro_ep, pos, warp, lane, dbg, it == 0 ? 200 : (it == 6 ? 220 : 100000));
            else if (st == ST_M2) prologue_copy(sm, pro_src, pro_ep, 2 * a.nb, d >> 1, 1, a.nb, tid, 8, false);   // hid: 2 half rows per beam; 16-column tiles
            else prologue_copy(sm, pro_src, pro_ep, a.nb, d >> 2, 0, a.nb, tid, 32, true);                          // attention: 64 columns per (beam, head)
        }
        dbg.mark(2 * it + 1);
        stage_gemv(sm, ring, red_buf, g, a, pos, nctas, warp, lane);
        dbg.mark(2 * it + 2);
    }
    if (a.do_sampling) {
        // ---- tail of the device-resident decode loop: two real grid barriers (every CTA's logits are needed) ----
        unsigned bar_k = 0;
        grid_sync(a.barrier, bar_k, nctas, tid);
        dbg.mark(2 * n_stages + 1);
        // Everything the two phases read is first copied into shared memory with 4-byte cp.async (one L2 round trip, all
        // in flight together); the generic bodies of sampling_dev.cuh then run against those copies.
        uint8_t* scratch = reinterpret_cast<uint8_t*>(sm.xs);
#ifdef B200_PROBES
        if (tid == 0 && cta == 0) g_probe = a.dbg;                     // experiments: cycle probes of CTA 0 in the bodies
#endif
        const int tb = a.spec.timestamp_begin, V = a.spec.n_vocab;
        for (int u = cta; u < SAMPLE_CHUNKS * a.nb; u += nctas) {
            const int chunk = u % SAMPLE_CHUNKS, b = u / SAMPLE_CHUNKS;
            int lo, hi;
            if (chunk < SAMPLE_TEXT_CHUNKS) { const int per = (tb + SAMPLE_TEXT_CHUNKS - 1) / SAMPLE_TEXT_CHUNKS; lo = chunk * per; hi = min(tb, lo + per); }
            else { lo = tb; hi = V; }
            float* s_logits = reinterpret_cast<float*>(scratch);                       // [<= 2048]
            uint8_t* s_sup = scratch + 8192;                                            // [<= 2064] suppress flags from lo & ~3
            int* s_tok = reinterpret_cast<int*>(scratch + 8192 + 2176);                 // [449] token row of beam b
            DecodeState* s_st = reinterpret_cast<DecodeState*>(scratch + 8192 + 2176 + 1856);
            consumer_sync();
            const float* gl = a.logits + (long)b * a.ld_logits + lo;
            for (int i = tid; i < hi - lo; i += MG_CONSUMERS) cp_async4(s_logits + i, gl + i);
            const int lo4 = lo & ~3;
            for (int i = tid; i < (hi - lo4 + 3) / 4; i += MG_CONSUMERS) cp_async4(s_sup + 4 * i, a.spec.d_suppress + lo4 + 4 * i);
            for (int i = tid; i < DEC_TOK_LD; i += MG_CONSUMERS) cp_async4(s_tok + i, a.tokens + b * DEC_TOK_LD + i);
            for (int i = tid; i < (int)(sizeof(DecodeState) / 4); i += MG_CONSUMERS) cp_async4(reinterpret_cast<int*>(s_st) + i, reinterpret_cast<const int*>(a.st) + i);
            cp_async_wait_all();
            consumer_sync();
            SampleArgs sa;                         // the bodies index by absolute token / beam: rebase the pointers onto the copies
            sa.logits = s_logits - ((long)b * a.ld_logits + lo); sa.ld_logits = a.ld_logits; sa.tokens = s_tok - b * DEC_TOK_LD; sa.st = s_st;
            sa.spec = a.spec; sa.spec.d_suppress = s_sup - lo4; sa.nb = a.nb; sa.k = a.k;
            sa.part = a.sp; sa.cand_lp = a.cand_lp; sa.cand_tok = a.cand_tok;
            sample_partial_body<MG_CONSUMERS>(sa, chunk, b, tid, ConsumerSync());
        }
        dbg.mark(2 * n_stages + 2);
        grid_sync(a.barrier, bar_k, nctas, tid);
        dbg.mark(2 * n_stages + 3);
        if (cta == 0) {                            // merge + greedy / beam update
            SamplePartials* s_part = reinterpret_cast<SamplePartials*>(scratch);
            constexpr int PART_BYTES = (int)((sizeof(SamplePartials) + 127) / 128 * 128);
            int* stage = reinterpret_cast<int*>(scratch + PART_BYTES);                 // [8][449] token rows
            int* stage2 = stage + DEC_MAX_BEAMS * DEC_TOK_LD;                            // [8][448] slot table
            DecodeState* s_st = reinterpret_cast<DecodeState*>(stage2 + DEC_MAX_BEAMS * 448);
            consumer_sync();
            for (int i = tid; i < (int)(sizeof(SamplePartials) / 4); i += MG_CONSUMERS) cp_async4(reinterpret_cast<int*>(s_part) + i, reinterpret_cast<const int*>(a.sp) + i);
            for (int i = tid; i < a.nb * DEC_TOK_LD; i += MG_CONSUMERS) cp_async4(stage + i, a.tokens + i);
            for (int i = tid; i < a.nb * 448; i += MG_CONSUMERS) cp_async4(stage2 + i, a.table + i);
            for (int i = tid; i < (int)(sizeof(DecodeState) / 4); i += MG_CONSUMERS) cp_async4(reinterpret_cast<int*>(s_st) + i, reinterpret_cast<const int*>(a.st) + i);
            cp_async_wait_all();
            consumer_sync();
            BeamUpdateArgs ba;
            ba.part = s_part; ba.timestamp_begin = a.spec.timestamp_begin; ba.update = 1; ba.cand_lp = a.cand_lp; ba.cand_tok = a.cand_tok;
            ba.nb = a.nb; ba.k = a.k; ba.tokens = a.tokens; ba.table = a.table; ba.fin_tokens = a.fin_tokens; ba.st = s_st;
            ba.eot = a.spec.eot; ba.n_text_ctx = N_TEXT_CTX;
            const int was_done = s_st->done;
            beam_update_body<MG_CONSUMERS, true>(ba, stage, stage2, tid, ConsumerSync());
            consumer_sync();
            if (!was_done)                         // publish the updated decode state (pos / done steer the next launch)
                for (int i = tid; i < (int)(sizeof(DecodeState) / 4); i += MG_CONSUMERS) reinterpret_cast<int*>(a.st)[i] = reinterpret_cast<const int*>(s_st)[i];
        }
        dbg.mark(2 * n_stages + 4);
    }
    // the last CTA to leave re-arms the barrier and advances the launch sequence number: by then every CTA has read it
    if (tid == 0) {
        const unsigned old = atomicAdd(&a.barrier[1], 1u);
        if (old == (unsigned)nctas - 1) { a.barrier[0] = 0; a.barrier[1] = 0; *a.seq = seq + 1; }
    }
}

void mega_set_model(const MegaModel& m) { B200_CHECK(cudaMemcpyToSymbol(c_model, &m, sizeof(MegaModel))); }

int mega_xs_rows(int nb) { return nb <= 5 ? 5 : 8; }
int mega_slots(int nb) { return nb <= 5 ? MG_MAX_SLOTS : MG_MAX_SLOTS - 1; }
size_t mega_ring_offset(int xs_cols, int nb) {                 // everything in front of the ring, 128-byte aligned
    const size_t b = (size_t)mega_xs_rows(nb) * (xs_cols + MG_XS_PAD) * 2 + (2 * MG_CWARPS * 128 + 8 * MG_SPLIT_KEYS + 8 * 64 + 64) * 4 +
                     2 * MG_MAX_SLOTS * 8;
    return (b + 127) / 128 * 128;
}
size_t mega_smem_bytes(int xs_cols, int nb) { return mega_ring_offset(xs_cols, nb) + (size_t)mega_slots(nb) * MG_SLOT; }
int mega_sa_cap(int xs_cols, int nb) {                          // cached positions a self-attention unit can stage in the xs rows
    const int rows = (int)((size_t)mega_xs_rows(nb) * (xs_cols + MG_XS_PAD) * 2 / 256);
    return rows < 256 ? rows : 256;
}

bool mega_launch(const MegaArgs& a, int n_ctas, cudaStream_t s) {
    const size_t smem = mega_smem_bytes(a.xs_cols, a.nb);
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
    // experiment (B200_MEGA_NOCOOP=1): a plain launch - every CTA needs a whole SM, so the grid is co-resident whenever
    // the lanes' grids add up to at most the SM count, which api_decode.cu guarantees
    static const bool nocoop = getenv("B200_MEGA_NOCOOP") && atoi(getenv("B200_MEGA_NOCOOP")) != 0;
    cudaError_t e = nocoop ? cudaLaunchKernel((const void*)decoder_mega_kernel, dim3(n_ctas), dim3(MG_THREADS), params, smem, s)
                           : cudaLaunchCooperativeKernel((const void*)decoder_mega_kernel, dim3(n_ctas), dim3(MG_THREADS), params, smem, s);
    ++g_launch_count;
    if (e != cudaSuccess) { cudaGetLastError(); record_error("decoder_mega launch: %s", cudaGetErrorString(e)); return false; }
    return true;
}

}  // namespace b200
