// decoder1 as ONE persistent kernel per token step for a BATCH of windows (whisper/decoder.py:241-257, 261-327;
// decoding.py:707-737).  Rows = windows x beams are the N dimension of every GEMV, so the 316 MB of decoder weights are
// streamed from HBM once per step for all windows of the batch; what stays per window is its cross
// K/V (read once for the window's beams, decoder.py:84-89), its KV-cache rows and its decode state.
//
// The step is a chain of ~33 dependent stages (per layer: LN+QKV, self-attention, out-proj, LN+cross-q, cross-attention with
// the merge of its key splits, cross-out, LN+MLP1, MLP2; then LN+vocabulary).  One CTA per SM stays resident:
//
//   * the producer warp (one thread) walks the CTA's static byte schedule - its share of every stage's weights and of the
//     windows' cross K/V, all stored fragment-major by the exporter so a share is one contiguous run - and streams it with
//     cp.async.bulk into a ring of 40 KB slots (one 16-row weight tile x K <= 1280 per slot), running as far ahead of the
//     math as the ring allows;
//   * the CW consumer warps wait for a slot, feed it to mma.m16n8k16 (weights = A fragments straight from the slot, 8 rows
//     per n-tile = the N dimension, NT n-tiles, activations = B fragments from a bf16 copy in shared memory), reduce across
//     warps and apply the stage epilogue.
//
// Stages hand their activations to each other WITHOUT grid barriers: every cross-CTA activation is an "LL" word, a 64-bit
// {payload, epoch} pair written with one single-copy-atomic st.relaxed.gpu.b64 and polled by its readers until the epoch
// matches (epoch = launch sequence number * 64 + layer + 1).  The flag travels inside the same word as the data, so no
// fence is needed; a buffer is rewritten one layer later and a CTA can only be two stages ahead of the slowest one.
// Readers first wait on one sentinel word per producer unit and only then read everything (tools/bench_barrier.cu,
// bench_bcast.cu measured both choices).
//
// Against the first persistent kernel (one window, 4 consumer warps at the 255-register limit) this one is templated on
// the n-tile count and the consumer-warp count: with 8 consumer warps the per-thread share of every prologue (LayerNorm,
// activation copies) halves, which is where a step's time goes (tools/step_timeline.py), and nothing needs to be batched
// by hand in asm to stay inside the register budget.  LayerNorm rows are staged as the bf16 pairs they arrive as and
// normalised in place; the MLP hidden row (4d columns) is consumed in K chunks of xs_cols columns so that 40 rows fit.
//
// Ring protocol.  full[slot] (1 arrival + tx bytes) / empty[slot] (4 arrivals) mbarriers; a wait names only the PARITY of the
// use it waits for, so a waiter must never be a whole ring cycle away from the barrier in either direction:
//   * whoever waits on a slot either releases it itself (GEMV tile group: 4 warps wait, 4 warps arrive) or passes a CTA-wide
//     barrier before the release (LayerNorm gamma | beta, cross-attention keys); warps that do not read a slot do not wait
//     on it (cross-attention values) - a late waiter would otherwise find the slot released, refilled, and the parity back
//     where it started, and wait forever;
//   * the two GEMV tile groups skip each other's slots without waiting, so the previous use of a group's slot may be one it
//     never watched: before its own wait it checks seen[slot] (the use index the last waiter observed complete) - otherwise
//     a wait issued while the previous use is still in flight passes at once, on stale bytes, and its release corrupts the
//     empty barrier's count (the large-v3 hang: 18-CTA lanes, 4-slot MLP2 tiles).
// Every bounded wait that gives up stores {where, CTA, thread} and the per-CTA stage table in a host-mapped word before the
// trap (db_fault_word), so a protocol bug reads as a log line instead of a hung box.
#include "decoder_batch.cuh"

#include <stdlib.h>
#include <algorithm>
#include <string.h>

namespace b200 {

constexpr int DB_SLOT_BLOCKS = 40, DB_SLOT = DB_SLOT_BLOCKS * 1024, DB_MAX_SLOTS = 4;
constexpr int DB_XS_PAD = 32;
constexpr int DB_SPLIT_TILES = 14;                    // 16-key tiles per cross-attention split (224 keys)
constexpr int DB_SPLIT_KEYS = DB_SPLIT_TILES * 16;
static_assert(DB_N_SPLITS == (CROSS_KEYS_PAD / 16 + DB_SPLIT_TILES - 1) / DB_SPLIT_TILES, "key splits");
constexpr unsigned DB_SPIN_LIMIT = 1u << 22;

template <int NT, int CW>
struct DbCfg {
    // 8 consumer warps = two TILE GROUPS of four (a GEMV tile is multiplied, reduced and stored by one group while the other
    // works on the next tile: two dependent chains in flight instead of one), plus one producer warpgroup (one active
    // thread); setmaxnreg hands the producer group's registers to the consumers
    static_assert(CW == 8, "two tile groups of four warps");
    static constexpr int PW = 4;                                       // producer warps (warp 0 lane 0 does the work)
    static constexpr bool SETREG = true;
    static constexpr int CONS = CW * 32, THREADS = CONS + PW * 32, ROWS = NT * 8;
    static constexpr int GW = 4, GT = GW * 32;                         // warps / threads of a tile group
    static constexpr int ARRIVALS = 4;                                 // warps that release a ring slot
    static constexpr int IPR = (320 + CONS - 1) / CONS;               // 16-byte items (4 columns) per row and thread, d <= 1280
    static constexpr int PASSES = (ROWS * 16 + GT - 1) / GT;          // epilogue passes: one (row, output) per group thread and pass
    static constexpr int RED_FLOATS = 2 * GW * NT * 128;              // one reduction buffer per tile group
};
// scratch behind the activation rows, in floats: red | sp | sq | stat | rowstat | ints (spos[8] stok[40]) | barriers | stage descriptors | seen[4]
__host__ __device__ constexpr size_t db_scratch_bytes(int nt, int cw) {
    return ((size_t)2 * 4 * nt * 128 + 8 * DB_SPLIT_KEYS + 8 * 64 + 64 + 2 * DB_MAX_ROWS + 48) * 4 + 2 * DB_MAX_SLOTS * 8 + 2 * 64 + DB_MAX_SLOTS * 4;
}

__device__ __forceinline__ void db_mma(float (&d)[4], const uint4& lo, const uint4& hi, const uint4& xb) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(lo.x), "r"(hi.x), "r"(lo.y), "r"(hi.y), "r"(xb.x), "r"(xb.y));
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(lo.z), "r"(hi.z), "r"(lo.w), "r"(hi.w), "r"(xb.z), "r"(xb.w));
}
template <int CONS>
__device__ __forceinline__ void csync() { asm volatile("bar.sync 1, %0;" ::"n"(CONS) : "memory"); }

__device__ __forceinline__ void db_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void db_cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void db_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// A protocol bug traps (-> launch error) instead of hanging the GPU box.  One instruction: the kernel is instruction-fetch
// bound (every stage's code runs once per layer from a cold instruction cache), so nothing cold may sit between hot code.
// Before the trap the wait that gave up leaves {where, CTA, thread, grid} in a host-mapped word (db_fault_word) for the error log.
__device__ unsigned long long* g_db_fault = nullptr;
// Every CTA also keeps its progress (stage index + 1 of its consumers, of its producer) in g_db_progress[decode lane][CTA]; the
// wait that gives up copies the table behind the fault word, so the log shows which CTA the others were waiting for.
__device__ unsigned g_db_progress[DB_PROGRESS_WORDS];
__device__ __forceinline__ void db_progress(int lane_id, int half, int it) {
    reinterpret_cast<volatile unsigned short*>(g_db_progress + (lane_id * DB_PROGRESS_LD + blockIdx.x) * 2)[half] = (unsigned short)(it + 1);
}
// The consumers' version is a real function (the table copy inlined at a dozen wait sites cost 2 % of the step: the kernel is
// instruction-fetch bound); the producer's stays one store + trap, because a call reachable from BOTH role branches makes ptxas
// drop the setmaxnreg register hand-over.
__device__ __forceinline__ void db_fault_store(int where) {
    *(volatile unsigned long long*)g_db_fault = 1ull << 63 | (unsigned long long)where << 48 | (unsigned long long)gridDim.x << 32 |
                                                (unsigned long long)blockIdx.x << 16 | threadIdx.x;
    __threadfence_system();
}
__device__ __noinline__ void db_timeout(int where) {
    if (g_db_fault) {
        volatile unsigned* dst = reinterpret_cast<volatile unsigned*>(g_db_fault + 2);
        for (int i = 0; i < DB_PROGRESS_WORDS; ++i) dst[i] = reinterpret_cast<volatile unsigned*>(g_db_progress)[i];
        db_fault_store(where);
    }
    asm volatile("trap;");
}
__device__ __forceinline__ void db_timeout_producer(int where) {
    if (g_db_fault) db_fault_store(where);
    asm volatile("trap;");
}
template <bool PRODUCER = false>
__device__ __forceinline__ void db_wait(uint64_t* bar, uint32_t parity, int where) {         // bounded mbarrier wait
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > (1u << 21)) { if (PRODUCER) db_timeout_producer(where); else db_timeout(where); }
}

__device__ __forceinline__ unsigned long long db_gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
template <bool ON>
struct DbDbgT {
    unsigned long long* buf; bool on;
    int sub;                                           // sub-marks of ONE probed stage (sub >= 0): indices 600 .. 639
    __device__ __forceinline__ void mark(int idx) { if (ON && on && idx < DB_DBG_LD) buf[idx] = db_gtimer(); }
    // (clock64: a %globaltimer read costs ~0.3 us, far too much for marks a few hundred cycles apart)
    __device__ __forceinline__ void smark(int k) { if (ON && on && sub >= 0 && k < 40) buf[600 + k] = (unsigned long long)clock64(); }
    __device__ __forceinline__ void set_sub(int v) { if (ON) sub = v; }
};

// ---- LL words -----------------------------------------------------------------------------------------------------
typedef unsigned long long u64;
__device__ __forceinline__ void ll_st(uint2* p, uint32_t payload, uint32_t epoch) {
    const u64 v = ((u64)epoch << 32) | payload;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 ll_ld1(const uint2* p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ll_ld2(const uint2* p, u64& a, u64& b) {           // p is 16-byte aligned
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ bool ll_good(u64 v, uint32_t epoch) { return (uint32_t)(v >> 32) == epoch; }
__device__ __forceinline__ void ll_pause(unsigned& spins, int where) {
    if (++spins > DB_SPIN_LIMIT) db_timeout(where);
    if (spins > 48) __nanosleep(64);                  // the first polls spin: __nanosleep's granularity is coarser than an L2 round trip
}
// A wait for an LL word that has used half of its budget notes {CTA, thread, where | address | last value | epoch expected} in a
// table (every long waiter does, the root cause among them); the wait that finally gives up copies the table behind the
// progress marks.
__device__ unsigned long long g_db_notes[DB_NOTE_ROWS * 4];
__device__ unsigned g_db_n_notes;
__device__ __noinline__ void db_note_wait(int where, const uint2* p, u64 v, uint32_t epoch) {
    const unsigned i = atomicAdd(&g_db_n_notes, 1u);
    if (i < (unsigned)DB_NOTE_ROWS) {
        volatile unsigned long long* x = g_db_notes + i * 4;
        x[0] = (unsigned long long)blockIdx.x | (unsigned long long)threadIdx.x << 16 | (unsigned long long)where << 32 | (unsigned long long)gridDim.x << 48;
        x[1] = (unsigned long long)p; x[2] = v; x[3] = epoch;
    }
}
__device__ __noinline__ void db_timeout_word(int where, const uint2* p, u64 v, uint32_t epoch) {
    if (g_db_fault) {
        volatile unsigned long long* x = g_db_fault + 2 + DB_PROGRESS_WORDS / 2;
        const unsigned n = min(*(volatile unsigned*)&g_db_n_notes, (unsigned)DB_NOTE_ROWS);
        for (unsigned i = 0; i < n * 4; ++i) x[1 + i] = ((volatile unsigned long long*)g_db_notes)[i];
        x[0] = n;
    }
    db_timeout(where);
}
__device__ __forceinline__ uint32_t ll_wait_word(const uint2* p, uint32_t epoch, int where) {
    u64 v;
    unsigned spins = 0;
    while (!ll_good(v = ll_ld1(p), epoch)) {
        if (++spins > DB_SPIN_LIMIT) db_timeout_word(where, p, v, epoch);
        if (spins == DB_SPIN_LIMIT / 2) db_note_wait(where, p, v, epoch);
        if (spins > 48) __nanosleep(64);
    }
    return (uint32_t)v;
}
// wait for the last word of every `stride`-word group in [word0, word0 + n_words) of rows [row_lo, row_hi) of an LL matrix
template <int CONS>
__device__ __forceinline__ void ll_wait_sentinels(const uint2* buf, uint32_t epoch, long row_words, int word0, int n_words, int stride,
                                                  int row_lo, int row_hi, int tid, int where) {
    const int per_row = n_words / stride, n_sent = (row_hi - row_lo) * per_row;
    for (int k = tid; k < n_sent; k += CONS) {
        const int r = row_lo + k / per_row, c = word0 + (k % per_row) * stride + stride - 1;
        ll_wait_word(buf + (long)r * row_words + c, epoch, where);
    }
    csync<CONS>();
}
// Copies a region of a bf16x2 LL matrix - rows [0, n_rows), words [word0, word0 + 2 * items) of rows `row_words` apart - into
// the shared-memory activation rows (row r, columns from 0).  A ROLLED loop with four 16-byte items in flight per thread: the
// kernel is instruction-fetch bound, so the body has to stay a few cache lines long (an unrolled version with 13 loads in
// flight measured 1.8x slower although it needs a third of the round trips).
template <int CONS>
__device__ __forceinline__ void ll_copy_region(const uint2* __restrict__ src, uint32_t epoch, int n_rows, long row_words, int word0, int items,
                                               bf16* xs, int ldx, int tid, int where) {
    int c = tid, r = 0;
    while (c >= items) { c -= items; ++r; }
#pragma unroll 1
    while (r < n_rows) {
        int rc[4];                                     // (row << 16) | item of every load of the batch, -1 = past the end
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            rc[u] = r < n_rows ? (r << 16) | c : -1;
            c += CONS;
            while (c >= items) { c -= items; ++r; }
        }
        u64 w[4][2];
        unsigned spins = 0;
        bool ok;
        do {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int rr = rc[u] < 0 ? 0 : rc[u] >> 16, cc = rc[u] < 0 ? 0 : rc[u] & 0xffff;
                ll_ld2(src + (long)rr * row_words + word0 + 2 * cc, w[u][0], w[u][1]);
            }
            ok = true;
#pragma unroll
            for (int u = 0; u < 4; ++u) ok = ok & ll_good(w[u][0], epoch) & ll_good(w[u][1], epoch);
            if (!ok) ll_pause(spins, where);
        } while (!ok);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (rc[u] >= 0) reinterpret_cast<uint2*>(xs + (long)(rc[u] >> 16) * ldx)[rc[u] & 0xffff] = make_uint2((uint32_t)w[u][0], (uint32_t)w[u][1]);
    }
}

// The model descriptor lives in constant memory: its pointers are read at every stage.
__constant__ DbModel c_db;

// ---- the CTA's position in the slot ring; producer and consumers advance identical copies -----------------
struct DbRing {
    int slot; uint32_t phase;
    int n;
    int use;                                           // running index of the slot use (consumers only)
    __device__ __forceinline__ void advance() { ++use; if (++slot == n) { slot = 0; phase ^= 1; } }
};
__device__ __forceinline__ void db_range(int n_tiles, int vcta, int nctas, int& u0, int& u1) {      // tiles [u0, u1) of a GEMV stage
    u0 = vcta * n_tiles / nctas; u1 = (vcta + 1) * n_tiles / nctas;
}
// blocks of the slot that starts at K block c0: at most 40, and never across a K-chunk boundary of the activations
__device__ __forceinline__ int db_slot_blocks(int c0, int n_kc, int chunk_kc) {
    const int chunk_end = min(n_kc, (c0 / chunk_kc + 1) * chunk_kc);
    return min(DB_SLOT_BLOCKS, chunk_end - c0);
}
// what a GEMV tile's epilogue needs to know about its stage: filled by ONE thread per stage (the look-ups are chains of selects
// over the stage kind; evaluated per tile they were most of the epilogue's instructions), read by everybody from shared memory
static_assert(true, "");
struct DbStageSm {
    const float* bias; const uint2* res_ll; uint2* out_ll; uint2* out_llb; long ld_out;
    int n_valid, epi, res_mode; uint32_t res_epoch;
};
struct DbSmem {
    uint8_t* ring; bf16* xs; float* red; float* sp; float* sq; float* stat; float* rowstat; int* spos; int* stok;
    uint64_t* full; uint64_t* empty; DbStageSm* desc;
    volatile int* seen;                                // [slot] latest use of the slot whose data some consumer has seen arrive (DbRing::use)
    int ldx;
};
__device__ __forceinline__ void db_seen_wait(const DbSmem& sm, const DbRing& ring) {
    unsigned spins = 0;
    while (sm.seen[ring.slot] < ring.use - ring.n) if (++spins > (1u << 23)) db_timeout(26);
}
// row r of the step -> index of its token history / slot-table row / physical KV slot
__device__ __forceinline__ int db_trow(const DbArgs& a, int r) { const int w = r / a.nbw; return w * a.slot_stride + (r - w * a.nbw); }

// ---- consumer: one GEMV stage -------------------------------------------------------------------------------------
// A unit = 16 output rows x all rows of the step; its K/32 weight blocks arrive <= 40 per slot and warp w multiplies blocks
// w, w + CW, ... of the slot against every n-tile.  Every thread owns (row, output) pairs: the additive term (bias,
// residual) is fetched BEFORE the MMA loop so that its latency hides behind the weight stream.
enum { DB_EPI_LL_F32 = 0, DB_EPI_LL_GELU_BF16 = 1, DB_EPI_LOGITS = 2 };
enum { DB_RES_NONE = 0, DB_RES_LL = 1, DB_RES_EMBED = 2, DB_RES_XIN = 3 };
enum { DBS_QKV = 0, DBS_SA, DBS_OUT, DBS_CQ, DBS_CA, DBS_CO, DBS_M1, DBS_M2, DBS_VOCAB };
// The stage descriptor is three registers (stage, layer, epoch); everything else - bias, residual source, output buffers - is
// looked up in constant memory (the model descriptor, the kernel arguments) at the point of use, so that nothing but the
// accumulators is live across the MMA loop (the kernel has 224 registers per consumer thread to work with).
struct DbGemv {
    int st, l;
    int n_tiles, n_kc, vcta;
    uint32_t ep;                 // this layer's epoch; the layer below wrote ep - 1
    int chunk_kc;                // K blocks of the activations staged at a time (MLP2), else n_kc
    __device__ __forceinline__ int epi() const { return st == DBS_M1 ? DB_EPI_LL_GELU_BF16 : st == DBS_VOCAB ? DB_EPI_LOGITS : DB_EPI_LL_F32; }
    __device__ __forceinline__ const float* bias() const {
        const DbLayer& L = c_db.layers[l];
        return st == DBS_QKV ? L.qkv_b : st == DBS_OUT ? L.attn_out_b : st == DBS_CQ ? L.cross_q_b : st == DBS_CO ? L.cross_out_b
             : st == DBS_M1 ? L.mlp1_b : st == DBS_M2 ? L.mlp2_b : nullptr;
    }
    __device__ __forceinline__ int res_mode(const DbArgs& a) const {
        return st == DBS_OUT ? (l == 0 ? (a.x_in ? DB_RES_XIN : DB_RES_EMBED) : DB_RES_LL) : (st == DBS_CO || st == DBS_M2) ? DB_RES_LL : DB_RES_NONE;
    }
    __device__ __forceinline__ const uint2* res_ll(const DbArgs& a) const { return st == DBS_OUT ? a.ll_x3 : st == DBS_CO ? a.ll_x1 : a.ll_x2; }
    __device__ __forceinline__ uint32_t res_epoch() const { return st == DBS_OUT ? ep - 1u : ep; }
    __device__ __forceinline__ uint2* out_ll(const DbArgs& a) const {
        return st == DBS_QKV ? a.ll_qkv : st == DBS_OUT ? a.ll_x1 : st == DBS_CQ ? a.ll_q : st == DBS_CO ? a.ll_x2 : st == DBS_M1 ? a.ll_hid : a.ll_x3;
    }
    // DB_EPI_LL_F32 stages whose outputs feed a LayerNorm publish them again as bf16x2 LL words
    __device__ __forceinline__ uint2* out_llb(const DbArgs& a) const { return st == DBS_OUT ? a.ll_x1b : st == DBS_CO ? a.ll_x2b : st == DBS_M2 ? a.ll_x3b : nullptr; }
    __device__ __forceinline__ long ld_out(const DbArgs& a) const { return st == DBS_QKV ? 3L * c_db.d : st == DBS_M1 ? 4L * c_db.d : st == DBS_VOCAB ? a.ld_logits : (long)c_db.d; }
    __device__ __forceinline__ int n_valid() const { return st == DBS_QKV ? 3 * c_db.d : st == DBS_M1 ? 4 * c_db.d : st == DBS_VOCAB ? c_db.V : c_db.d; }
};

template <int CONS>
__device__ __forceinline__ void gsync(int group) { asm volatile("bar.sync %0, %1;" ::"r"(2 + group), "n"(CONS) : "memory"); }

template <int NT, int CW, class Dbg>
__device__ __forceinline__ void db_stage_gemv(const DbSmem& sm, DbRing& ring, const DbGemv& g, const DbStageSm& D, const DbArgs& a, int R,
                                              int nctas, int warp, int lane, Dbg& dbg) {
    using C = DbCfg<NT, CW>;
    const DbModel& M = c_db;
    const int gq = lane >> 2, tq = lane & 3, tid = warp * 32 + lane;
    const int group = warp >> 2, gw = warp & 3, gtid = tid & (C::GT - 1);
    const bf16* xrow = sm.xs + (long)gq * sm.ldx + tq * 8;
    const bool chunked = g.chunk_kc < g.n_kc;          // MLP2 with the hidden row staged in K chunks: one group owns every tile
    int u0, u1;
    db_range(g.n_tiles, g.vcta, nctas, u0, u1);
    for (int t = u0; t < u1; ++t) {
        const bool mine = chunked ? group == 0 : ((t - u0) & 1) == group;          // (uniform within a group)
        if (t - u0 == 2) dbg.smark(27);                                            // (probe: top of the group's second tile)
        float add[C::PASSES];
        // seen[] of the tile's first slot, read early (its latency hides behind the residual fetch; the value only grows)
        int sv = mine ? sm.seen[ring.slot] : 0;
        if (mine) {
#pragma unroll
            for (int p = 0; p < C::PASSES; ++p) {
                const int idx = gtid + p * C::GT, r = idx >> 4, n = t * 16 + (idx & 15);
                add[p] = 0.f;
                if (r < R && n < D.n_valid) {
                    const float* bias = D.bias;
                    const int res_mode = D.res_mode;
                    if (bias) add[p] = __ldg(bias + n);
                    if (res_mode == DB_RES_LL) add[p] += __uint_as_float(ll_wait_word(D.res_ll + (long)r * M.d + n, D.res_epoch, 1));
                    else if (res_mode == DB_RES_EMBED) {
                        const int pos = sm.spos[r / a.nbw];
                        add[p] += __bfloat162float(M.tok_emb[(long)a.tokens[db_trow(a, r) * DEC_TOK_LD + pos] * M.d + n]) + __ldg(M.pos_emb + (long)pos * M.d + n);
                    } else if (res_mode == DB_RES_XIN) add[p] += __ldg(a.x_in + (long)r * M.d + n);
                }
            }
        }
        dbg.smark(20 + 4 * (t - u0));
        float acc[NT][2][4];
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) { acc[j][0][e] = 0.f; acc[j][1][e] = 0.f; }
        for (int c0 = 0, nblk = 0; c0 < g.n_kc; c0 += nblk) {
            // first block of the slot within the staged columns, blocks of the slot (no divisions on the common path)
            const int cin = chunked ? c0 % g.chunk_kc : c0;
            nblk = chunked ? db_slot_blocks(c0, g.n_kc, g.chunk_kc) : min(DB_SLOT_BLOCKS, g.n_kc - c0);
            if (g.st == DBS_M2 && cin == 0 && !(!chunked && t > u0)) {  // a new K chunk of the MLP hidden row (one chunk: staged once)
                csync<C::CONS>();                                       // everybody is done with the previous chunk
                const int word0 = c0 * 16, n_words = min(g.chunk_kc, g.n_kc - c0) * 16;
                ll_wait_sentinels<C::CONS>(a.ll_hid, g.ep, 2L * M.d, word0, n_words, 8, R - 1, R, tid, 13);
                ll_copy_region<C::CONS>(a.ll_hid, g.ep, R, 2L * M.d, word0, n_words / 2, sm.xs, sm.ldx, tid, 3);
                csync<C::CONS>();
            }
            if (mine) {
                const uint4* sl = reinterpret_cast<const uint4*>(sm.ring + (size_t)ring.slot * DB_SLOT) + lane;
                const bf16* xk = xrow + cin * 32;
                // A wait knows only the PARITY of the use it waits for, and the slot's previous use may have been the other
                // group's (skipped here without a wait): were that one still in flight, this wait would pass at once on stale
                // bytes.  So first see that some consumer has watched the previous use arrive (db_seen_wait).
                if (sv < ring.use - ring.n) db_seen_wait(sm, ring);
                db_wait(&sm.full[ring.slot], ring.phase, 20);
                if (gw == 0 && lane == 0) sm.seen[ring.slot] = ring.use;
                sv = sm.seen[ring.slot + 1 == ring.n ? 0 : ring.slot + 1];         // the next slot of the tile
#pragma unroll
                for (int q0 = 0; q0 < DB_SLOT_BLOCKS / C::GW; q0 += 5) {           // two batches of five blocks per warp
                    uint4 lo[5], hi[5];
#pragma unroll
                    for (int q = 0; q < 5; ++q) {
                        const int blk = gw + C::GW * (q0 + q);
                        if (blk < nblk) { lo[q] = sl[blk * 64]; hi[q] = sl[blk * 64 + 32]; }
                    }
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        if (NT <= 2) {                                  // activation fragments fetched ahead of their MMAs (register budget)
                            uint4 xb[5];
#pragma unroll
                            for (int q = 0; q < 5; ++q) {
                                const int blk = gw + C::GW * (q0 + q);
                                if (blk < nblk) xb[q] = *reinterpret_cast<const uint4*>(xk + (long)j * 8 * sm.ldx + blk * 32);
                            }
#pragma unroll
                            for (int q = 0; q < 5; ++q)
                                if (gw + C::GW * (q0 + q) < nblk) db_mma(acc[j][q & 1], lo[q], hi[q], xb[q]);
                        } else {
#pragma unroll
                            for (int q = 0; q < 5; ++q) {
                                const int blk = gw + C::GW * (q0 + q);
                                if (blk < nblk) db_mma(acc[j][q & 1], lo[q], hi[q], *reinterpret_cast<const uint4*>(xk + (long)j * 8 * sm.ldx + blk * 32));
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.empty[ring.slot]);
            }
            // (The other group's slots are skipped without a wait: a full barrier cannot complete a second phase before its
            // owner group has released the slot, so a group's waits on its OWN slots never alias; waiting on the other group's
            // slots would, because the producer refills them as soon as their owner is done.)
            ring.advance();                                             // both groups track every slot of the CTA
        }
        if (!mine) { dbg.smark(21 + 4 * (t - u0)); continue; }
        dbg.smark(21 + 4 * (t - u0));
        float* red = sm.red + group * (C::GW * NT * 128);
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            float* r = red + gw * (NT * 128) + j * 128;
            r[gq * 8 + tq * 2] = acc[j][0][0] + acc[j][1][0]; r[gq * 8 + tq * 2 + 1] = acc[j][0][1] + acc[j][1][1];
            r[(gq + 8) * 8 + tq * 2] = acc[j][0][2] + acc[j][1][2]; r[(gq + 8) * 8 + tq * 2 + 1] = acc[j][0][3] + acc[j][1][3];
        }
        gsync<C::GT>(group);
        dbg.smark(22 + 4 * (t - u0));
#pragma unroll
        for (int p = 0; p < C::PASSES; ++p) {
            const int idx = gtid + p * C::GT, r = idx >> 4, o = idx & 15, n = t * 16 + o;
            const bool owner = r < R && n < D.n_valid;
            const long ld_out = D.ld_out;
            const int epi = D.epi;
            const float* rr = red + (r >> 3) * 128 + o * 8 + (r & 7);
            float v = add[p];
            if (idx < C::ROWS * 16) v += (rr[0] + rr[NT * 128]) + (rr[2 * NT * 128] + rr[3 * NT * 128]);
            if (epi == DB_EPI_LL_GELU_BF16) {
                v = gelu_erf(v);
                const float nxt = __shfl_down_sync(0xffffffffu, v, 1);          // lanes are (row, o): o + 1 is the next lane
                if (owner && !(o & 1)) ll_st(D.out_ll + ((long)r * ld_out + n) / 2, pack_bf16(v, nxt), g.ep);
            } else if (epi == DB_EPI_LL_F32) {
                if (owner) ll_st(D.out_ll + (long)r * ld_out + n, __float_as_uint(v), g.ep);
                uint2* out_llb = D.out_llb;
                if (out_llb) {                                        // (stage uniform)
                    const float nxt = __shfl_down_sync(0xffffffffu, v, 1);
                    if (owner && !(o & 1)) ll_st(out_llb + ((long)r * ld_out + n) / 2, pack_bf16(v, nxt), g.ep);
                }
            } else if (owner) {
                __stcg(a.logits + (long)r * ld_out + n, v);
            }
        }
        dbg.smark(23 + 4 * (t - u0));
        gsync<C::GT>(group);                            // the group's next tile reuses its reduction buffer
    }
}

// ---- prologue: LayerNorm of the residual stream into the bf16 activation rows -------------------------------------------
// LL rows (every LayerNorm but the step's first): every thread owns 16-byte items (4 columns) c = tid + CONS * i of every row.
// Pass 1 fetches them as bf16x2 LL words, DB_G rows in flight, adds them into the row's sum and sum of squares in fp32 and parks
// them in the row as the bf16 pairs they are; row statistics are combined across warps through shared memory; pass 2 normalises
// the thread's own items in place.  The step's first LayerNorm (token + position embeddings, or x_in) is done a row per warp
// with the fp32 sums in registers: rounding them to bf16 before the normalisation costs the soft-logit test cases their token
// parity.  gamma | beta arrive through the ring.
enum { DB_PRO_EMBED = 0, DB_PRO_LL = 1 };
constexpr int DB_G = 5;                               // rows whose loads are in flight together
constexpr int DB_LPR = 10;                            // 16-byte items per lane and row in the row-per-warp form (d <= 1280)
template <int NT, int CW, class Dbg>
__device__ __forceinline__ void db_prologue_ln(const DbSmem& sm, DbRing& ring, const DbArgs& a, int R, int mode, const uint2* __restrict__ x_ll,
                                               uint32_t epoch, int warp, int lane, Dbg& dbg) {
    using C = DbCfg<NT, CW>;
    const DbModel& M = c_db;
    const int d = M.d, tid = warp * 32 + lane, items = d >> 2, row_words = d >> 1;
    float* part = sm.red;                                                  // [CW][ROWS][sum, sum of squares]
    csync<C::CONS>();                                                      // the previous stage is done with xs / red
    if (mode == DB_PRO_EMBED) {
        if (!a.x_in) {
            if (tid < R) sm.stok[tid] = a.tokens[db_trow(a, tid) * DEC_TOK_LD + sm.spos[tid / a.nbw]];
            csync<C::CONS>();
        }
        db_wait(&sm.full[ring.slot], ring.phase, 22);                          // gamma | beta: one slot ahead of the stage's tiles
        if (tid == 0) sm.seen[ring.slot] = ring.use;
        const float4* gsl = reinterpret_cast<const float4*>(sm.ring + (size_t)ring.slot * DB_SLOT);
        const float inv_d = __fdividef(1.f, (float)d);
#pragma unroll 1
        for (int r = warp; r < R; r += CW) {
            float x[DB_LPR][4];
            if (a.x_in) {
                const float4* xp = reinterpret_cast<const float4*>(a.x_in + (long)r * d);
#pragma unroll
                for (int i = 0; i < DB_LPR; ++i) {
                    const int c = lane + 32 * i;
                    const float4 v = c < items ? __ldg(xp + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    x[i][0] = v.x; x[i][1] = v.y; x[i][2] = v.z; x[i][3] = v.w;
                }
            } else {
                uint2 tv[DB_LPR];
                const uint2* tp = reinterpret_cast<const uint2*>(M.tok_emb + (long)sm.stok[r] * d);
                const float4* pp = reinterpret_cast<const float4*>(M.pos_emb + (long)sm.spos[r / a.nbw] * d);
#pragma unroll
                for (int i = 0; i < DB_LPR; ++i) {
                    const int c = lane + 32 * i;
                    const bool v = c < items;
                    tv[i] = v ? __ldg(tp + c) : make_uint2(0u, 0u);
                    const float4 pv = v ? __ldg(pp + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    x[i][0] = pv.x; x[i][1] = pv.y; x[i][2] = pv.z; x[i][3] = pv.w;
                }
#pragma unroll
                for (int i = 0; i < DB_LPR; ++i) { x[i][0] += bf16lo(tv[i].x); x[i][1] += bf16hi(tv[i].x); x[i][2] += bf16lo(tv[i].y); x[i][3] += bf16hi(tv[i].y); }
            }
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int i = 0; i < DB_LPR; ++i) {                             // items past the row are zeros: they add nothing
                s1 += (x[i][0] + x[i][1]) + (x[i][2] + x[i][3]);
                s2 = fmaf(x[i][0], x[i][0], fmaf(x[i][1], x[i][1], fmaf(x[i][2], x[i][2], fmaf(x[i][3], x[i][3], s2))));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
            const float mean = s1 * inv_d, rstd = rsqrtf(fmaxf(s2 * inv_d - mean * mean, 0.f) + 1e-5f);
            uint2* row = reinterpret_cast<uint2*>(sm.xs + (long)r * sm.ldx);
#pragma unroll
            for (int i = 0; i < DB_LPR; ++i) {
                const int c = lane + 32 * i;
                if (c < items) {
                    const float4 ga = gsl[c], be = gsl[items + c];
                    row[c] = make_uint2(pack_bf16((x[i][0] - mean) * rstd * ga.x + be.x, (x[i][1] - mean) * rstd * ga.y + be.y),
                                        pack_bf16((x[i][2] - mean) * rstd * ga.z + be.z, (x[i][3] - mean) * rstd * ga.w + be.w));
                }
            }
        }
        csync<C::CONS>();                                                  // every warp is done with gamma | beta
        if (warp < C::ARRIVALS && lane == 0) mbar_arrive(&sm.empty[ring.slot]);
        ring.advance();
        return;
    }
    bool v[C::IPR];
    int cc[C::IPR];
#pragma unroll
    for (int i = 0; i < C::IPR; ++i) { v[i] = tid + i * C::CONS < items; cc[i] = v[i] ? tid + i * C::CONS : 0; }   // out of range: re-read item 0
    {
        ll_wait_sentinels<C::CONS>(x_ll, epoch, row_words, 0, row_words, 8, R - 1, R, tid, 12);   // written by 16-column GEMV tiles (8 words)
#pragma unroll 1
        for (int r0 = 0; r0 < R; r0 += DB_G) {
            u64 w[DB_G][C::IPR][2];
            const uint2* pr[DB_G];
#pragma unroll
            for (int q = 0; q < DB_G; ++q) pr[q] = x_ll + (long)min(r0 + q, R - 1) * row_words;      // rows past R re-read the last row
            unsigned spins = 0;
            bool ok;
            do {
#pragma unroll
                for (int q = 0; q < DB_G; ++q)
#pragma unroll
                    for (int i = 0; i < C::IPR; ++i) ll_ld2(pr[q] + 2 * cc[i], w[q][i][0], w[q][i][1]);
                ok = true;
#pragma unroll
                for (int q = 0; q < DB_G; ++q)
#pragma unroll
                    for (int i = 0; i < C::IPR; ++i) ok = ok & ll_good(w[q][i][0], epoch) & ll_good(w[q][i][1], epoch);
                if (!ok) ll_pause(spins, 2);
            } while (!ok);
            float s1[DB_G], s2[DB_G];
#pragma unroll
            for (int q = 0; q < DB_G; ++q) {
                uint2* dst = reinterpret_cast<uint2*>(sm.xs + (long)min(r0 + q, R - 1) * sm.ldx);
                s1[q] = 0.f; s2[q] = 0.f;
#pragma unroll
                for (int i = 0; i < C::IPR; ++i)
                    if (v[i]) {
                        const uint32_t p0 = (uint32_t)w[q][i][0], p1 = (uint32_t)w[q][i][1];
                        dst[cc[i]] = make_uint2(p0, p1);
                        const float x0 = bf16lo(p0), x1 = bf16hi(p0), x2 = bf16lo(p1), x3 = bf16hi(p1);
                        s1[q] += (x0 + x1) + (x2 + x3);
                        s2[q] = fmaf(x0, x0, fmaf(x1, x1, fmaf(x2, x2, fmaf(x3, x3, s2[q]))));
                    }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int q = 0; q < DB_G; ++q) { s1[q] += __shfl_xor_sync(0xffffffffu, s1[q], o); s2[q] += __shfl_xor_sync(0xffffffffu, s2[q], o); }
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < DB_G; ++q)
                    if (r0 + q < R) *reinterpret_cast<float2*>(part + (warp * C::ROWS + r0 + q) * 2) = make_float2(s1[q], s2[q]);
            }
        }
    }
    csync<C::CONS>();
    if (tid < R) {                                                         // one thread per row: (mean, rstd)
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int w = 0; w < CW; ++w) { const float2 p = *reinterpret_cast<const float2*>(part + (w * C::ROWS + tid) * 2); s1 += p.x; s2 += p.y; }
        const float inv_d = __fdividef(1.f, (float)d), mean = s1 * inv_d;
        *reinterpret_cast<float2*>(sm.rowstat + 2 * tid) = make_float2(mean, rsqrtf(fmaxf(s2 * inv_d - mean * mean, 0.f) + 1e-5f));
    }
    db_wait(&sm.full[ring.slot], ring.phase, 23);                              // gamma | beta: one slot ahead of the stage's tiles
    if (tid == 0) sm.seen[ring.slot] = ring.use;
    csync<C::CONS>();
    {
        float4 ga[C::IPR], be[C::IPR];
        const float4* gsl = reinterpret_cast<const float4*>(sm.ring + (size_t)ring.slot * DB_SLOT);
#pragma unroll
        for (int i = 0; i < C::IPR; ++i) { ga[i] = gsl[cc[i]]; be[i] = gsl[items + cc[i]]; }
#pragma unroll 1
        for (int r0 = 0; r0 < R; r0 += DB_G) {
            uint2 xv[DB_G][C::IPR];
            float2 ms[DB_G];
#pragma unroll
            for (int q = 0; q < DB_G; ++q) {
                const int r = min(r0 + q, R - 1);
                const uint2* row = reinterpret_cast<const uint2*>(sm.xs + (long)r * sm.ldx);
#pragma unroll
                for (int i = 0; i < C::IPR; ++i) xv[q][i] = row[cc[i]];
                ms[q] = *reinterpret_cast<const float2*>(sm.rowstat + 2 * r);
            }
#pragma unroll
            for (int q = 0; q < DB_G; ++q) {
                if (r0 + q < R) {
                    uint2* row = reinterpret_cast<uint2*>(sm.xs + (long)(r0 + q) * sm.ldx);
#pragma unroll
                    for (int i = 0; i < C::IPR; ++i)
                        if (v[i]) {
                            const float mean = ms[q].x, rstd = ms[q].y;
                            row[cc[i]] = make_uint2(pack_bf16((bf16lo(xv[q][i].x) - mean) * rstd * ga[i].x + be[i].x, (bf16hi(xv[q][i].x) - mean) * rstd * ga[i].y + be[i].y),
                                                    pack_bf16((bf16lo(xv[q][i].y) - mean) * rstd * ga[i].z + be[i].z, (bf16hi(xv[q][i].y) - mean) * rstd * ga[i].w + be[i].w));
                        }
                }
            }
        }
    }
    csync<C::CONS>();                                                      // every warp is done with gamma | beta
    if (warp < C::ARRIVALS && lane == 0) mbar_arrive(&sm.empty[ring.slot]);
    ring.advance();
}

// ---- self-attention of the new token, unit = (row, head) ------------------------------------------------------------------
// The cached K | V rows of the first `cap` positions are copied into shared memory (256 bytes per key, over the idle xs
// rows) BEFORE the unit waits for this step's q | k | v, so only the arithmetic is left on the critical path: thread j
// owns keys j, j + CONS, ... and reads its K row in 16-byte chunks rotated by j (conflict free without padding).
// Positions >= cap (long segments only) are read from global memory on demand.
template <int NT, int CW>
__device__ __forceinline__ void db_stage_self_attn(const DbSmem& sm, const DbArgs& a, int R, bf16* cache_k, bf16* cache_v, uint32_t ep,
                                                   int cta, int nctas, int rot, int warp, int lane) {
    using C = DbCfg<NT, CW>;
    const DbModel& M = c_db;
    const int d = M.d, H = M.H, tid = warp * 32 + lane, cap = a.sa_cap;
    if ((cta + rot) % nctas >= R * H) csync<C::CONS>();    // every stage passes a CTA-wide barrier, units or not (see the stage loop)
    for (int u = (cta + rot) % nctas; u < R * H; u += nctas) {
        const int r = u / H, h = u - r * H, trow = db_trow(a, r), pos = sm.spos[r / a.nbw];
        float* ss = sm.sp;                          // [<= 449] scores
        float* sqv = sm.sq;                         // q fp32 [64]
        bf16* sknew = reinterpret_cast<bf16*>(sm.sq + 64); bf16* svnew = sknew + 64;       // new K | V row when pos >= cap
        bf16* kv = sm.xs;                           // [cap][K 64 | V 64] bf16
        csync<C::CONS>();
        const int* tab = a.table + trow * 448;
        const int n_sm = min(pos, cap);
#pragma unroll 1
        for (int j = tid; j < n_sm; j += C::CONS) {
            const long row = ((long)tab[j] * 448 + j) * d + h * 64;
            const uint4* kp = reinterpret_cast<const uint4*>(cache_k + row);
            const uint4* vp = reinterpret_cast<const uint4*>(cache_v + row);
#pragma unroll
            for (int i = 0; i < 8; ++i) { db_cp_async16(kv + (long)j * 128 + i * 8, kp + i); db_cp_async16(kv + (long)j * 128 + 64 + i * 8, vp + i); }
        }
        bf16* knew = pos < cap ? kv + (long)pos * 128 : sknew;
        bf16* vnew = pos < cap ? kv + (long)pos * 128 + 64 : svnew;
        // q | k | v of this (row, head): 192 fp32 LL words; word i < 64 is q, then k, then v
#pragma unroll 1
        for (int i = tid; i < 192; i += C::CONS) {
            const int part = i >> 6, c = i & 63;
            const uint32_t wv = ll_wait_word(a.ll_qkv + (long)r * 3 * d + part * d + h * 64 + c, ep, 4);
            const long off = ((long)trow * 448 + pos) * d + h * 64 + c;               // the new row lives in the row's own physical slot
            if (part == 0) sqv[c] = __uint_as_float(wv);
            else if (part == 1) { const bf16 kb = __float2bfloat16(__uint_as_float(wv)); knew[c] = kb; cache_k[off] = kb; }
            else { const bf16 vb = __float2bfloat16(__uint_as_float(wv)); vnew[c] = vb; cache_v[off] = vb; }
        }
        if (h == 0 && tid == 0) a.table[trow * 448 + pos] = trow;
        db_cp_async_wait_all();
        csync<C::CONS>();
        float m = -INFINITY;
#pragma unroll 1
        for (int j = tid; j <= pos; j += C::CONS) {
            const bf16* kp = j < cap ? kv + (long)j * 128 : (j == pos ? sknew : cache_k + ((long)tab[j] * 448 + j) * d + h * 64);
            const int rot_j = j < cap ? j : 0;
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int ch = (i + rot_j) & 7;
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
        csync<C::CONS>();
        m = sm.stat[0];
#pragma unroll
        for (int w = 1; w < CW; ++w) m = fmaxf(m, sm.stat[w]);
        float lsum = 0.f;
#pragma unroll 1
        for (int j = tid; j <= pos; j += C::CONS) { const float p = __expf(ss[j] - m); ss[j] = p; lsum += p; }   // own entries only
        lsum = warp_sum(lsum);
        if (lane == 0) sm.stat[16 + warp] = lsum;
        csync<C::CONS>();
        lsum = 0.f;
#pragma unroll
        for (int w = 0; w < CW; ++w) lsum += sm.stat[16 + w];
        float o0 = 0.f, o1 = 0.f;
#pragma unroll 4
        for (int j = warp; j <= pos; j += CW) {
            const bf16* vp = j < cap ? kv + (long)j * 128 + 64 : (j == pos ? svnew : cache_v + ((long)tab[j] * 448 + j) * d + h * 64);
            const uint32_t vv = *reinterpret_cast<const uint32_t*>(vp + 2 * lane);
            const float p = ss[j];
            o0 = fmaf(p, bf16lo(vv), o0); o1 = fmaf(p, bf16hi(vv), o1);
        }
        float* sred = sm.red;                       // [CW][64]
        *reinterpret_cast<float2*>(sred + warp * 64 + 2 * lane) = make_float2(o0, o1);
        csync<C::CONS>();
        if (tid < 32) {
            float e0 = 0.f, e1 = 0.f;
#pragma unroll
            for (int w = 0; w < CW; ++w) { const float2 p = *reinterpret_cast<const float2*>(sred + w * 64 + 2 * tid); e0 += p.x; e1 += p.y; }
            ll_st(a.ll_att + ((long)r * d + h * 64) / 2 + tid, pack_bf16(__fdividef(e0, lsum), __fdividef(e1, lsum)), ep);
        }
    }
}

// ---- cross-attention, unit = (window, head, key split): K/V of a head are read once for all beams of the window; the group of
// a head's last (shortest) split then merges the head's partials in split order.  The CTA's units alternate between the two
// tile groups, which work CONCURRENTLY, each with its own scratch: a unit is a chain of four barriers (q staged -> scores ->
// softmax -> P V), so with two or more units per CTA (any launch with fewer CTAs than units: decode lanes) one unit at a time
// left the SM mostly waiting.  Group g keeps q / p of its unit in activation rows 0 .. nbw - 1 at columns DB_CA_COLS * g and
// its scores in sm.sp (g = 0) or behind them in the same rows (g = 1, columns DB_CA_SP1 ..; db_geometry keeps rows that long).
constexpr int DB_CA_COLS = 512, DB_CA_SP1 = 736, DB_CA_MIN_COLS = DB_CA_SP1 + 2 * DB_SPLIT_KEYS;
template <int NT, int CW>
__device__ __forceinline__ void db_stage_cross_attn(const DbSmem& sm, DbRing& ring, const DbArgs& a, uint32_t ep, int cta, int nctas, int rot,
                                                    int warp, int lane) {
    using C = DbCfg<NT, CW>;
    const DbModel& M = c_db;
    const int d = M.d, H = M.H, gq = lane >> 2, tq = lane & 3, nbw = a.nbw;
    const int group = warp >> 2, gw = warp & 3, gtid = gw * 32 + lane;
    constexpr int n_ktiles = CROSS_KEYS_PAD / 16, n_vkc = CROSS_KEYS_PAD / 32;
    bf16* xg = sm.xs + group * DB_CA_COLS;                                       // q (64 columns), then p (224) of the group's unit
    float* spg = group == 0 ? sm.sp : reinterpret_cast<float*>(sm.xs + DB_CA_SP1);      // scores [beam][224]
    const int sp_ld = group == 0 ? DB_SPLIT_KEYS : sm.ldx / 2;                   // (floats)
    float* statg = sm.stat + 32 + group * 16;                                    // [0 .. 7] max, [8 .. 15] sum per beam
    csync<C::CONS>();                                                            // the previous stage's tiles are done with the activation rows
    int j = 0;
    for (int u = (cta + rot) % nctas; u < a.W * H * DB_N_SPLITS; u += nctas, ++j) {
        if ((j & 1) != group) { ring.advance(); ring.advance(); continue; }      // the other group's unit: its K and V slots
        const int wh = u / DB_N_SPLITS, s = u - wh * DB_N_SPLITS, w = wh / H, h = wh - w * H;
        const int t0 = s * DB_SPLIT_TILES, nt = min(DB_SPLIT_TILES, n_ktiles - t0);
        const int nkeys = min(nt * 16, N_AUDIO_CTX - t0 * 16);                  // valid (unpadded) keys of the split
        gsync<C::GT>(group);                                                     // the group's previous unit is done with its scratch
        // q of this (window, head) -> rows 0 .. nbw - 1 (bf16, 64 columns): 16-byte items of 2 fp32 words
#pragma unroll 1
        for (int i = gtid; i < nbw * 32; i += C::GT) {
            const uint2* p = a.ll_q + (long)(w * nbw + (i >> 5)) * d + h * 64 + (i & 31) * 2;
            u64 q0, q1;
            unsigned spins = 0;
            for (;;) {
                ll_ld2(p, q0, q1);
                if (ll_good(q0, ep) && ll_good(q1, ep)) break;
                ll_pause(spins, 5);
            }
            reinterpret_cast<uint32_t*>(xg + (long)(i >> 5) * sm.ldx)[i & 31] = pack_bf16(__uint_as_float((uint32_t)q0), __uint_as_float((uint32_t)q1));
        }
        gsync<C::GT>(group);
        // scores: one slot carries the split's key tiles (2 blocks each); warp gw of the group takes tiles gw, gw + 4, ...
        db_seen_wait(sm, ring);
        db_wait(&sm.full[ring.slot], ring.phase, 24);
        if (gtid == 0) sm.seen[ring.slot] = ring.use;
        {
            const uint4* sl = reinterpret_cast<const uint4*>(sm.ring + (size_t)ring.slot * DB_SLOT) + lane;
            const uint4 xb0 = *reinterpret_cast<const uint4*>(xg + (long)gq * sm.ldx + tq * 8);
            const uint4 xb1 = *reinterpret_cast<const uint4*>(xg + (long)gq * sm.ldx + 32 + tq * 8);
#pragma unroll
            for (int q = 0; q < (DB_SPLIT_TILES + C::GW - 1) / C::GW; ++q) {
                const int tt = gw + C::GW * q;
                if (tt < nt) {
                    float acc[4] = {0.f, 0.f, 0.f, 0.f};
                    db_mma(acc, sl[tt * 128], sl[tt * 128 + 32], xb0);
                    db_mma(acc, sl[tt * 128 + 64], sl[tt * 128 + 96], xb1);
                    float* spb = spg + tt * 16;                                  // (beams >= nbw are rows that do not exist)
                    if (2 * tq < nbw) { spb[(2 * tq) * sp_ld + gq] = acc[0]; spb[(2 * tq) * sp_ld + gq + 8] = acc[2]; }
                    if (2 * tq + 1 < nbw) { spb[(2 * tq + 1) * sp_ld + gq] = acc[1]; spb[(2 * tq + 1) * sp_ld + gq + 8] = acc[3]; }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[ring.slot]);                       // every warp of the group releases what it has read
        ring.advance();
        gsync<C::GT>(group);                                                     // every score is written
        // partial softmax per beam over the split's valid keys; p (bf16) becomes the B operand of P V
        for (int b = gw; b < nbw; b += C::GW) {
            float sv[DB_SPLIT_KEYS / 32];
            float m = -INFINITY;
#pragma unroll
            for (int i = 0; i < DB_SPLIT_KEYS / 32; ++i) { const int jj = lane + 32 * i; sv[i] = jj < nkeys ? spg[b * sp_ld + jj] : -INFINITY; m = fmaxf(m, sv[i]); }
            m = warp_max(m);
            float lsum = 0.f;
#pragma unroll
            for (int i = 0; i < DB_SPLIT_KEYS / 32; ++i) {
                const float p = __expf(sv[i] - m);                              // exp(-inf) = 0 for the padded keys
                lsum += p;
                xg[(long)b * sm.ldx + lane + 32 * i] = __float2bfloat16(p);
            }
            lsum = warp_sum(lsum);
            if (lane == 0) { statg[b] = m; statg[8 + b] = lsum; }
        }
        gsync<C::GT>(group);
        // o[dim][beam] = V^T[dim][key] p[key][beam]: the slot holds the split's key blocks of the four dim tiles; warp dt owns tile dt
        const int nkc = min(DB_SPLIT_TILES / 2, n_vkc - s * (DB_SPLIT_TILES / 2));
        uint2* part = a.ll_cap + (((long)(w * H + h) * DB_N_SPLITS + s) * 8) * 66;
        db_seen_wait(sm, ring);
        db_wait(&sm.full[ring.slot], ring.phase, 25);
        if (gtid == 0) sm.seen[ring.slot] = ring.use;
        {
            const uint4* sl = reinterpret_cast<const uint4*>(sm.ring + (size_t)ring.slot * DB_SLOT + (size_t)gw * nkc * 1024) + lane;
            float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int kc = 0; kc < DB_SPLIT_TILES / 2; ++kc)
                if (kc < nkc) db_mma(acc[kc & 1], sl[kc * 64], sl[kc * 64 + 32], *reinterpret_cast<const uint4*>(xg + (long)gq * sm.ldx + kc * 32 + tq * 8));
            const int b0 = 2 * tq, dim = gw * 16 + gq;
            if (b0 < nbw) {
                ll_st(part + b0 * 66 + 2 + dim, __float_as_uint(acc[0][0] + acc[1][0]), ep);
                ll_st(part + b0 * 66 + 2 + dim + 8, __float_as_uint(acc[0][2] + acc[1][2]), ep);
            }
            if (b0 + 1 < nbw) {
                ll_st(part + (b0 + 1) * 66 + 2 + dim, __float_as_uint(acc[0][1] + acc[1][1]), ep);
                ll_st(part + (b0 + 1) * 66 + 2 + dim + 8, __float_as_uint(acc[0][3] + acc[1][3]), ep);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[ring.slot]);
        ring.advance();
        if (gtid < nbw) { ll_st(part + gtid * 66, __float_as_uint(statg[gtid]), ep); ll_st(part + gtid * 66 + 1, __float_as_uint(statg[8 + gtid]), ep); }
        if (s == DB_N_SPLITS - 1) {
            const uint2* ph = a.ll_cap + ((long)(w * H + h) * DB_N_SPLITS * 8) * 66;
#pragma unroll 1
            for (int e = gtid; e < nbw * 64; e += C::GT) {
                const int b = e >> 6, c = e & 63;
                u64 rm[DB_N_SPLITS], rl[DB_N_SPLITS], ro[DB_N_SPLITS];         // all loads of a round are in flight together
                unsigned spins = 0;
                bool ok;
                do {
                    ok = true;
#pragma unroll
                    for (int q = 0; q < DB_N_SPLITS; ++q) { const uint2* pq = ph + (q * 8 + b) * 66; ll_ld2(pq, rm[q], rl[q]); ro[q] = ll_ld1(pq + 2 + c); }
#pragma unroll
                    for (int q = 0; q < DB_N_SPLITS; ++q) ok = ok && ll_good(rm[q], ep) && ll_good(rl[q], ep) && ll_good(ro[q], ep);
                    if (!ok) ll_pause(spins, 6);
                } while (!ok);
                float mm = -INFINITY;
#pragma unroll
                for (int q = 0; q < DB_N_SPLITS; ++q) mm = fmaxf(mm, __uint_as_float((uint32_t)rm[q]));
                float ll = 0.f, oo = 0.f;
#pragma unroll
                for (int q = 0; q < DB_N_SPLITS; ++q) {
                    const float wq = __expf(__uint_as_float((uint32_t)rm[q]) - mm);
                    ll = fmaf(__uint_as_float((uint32_t)rl[q]), wq, ll); oo = fmaf(__uint_as_float((uint32_t)ro[q]), wq, oo);
                }
                const float o = __fdividef(oo, ll);
                const float nxt = __shfl_down_sync(0xffffffffu, o, 1);
                if (!(c & 1)) ll_st(a.ll_catt + ((long)(w * nbw + b) * d + h * 64 + c) / 2, pack_bf16(o, nxt), ep);
            }
        }
    }
}

// ---- the stage table: stage `it` of the step (8 per layer + the vocabulary projection) ----------------------------------------
struct DbStageDesc {                   // what the producer needs: the weight matrix of a GEMV stage and who owns which tile
    const bf16* w; int n_tiles, n_kc, vcta;
    const float *ln_g, *ln_b;          // LayerNorm in front of the stage (gamma | beta ride in one slot) or nullptr
};
__device__ __forceinline__ int db_stage_rot(int st, int nctas) {
    // stages with fewer units than CTAs start at different CTAs so that every CTA streams about the same bytes per layer
    return st == DBS_CQ ? nctas / 2 : st == DBS_CA ? nctas / 4 : st == DBS_CO ? (3 * nctas) / 4 : st == DBS_M2 ? nctas / 3 : st == DBS_SA ? nctas / 8 : 0;
}
__device__ __forceinline__ DbStageDesc db_stage_desc(int st, int l, int cta, int nctas) {
    const DbModel& M = c_db;
    const DbLayer& L = M.layers[l];
    const int d = M.d;
    DbStageDesc s{nullptr, d / 16, d >> 5, (cta + db_stage_rot(st, nctas)) % nctas, nullptr, nullptr};
    switch (st) {
    case DBS_QKV: s.w = L.qkv; s.n_tiles = 3 * d / 16; s.ln_g = L.ln1_w; s.ln_b = L.ln1_b; break;
    case DBS_OUT: s.w = L.attn_out; break;
    case DBS_CQ: s.w = L.cross_q; s.ln_g = L.ln2_w; s.ln_b = L.ln2_b; break;
    case DBS_CO: s.w = L.cross_out; break;
    case DBS_M1: s.w = L.mlp1; s.n_tiles = 4 * d / 16; s.ln_g = L.ln3_w; s.ln_b = L.ln3_b; break;
    case DBS_M2: s.w = L.mlp2; s.n_kc = d >> 3; break;
    case DBS_VOCAB: s.w = M.tok_emb_frag; s.n_tiles = M.n_tiles_vocab; s.ln_g = M.ln_w; s.ln_b = M.ln_b; break;
    default: break;
    }
    return s;
}

// =================================================================================================================
template <int NT, int CW>
__device__ __forceinline__ DbSmem db_smem(const DbArgs& a, uint8_t* raw) {
    using C = DbCfg<NT, CW>;
    // xs comes first: the MMA B operand always reads 8 rows per n-tile, rows >= xs_rows alias what follows (harmless garbage in
    // the output columns of rows that do not exist)
    DbSmem sm;
    sm.ldx = a.xs_cols + DB_XS_PAD;
    sm.xs = reinterpret_cast<bf16*>(raw);
    sm.red = reinterpret_cast<float*>(raw + (size_t)a.xs_rows * sm.ldx * 2);
    sm.sp = sm.red + C::RED_FLOATS;                    // [8][224] cross scores / self-attention scores [<= 449]
    sm.sq = sm.sp + 8 * DB_SPLIT_KEYS;                 // [8][64]
    sm.stat = sm.sq + 8 * 64;                          // 64 floats
    sm.rowstat = sm.stat + 64;                         // [DB_MAX_ROWS][mean, rstd]
    sm.spos = reinterpret_cast<int*>(sm.rowstat + 2 * DB_MAX_ROWS);     // [8] text_offset of each window
    sm.stok = sm.spos + 8;                             // [40] current token of each row
    sm.full = reinterpret_cast<uint64_t*>(sm.stok + 40);
    sm.empty = sm.full + DB_MAX_SLOTS;
    sm.desc = reinterpret_cast<DbStageSm*>(sm.empty + DB_MAX_SLOTS);       // [2], 64 bytes each
    sm.seen = reinterpret_cast<volatile int*>(sm.desc + 2);
    sm.ring = raw + a.ring_offset;
    return sm;
}

// ---- producer: one thread streams the CTA's static byte schedule into the ring ------------------------------------------
template <int NT, int CW>
__device__ __forceinline__ void db_producer(const DbArgs& a, uint8_t* raw) {
    const DbModel& M = c_db;
    const int d = M.d, H = M.H, nctas = gridDim.x, cta = blockIdx.x;
    const DbSmem sm = db_smem<NT, CW>(a, raw);
    const int n_stages = M.Ld * 8 + (a.no_vocab ? 0 : 1);
    DbRing ring{0, 0, a.n_slots};
    constexpr int n_ktiles = CROSS_KEYS_PAD / 16, n_vkc = CROSS_KEYS_PAD / 32;
    const long head_elems = (long)64 * CROSS_KEYS_PAD;
    for (int it = 0; it < n_stages; ++it) {
        const int l = it >> 3, st = it == M.Ld * 8 ? DBS_VOCAB : (it & 7);
        db_progress(a.lane_id, 1, it);
        if (st == DBS_SA) continue;
        if (st == DBS_CA) {
            for (int u = (cta + db_stage_rot(DBS_CA, nctas)) % nctas; u < a.W * H * DB_N_SPLITS; u += nctas) {
                // cross-attention unit (window, head, split): one slot of K tiles, one slot with the four V^T dim tiles
                const int wh = u / DB_N_SPLITS, s = u - wh * DB_N_SPLITS, w = wh / H, h = wh - w * H;
                const bf16* base = a.ckv_frag + (long)a.win[w] * a.ckv_window_elems;
                const bf16* kf = base + (long)(l * 2) * H * head_elems + h * head_elems;
                const bf16* vf = base + (long)(l * 2 + 1) * H * head_elems + h * head_elems;
                const int t0 = s * DB_SPLIT_TILES, nt = min(DB_SPLIT_TILES, n_ktiles - t0);
                const int kc0 = s * (DB_SPLIT_TILES / 2), nkc = min(DB_SPLIT_TILES / 2, n_vkc - kc0);
                db_wait<true>(&sm.empty[ring.slot], ring.phase ^ 1, 30);
                mbar_expect_tx(&sm.full[ring.slot], (uint32_t)nt * 2048);
                db_bulk_g2s(sm.ring + (size_t)ring.slot * DB_SLOT, kf + (long)t0 * 1024, (uint32_t)nt * 2048, &sm.full[ring.slot]);
                ring.advance();
                db_wait<true>(&sm.empty[ring.slot], ring.phase ^ 1, 31);
                mbar_expect_tx(&sm.full[ring.slot], (uint32_t)nkc * 4096);
#pragma unroll 1
                for (int dt = 0; dt < 4; ++dt)
                    db_bulk_g2s(sm.ring + (size_t)ring.slot * DB_SLOT + (size_t)dt * nkc * 1024, vf + ((long)dt * n_vkc + kc0) * 512, (uint32_t)nkc * 1024,
                                &sm.full[ring.slot]);
                ring.advance();
            }
            continue;
        }
        const DbStageDesc sd = db_stage_desc(st, l < M.Ld ? l : 0, cta, nctas);
        int u0, u1;
        db_range(sd.n_tiles, sd.vcta, nctas, u0, u1);
        if (sd.ln_g && u1 > u0) {
            db_wait<true>(&sm.empty[ring.slot], ring.phase ^ 1, 32);
            mbar_expect_tx(&sm.full[ring.slot], (uint32_t)d * 8);
            db_bulk_g2s(sm.ring + (size_t)ring.slot * DB_SLOT, sd.ln_g, (uint32_t)d * 4, &sm.full[ring.slot]);
            db_bulk_g2s(sm.ring + (size_t)ring.slot * DB_SLOT + (size_t)d * 4, sd.ln_b, (uint32_t)d * 4, &sm.full[ring.slot]);
            ring.advance();
        }
        const int chunk_kc = st == DBS_M2 ? min(a.xs_cols >> 5, sd.n_kc) : sd.n_kc;
#pragma unroll 1
        for (int t = u0; t < u1; ++t)
#pragma unroll 1
            for (int c0 = 0, nblk = 0; c0 < sd.n_kc; c0 += nblk) {
                nblk = chunk_kc < sd.n_kc ? db_slot_blocks(c0, sd.n_kc, chunk_kc) : min(DB_SLOT_BLOCKS, sd.n_kc - c0);
                const uint32_t bytes = (uint32_t)nblk * 1024;
                db_wait<true>(&sm.empty[ring.slot], ring.phase ^ 1, 33);
                mbar_expect_tx(&sm.full[ring.slot], bytes);
                db_bulk_g2s(sm.ring + (size_t)ring.slot * DB_SLOT, sd.w + ((long)t * sd.n_kc + c0) * 512, bytes, &sm.full[ring.slot]);
                ring.advance();
            }
    }
}

// ---- consumers: the stage loop ------------------------------------------------------------------------------------------
template <int NT, int CW, bool DBG>
__device__ __forceinline__ void db_consumer(const DbArgs& a, uint8_t* raw, unsigned seq, int warp, int lane) {
    using C = DbCfg<NT, CW>;
    const DbModel& M = c_db;
    const int d = M.d, nctas = gridDim.x, cta = blockIdx.x, tid = warp * 32 + lane;
    const int R = a.W * a.nbw;
    const DbSmem sm = db_smem<NT, CW>(a, raw);
    const int n_stages = M.Ld * 8 + (a.no_vocab ? 0 : 1);
    DbRing ring{0, 0, a.n_slots};
    DbDbgT<DBG> dbg{a.dbg + (size_t)cta * DB_DBG_LD, a.dbg != nullptr && tid == 0, -1};
    dbg.mark(0);

    for (int it = 0; it < n_stages; ++it) {
        const int l = it >> 3, st = it == M.Ld * 8 ? DBS_VOCAB : (it & 7);
        const uint32_t ep = seq * 64u + (uint32_t)l + 1u, ep_prev = ep - 1u;      // ep_prev: x3 of the layer below
        if ((tid & 127) == 0) db_progress(a.lane_id, tid == 0 ? 0 : 2, it);       // the leaders of the two tile groups
        if (st == DBS_QKV && l < M.Ld) {
            // Hide DRAM latency: the layer's bias vectors (read by every tile's epilogue) and the cached K / V rows of the
            // self-attention units this CTA will run two stages from now were last touched a step (360 MB of traffic) ago, so
            // they come from HBM; an L2 prefetch now makes them L2 hits (~0.2 us instead of > 1 us on the critical path).
            const DbLayer& L = M.layers[l];
            const float* const* bv = &L.qkv_b;
            for (int i = tid; i < 6 * (4 * d / 32); i += C::CONS) {       // 128-byte lines of the six bias vectors (the longest has 4d floats)
                const int v = i / (4 * d / 32), line = i - v * (4 * d / 32);
                const int n = v == 0 ? 3 * d : v == 4 ? 4 * d : d;
                if (line * 32 < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(bv[v] + line * 32));
            }
            const int H = M.H;
            for (int u = (cta + db_stage_rot(DBS_SA, nctas)) % nctas; u < R * H; u += nctas) {
                const int r = u / H, h = u - r * H, trow = db_trow(a, r), pos = sm.spos[r / a.nbw];
                const int* tab = a.table + trow * 448;
                for (int j = tid; j < 2 * pos; j += C::CONS) {            // one 128-byte line per (K | V, position)
                    const int jj = j >> 1;
                    const bf16* base = a.mkv + (long)(2 * l + (j & 1)) * a.kv_stride;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(base + ((long)tab[jj] * 448 + jj) * d + h * 64));
                }
            }
        }
        if (st == DBS_SA) {
            dbg.mark(2 * it + 1);
            db_stage_self_attn<NT, CW>(sm, a, R, a.mkv + (long)(2 * l) * a.kv_stride, a.mkv + (long)(2 * l + 1) * a.kv_stride, ep, cta, nctas,
                                       db_stage_rot(DBS_SA, nctas), warp, lane);
            dbg.mark(2 * it + 2);
            continue;
        }
        if (st == DBS_CA) {
            dbg.mark(2 * it + 1);
            db_stage_cross_attn<NT, CW>(sm, ring, a, ep, cta, nctas, db_stage_rot(DBS_CA, nctas), warp, lane);
            dbg.mark(2 * it + 2);
            continue;
        }
        dbg.set_sub(it == a.dbg_stage ? it : -1);
        // (double-buffered: a group may still be in the epilogue of stage it - 1's last tile.  Stage it - 2 is finished: EVERY stage
        // passes at least one CTA-wide barrier - a CTA without a tile or an attention unit in a stage too, or thread 0 could
        // refill this descriptor while the other group's last tile of stage it - 2 still reads it and stores its outputs through
        // the wrong pointers: the rare hang of the prompt launches, where only 20 CTAs have a self-attention unit)
        DbStageSm& D = sm.desc[it & 1];
        if (tid == 0) {
            const DbGemv q{st, l < M.Ld ? l : 0, 0, 0, 0, ep, 0};
            D.bias = q.bias(); D.res_ll = q.res_ll(a); D.out_ll = q.out_ll(a); D.out_llb = q.out_llb(a); D.ld_out = q.ld_out(a);
            D.n_valid = q.n_valid(); D.epi = q.epi(); D.res_mode = q.res_mode(a); D.res_epoch = q.res_epoch();
        }
        if (st == DBS_M2) csync<C::CONS>();            // (the other stages' prologues start with this barrier)
        {   // prologue first, with as little live state as possible (the stage descriptor is built after it)
            const DbStageDesc sd = db_stage_desc(st, l < M.Ld ? l : 0, cta, nctas);
            int u0, u1;
            db_range(sd.n_tiles, sd.vcta, nctas, u0, u1);
            if (u1 <= u0 && st != DBS_M2) csync<C::CONS>();
            if (u1 > u0 && st != DBS_M2) {             // a CTA without a tile in this stage does not read its input at all
                if (sd.ln_g) {
                    const uint2* src = st == DBS_CQ ? a.ll_x1b : st == DBS_M1 ? a.ll_x2b : a.ll_x3b;
                    const uint32_t pep = st == DBS_QKV ? ep_prev : st == DBS_VOCAB ? seq * 64u + (uint32_t)M.Ld : ep;
                    db_prologue_ln<NT, CW>(sm, ring, a, R, it == 0 ? DB_PRO_EMBED : DB_PRO_LL, src, pep, warp, lane, dbg);
                } else {                               // attention outputs: 32 words (64 columns) per (row, head)
                    const uint2* src = st == DBS_OUT ? a.ll_att : a.ll_catt;
                    dbg.smark(0);
                    csync<C::CONS>();
                    dbg.smark(1);
                    ll_wait_sentinels<C::CONS>(src, ep, d >> 1, 0, d >> 1, 32, 0, R, tid, 13);
                    dbg.smark(2);
                    ll_copy_region<C::CONS>(src, ep, R, d >> 1, 0, d >> 2, sm.xs, sm.ldx, tid, 3);
                    dbg.smark(3);
                    csync<C::CONS>();
                    dbg.smark(4);
                }
            }
        }
        const DbStageDesc sd = db_stage_desc(st, l < M.Ld ? l : 0, cta, nctas);
        const DbGemv g{st, l < M.Ld ? l : 0, sd.n_tiles, sd.n_kc, sd.vcta, ep, st == DBS_M2 ? min(a.xs_cols >> 5, sd.n_kc) : sd.n_kc};
        dbg.mark(2 * it + 1);
        db_stage_gemv<NT, CW>(sm, ring, g, D, a, R, nctas, warp, lane, dbg);
        dbg.mark(2 * it + 2);
    }
    // the last CTA to leave advances the launch sequence number: by then every CTA has read it
    if (tid == 0) {
        const unsigned old = atomicAdd(&a.barrier[1], 1u);
        if (old == (unsigned)nctas - 1) { a.barrier[1] = 0; a.barrier[2] = seq + 1; }
    }
}

// The role split comes first: whatever is live across setmaxnreg must fit the smaller register file, so each role derives its
// own state afterwards.
template <int NT, int CW, bool DBG>
__global__ void __launch_bounds__(DbCfg<NT, CW>::THREADS, 1) decoder_batch_kernel(const __grid_constant__ DbArgs a) {
    using C = DbCfg<NT, CW>;
    extern __shared__ __align__(128) uint8_t db_raw[];
    if (a.st) {                                        // uniform: every window of the batch already finished (graph replays past the end)
        bool all_done = true;
        for (int w = 0; w < a.W; ++w) all_done = all_done && a.st[w].done != 0;
        if (all_done) return;
    }
    if (threadIdx.x < C::PW * 32) {
        if (C::SETREG) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (threadIdx.x == 0) {
            uint64_t* full = db_smem<NT, CW>(a, db_raw).full;
            for (int s = 0; s < DB_MAX_SLOTS; ++s) { mbar_init(&full[s], 1); mbar_init(&full[DB_MAX_SLOTS + s], C::ARRIVALS); }
            fence_barrier_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) db_producer<NT, CW>(a, db_raw);
        return;
    }
    if (C::SETREG) asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int tid = (int)threadIdx.x - C::PW * 32;
    // a window that has finished keeps stepping (its rows are ignored) at its last position, never past the cache
    if (tid < a.W) db_smem<NT, CW>(a, db_raw).spos[tid] = min(a.st ? a.st[tid].pos : a.text_offset, N_TEXT_CTX - 1);
    if (tid >= 32 && tid < 32 + DB_MAX_SLOTS) db_smem<NT, CW>(a, db_raw).seen[tid - 32] = -1;
    const unsigned seq = a.barrier[2];                 // written by the previous launch
    __syncthreads();
    db_consumer<NT, CW, DBG>(a, db_raw, seq, tid >> 5, tid & 31);
}

// ---- host side -----------------------------------------------------------------------------------------------------------
static unsigned long long* g_h_fault = nullptr;
void db_set_model(const DbModel& m) {
    B200_CHECK(cudaMemcpyToSymbol(c_db, &m, sizeof(DbModel)));
    if (!g_h_fault && cudaHostAlloc((void**)&g_h_fault, 16 + DB_PROGRESS_WORDS * 4 + 8 + DB_NOTE_ROWS * 32, cudaHostAllocMapped) == cudaSuccess) {
        memset(g_h_fault, 0, 16 + DB_PROGRESS_WORDS * 4 + 8 + DB_NOTE_ROWS * 32);
        unsigned long long* dp = nullptr;
        if (cudaHostGetDevicePointer((void**)&dp, g_h_fault, 0) == cudaSuccess) B200_CHECK(cudaMemcpyToSymbol(g_db_fault, &dp, sizeof(dp)));
    }
}
// {where, grid, CTA, thread} of the bounded wait that trapped (0 = none): decoder_batch.cu's `where` codes are 2..13 for LL polls,
// 20..25 for consumer waits on a ring slot, 30..33 for the producer's waits on a free slot
unsigned long long db_fault_word() { return g_h_fault ? *(volatile unsigned long long*)g_h_fault : 0ull; }
// progress table saved with the fault word: [decode lane][DB_PROGRESS_LD][4] 16-bit marks: (stage + 1) of tile group 0, the producer, tile group 1
const unsigned* db_fault_progress() { return g_h_fault ? reinterpret_cast<const unsigned*>(g_h_fault + 2) : nullptr; }
// [0] = n, then n rows {CTA | thread << 16 | where << 32 | grid << 48, address, last value read, epoch expected}: the LL word waits that were long
const unsigned long long* db_fault_ll_word() { return g_h_fault ? g_h_fault + 2 + DB_PROGRESS_WORDS / 2 : nullptr; }

int db_consumer_warps() {
    return 8;
}

// Shared-memory plan for `rows` rows: the activation rows hold K chunks of xs_cols columns (>= d: LayerNorm rows are staged
// whole; the MLP hidden row is consumed in chunks), and the same bytes stage the cached K | V rows of a self-attention unit.
bool db_geometry(int d, int rows, int smem_optin, DbGeometry* g) {
    const int cw = db_consumer_warps();
    const int nt = (rows + 7) / 8;
    if (nt < 1 || nt > DB_MAX_ROWS / 8 || d % 64 != 0 || d > 1280) return false;
    static const int force_slots = getenv("B200_STEP_SLOTS") ? atoi(getenv("B200_STEP_SLOTS")) : 0;
    const int xs_rows = rows < 5 ? 5 : rows;
    // prefer >= 3 ring slots with the largest K chunk; with many rows fall back to 2 slots
    for (int min_slots = 3; min_slots >= 2; --min_slots)
        for (int div = 1; div <= 4; div *= 2) {        // K chunk = 4d, 2d, d columns
            const int xs_cols = std::max(4 * d / div, DB_CA_MIN_COLS - DB_XS_PAD);      // (cross-attention: the second group's scratch, DB_CA_MIN_COLS)
            const size_t front = ((size_t)xs_rows * (xs_cols + DB_XS_PAD) * 2 + db_scratch_bytes(nt, cw) + 127) / 128 * 128;
            int slots = (int)(((size_t)smem_optin - (front < (size_t)smem_optin ? front : (size_t)smem_optin)) / DB_SLOT);
            if (slots > DB_MAX_SLOTS) slots = DB_MAX_SLOTS;
            if (force_slots >= 2 && slots > force_slots) slots = force_slots;
            if (slots < min_slots) continue;
            g->nt = nt; g->xs_cols = xs_cols; g->xs_rows = xs_rows; g->ring_offset = (int)front; g->n_slots = slots;
            const int cap = (int)((size_t)xs_rows * (xs_cols + DB_XS_PAD) * 2 / 256);
            g->sa_cap = cap < 448 ? cap : 448;
            g->smem = front + (size_t)slots * DB_SLOT;
            return true;
        }
    return false;
}

size_t db_ll_words(size_t d, size_t H) {
    const size_t R = DB_MAX_ROWS;
    return R * 3 * d + 2 * (R * d / 2) + 4 * R * d + (size_t)DB_MAX_WINDOWS * H * DB_N_SPLITS * 8 * 66 + R * 2 * d + 3 * (R * d / 2);
}
void db_carve_ll(DbArgs& a, uint2* p, size_t d, size_t H) {
    const size_t R = DB_MAX_ROWS;
    a.ll_qkv = p; p += R * 3 * d; a.ll_att = p; p += R * d / 2; a.ll_catt = p; p += R * d / 2;
    a.ll_x1 = p; p += R * d; a.ll_x2 = p; p += R * d; a.ll_x3 = p; p += R * d; a.ll_q = p; p += R * d;
    a.ll_cap = p; p += (size_t)DB_MAX_WINDOWS * H * DB_N_SPLITS * 8 * 66; a.ll_hid = p; p += R * 2 * d;
    a.ll_x1b = p; p += R * d / 2; a.ll_x2b = p; p += R * d / 2; a.ll_x3b = p;
}

template <int NT, int CW, bool DBG>
static bool db_launch_t(const DbArgs& a, int n_ctas, size_t smem, cudaStream_t s) {
    static size_t attr = 0;
    auto* kern = decoder_batch_kernel<NT, CW, DBG>;
    if (smem > attr) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            record_error("decoder_batch: %zu bytes of shared memory unavailable", smem);
            return false;
        }
        attr = smem;
    }
    DbArgs args = a;
    void* params[] = {(void*)&args};
    // cooperative launch: the LL hand-offs need every CTA of the grid resident at once
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)kern, dim3(n_ctas), dim3(DbCfg<NT, CW>::THREADS), params, smem, s);
    ++g_launch_count;
    if (e != cudaSuccess) { cudaGetLastError(); record_error("decoder_batch launch: %s", cudaGetErrorString(e)); return false; }
    return true;
}

bool db_launch(const DbArgs& a, int n_ctas, cudaStream_t s) {
    const int rows = a.W * a.nbw, nt = (rows + 7) / 8;
    const size_t smem = (size_t)a.ring_offset + (size_t)a.n_slots * DB_SLOT;
    // the stage-timeline marks are a separate instantiation (tools/step_timeline.py): the production kernel carries none
#define DB_CASE(NT_)                                                                                                   \
    case NT_:                                                                                                          \
        return a.dbg ? db_launch_t<NT_, 8, true>(a, n_ctas, smem, s) : db_launch_t<NT_, 8, false>(a, n_ctas, smem, s);
    switch (nt) {
        DB_CASE(1) DB_CASE(2) DB_CASE(3) DB_CASE(4) DB_CASE(5)
    default: record_error("decoder_batch: %d rows", rows); return false;
    }
#undef DB_CASE
}

}  // namespace b200
