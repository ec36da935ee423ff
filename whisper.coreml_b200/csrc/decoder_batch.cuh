// Batched persistent decoder1 step kernel (decoder_batch.cu): ONE launch advances every beam of every window of a batch by
// one token, so the decoder weights are streamed from HBM once per step for all windows.
#pragma once
#include "common.cuh"
#include "sampling.cuh"
#include "state.cuh"

namespace b200 {

constexpr int DB_MAX_LAYERS = 32;
constexpr int DB_MAX_WINDOWS = 8;        // windows per batched step (each with its own cross K/V, KV cache rows and decode state)
constexpr int DB_MAX_ROWS = 40;          // rows = windows x beams per step: 5 n-tiles of mma.m16n8k16
constexpr int DB_DBG_LD = 640;           // timeline marks per CTA (2 per stage + 1)
constexpr int DB_NOTE_ROWS = 256;         // long LL-word waits noted for the fault report
constexpr int DB_PROGRESS_LD = 160, DB_PROGRESS_WORDS = 8 * DB_PROGRESS_LD * 2; // progress: [decode lane][CTA][4] 16-bit marks (group 0, producer, group 1, -)
constexpr int DB_N_SPLITS = 7;           // key splits of a cross-attention head (224 keys each)

struct DbLayer {
    const bf16 *qkv, *attn_out, *cross_q, *cross_out, *mlp1, *mlp2;                  // fragment-major weights (export.py: to_frag)
    const float *qkv_b, *attn_out_b, *cross_q_b, *cross_out_b, *mlp1_b, *mlp2_b;
    const float *ln1_w, *ln1_b, *ln2_w, *ln2_b, *ln3_w, *ln3_b;
};
struct DbModel {                         // copied into constant memory when both decoders are loaded
    int d, H, Ld, V, n_tiles_vocab;
    const bf16* tok_emb; const bf16* tok_emb_frag; const float* pos_emb; const float* ln_w; const float* ln_b;
    DbLayer layers[DB_MAX_LAYERS];
};

// Row r of a step = (window w = r / nbw, beam b = r % nbw).  Its token history, KV slot table row and physical KV cache slot are
// all index w * slot_stride + b (slot_stride = beams of the decode; the prompt runs with nbw = 1 into the first slot of each window).
struct DbArgs {
    int W, nbw, slot_stride;
    int lane_id;                         // decode lane of the launch (row of the progress table, 0 .. 7)
    int win[DB_MAX_WINDOWS];             // cross K/V window (index into ckv_frag) of each batch window
    const bf16* ckv_frag; long ckv_window_elems;
    // shared-memory geometry (db_geometry): bf16 activation rows [xs_rows][xs_cols + 32], scratch, then the ring of n_slots 40 KB slots
    int xs_cols, xs_rows, ring_offset, n_slots, sa_cap;
    // activations handed from stage to stage as LL words {payload, epoch}; all sized for DB_MAX_ROWS rows
    uint2 *ll_qkv;                       // [R][3d]   fp32   q | k | v of the new token
    uint2 *ll_att, *ll_catt;             // [R][d/2]  bf16x2 self- / cross-attention output
    uint2 *ll_x1, *ll_x2, *ll_x3;        // [R][d]    fp32   residual stream after self-attention / cross-attention / MLP
    uint2 *ll_x1b, *ll_x2b, *ll_x3b;     // [R][d/2]  bf16x2 the same values for the LayerNorm prologues
    uint2 *ll_q;                         // [R][d]    fp32   cross-attention query
    uint2 *ll_cap;                       // [W][H][7][8][66] fp32 cross-attention partials (max, sum, o[64]) per key split
    uint2 *ll_hid;                       // [R][2d]   bf16x2 MLP hidden activations
    float* logits; long ld_logits;       // [R][ld_logits]
    bf16* mkv; long kv_stride;           // KV cache [2Ld][slots][448][d]; elements per [slots][448][d] plane
    int* table; const int* tokens;       // KV slot table [slots][448]; token histories [slots][DEC_TOK_LD] (nullptr with x_in)
    const DecodeState* st;               // device-resident decode loop: [W] decode states (pos, done), else nullptr
    const float* mask;                   // reference ABI: additive (449) mask on the device, else nullptr
    const float* x_in;                   // reference ABI: embedded tokens fp32 [R][d] on the device, else nullptr
    int text_offset;                     // used when st == nullptr
    int no_vocab;                        // stop after the last layer (prompt positions whose logits nobody reads)
    unsigned* barrier;                   // [1] CTAs that have left the kernel, [2] launch sequence number
    unsigned long long* dbg;             // optional stage timeline: [n_ctas][DB_DBG_LD] %globaltimer values (0 = not reached)
    int dbg_stage;                       // stage index whose inner marks are recorded at 600.. (B200_STEP_PROBE), -1 = none
};

struct DbGeometry { int nt, xs_cols, xs_rows, ring_offset, n_slots, sa_cap; size_t smem; };
void db_set_model(const DbModel& m);
int db_consumer_warps();                               // B200_STEP_WARPS (4, 7 or 8)
bool db_geometry(int d, int rows, int smem_optin, DbGeometry* g);
size_t db_ll_words(size_t d, size_t H);
void db_carve_ll(DbArgs& a, uint2* base, size_t d, size_t H);
bool db_launch(const DbArgs& a, int n_ctas, cudaStream_t s);
const unsigned* db_fault_progress();
const unsigned long long* db_fault_ll_word();
unsigned long long db_fault_word();                     // what the kernel left before a timeout trap (0 = nothing)

}  // namespace b200
