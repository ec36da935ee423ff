// Batched device-resident decode loop: every window of a batch advances by one token per launch of decoder_batch_kernel, so
// the decoder weights are read from HBM once per step for ALL windows.  DecodingTask.run /
// _main_loop (whisper/decoding.py:707-816) per window; windows are independent (condition_on_previous_text=False).
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <vector>

#include "api_batch.cuh"
#include "decoder_batch.cuh"
#include "sampling.cuh"
#include "state.cuh"

namespace b200 {

struct BatchCtx {
    bool ready = false;
    size_t d = 0, Ld = 0, V = 0;
    bf16* mkv = nullptr;                 // [2Ld][DB_MAX_ROWS][448][d]
    int* table = nullptr;                // [DB_MAX_ROWS][448]
    float* logits = nullptr;             // [DB_MAX_ROWS][V]
    uint2* ll = nullptr;
    unsigned* barrier = nullptr;         // [4]: [1] leavers, [2] launch sequence number
    DecodeState* st = nullptr;           // [DB_MAX_WINDOWS]
    int* tokens = nullptr;               // [DB_MAX_ROWS][DEC_TOK_LD]
    int* fin_tokens = nullptr;           // [DB_MAX_WINDOWS][DEC_MAX_BEAMS][DEC_TOK_LD]
    float* cand_lp = nullptr; int* cand_tok = nullptr;
    SamplePartials* part = nullptr;      // [DB_MAX_WINDOWS]
    int* d_init = nullptr;               // [256] prompt tokens
    DecodeState* pin_st = nullptr;       // pinned host [2][DB_MAX_WINDOWS]
    unsigned long long* dbg = nullptr;   // stage timeline (b200TestStepTimeline)
    // reference-ABI step through the same kernel (decoder1Predict): its own LL words and sequence counter
    uint2* abi_ll = nullptr; unsigned* abi_barrier = nullptr;
};
static BatchCtx g_bc;
static DbModel g_db_model;
static bool g_db_model_set = false;

struct BatchGraph { cudaGraphExec_t exec = nullptr; long launches = 0; };
static std::map<std::vector<long>, BatchGraph> g_batch_graphs;
void batch_clear_graphs() {
    for (auto& kv : g_batch_graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    g_batch_graphs.clear();
}

void batch_set_model(const DbModel& m) { g_db_model = m; g_db_model_set = true; db_set_model(m); batch_clear_graphs(); }

void batch_free() {
    BatchCtx& c = g_bc;
    batch_clear_graphs();
    dev_free(&c.mkv); dev_free(&c.table); dev_free(&c.logits); dev_free(&c.ll); dev_free(&c.barrier); dev_free(&c.st); dev_free(&c.tokens);
    dev_free(&c.fin_tokens); dev_free(&c.cand_lp); dev_free(&c.cand_tok); dev_free(&c.part); dev_free(&c.d_init); dev_free(&c.dbg);
    dev_free(&c.abi_ll); dev_free(&c.abi_barrier);
    if (c.pin_st) { cudaFreeHost(c.pin_st); c.pin_st = nullptr; }
    c.ready = false;
    g_db_model_set = false;
}

static int step_impl() {                                 // 0 = batched kernel (default), 1 = one-window persistent kernel, 2 = one kernel per stage
    static const int impl = [] {
        const char* e = getenv("B200_STEP_IMPL");
        if (!e) return 0;
        if (!strcmp(e, "mega")) return 1;
        if (!strcmp(e, "v1") || !strcmp(e, "1")) return 2;
        return 0;
    }();
    return impl;
}

bool batch_available() {
    State& s = S();
    if (step_impl() != 0 || !g_db_model_set || !s.dec1_loaded || !s.dec256_loaded) return false;
    mega_available();                                    // fills n_sms / smem_optin
    DbGeometry g;
    return s.Ld <= DB_MAX_LAYERS && db_geometry(s.d, 1, s.smem_optin, &g) && db_geometry(s.d, DEC_MAX_BEAMS, s.smem_optin, &g);
}

// windows one batched step can take with nb beams each (shared-memory plan permitting)
int batch_max_windows(int nb) {
    State& s = S();
    static const int cap = [] { const char* e = getenv("B200_BATCH_WINDOWS"); const int v = e ? atoi(e) : DB_MAX_WINDOWS; return v < 1 ? 1 : (v > DB_MAX_WINDOWS ? DB_MAX_WINDOWS : v); }();
    int w = std::min(cap, DB_MAX_ROWS / nb);
    DbGeometry g;
    while (w > 1 && !db_geometry(s.d, w * nb, s.smem_optin, &g)) --w;
    return w < 1 ? 1 : w;
}

static bool ensure_batch_ctx() {
    BatchCtx& c = g_bc;
    State& s = S();
    if (c.ready && c.d == (size_t)s.d && c.Ld == (size_t)s.Ld && c.V == (size_t)s.V) return true;
    batch_clear_graphs();
    bool ok = true;
    const size_t d = s.d, R = DB_MAX_ROWS;
    ok &= dev_alloc(&c.mkv, (size_t)2 * s.Ld * R * N_TEXT_CTX * d, true);
    ok &= dev_alloc(&c.table, R * N_TEXT_CTX, true);
    ok &= dev_alloc(&c.logits, R * (size_t)s.V);
    ok &= dev_alloc(&c.ll, db_ll_words(d, s.H), true);
    ok &= dev_alloc(&c.barrier, (size_t)4, true);
    if (ok) { const unsigned one = 1; B200_CHECK(cudaMemcpy(c.barrier + 2, &one, sizeof(one), cudaMemcpyHostToDevice)); }   // epoch 0 = "never written"
    ok &= dev_alloc(&c.st, (size_t)DB_MAX_WINDOWS, true);
    ok &= dev_alloc(&c.tokens, R * DEC_TOK_LD, true);
    ok &= dev_alloc(&c.fin_tokens, (size_t)DB_MAX_WINDOWS * DEC_MAX_BEAMS * DEC_TOK_LD, true);
    ok &= dev_alloc(&c.cand_lp, (size_t)DB_MAX_WINDOWS * DEC_MAX_BEAMS * SAMPLE_MAX_K);
    ok &= dev_alloc(&c.cand_tok, (size_t)DB_MAX_WINDOWS * DEC_MAX_BEAMS * SAMPLE_MAX_K);
    ok &= dev_alloc(&c.part, (size_t)DB_MAX_WINDOWS, true);
    ok &= dev_alloc(&c.d_init, (size_t)PREFILL_CTX);
    if (!c.pin_st) ok &= cudaMallocHost((void**)&c.pin_st, 2 * DB_MAX_WINDOWS * sizeof(DecodeState)) == cudaSuccess;
    c.d = d; c.Ld = s.Ld; c.V = s.V;
    c.ready = ok;
    return ok;
}

static bool fill_geometry(DbArgs& a, int rows) {
    State& s = S();
    DbGeometry g;
    if (!db_geometry(s.d, rows, s.smem_optin, &g)) { record_error("decoder_batch: no shared-memory plan for %d rows (d = %d)", rows, s.d); return false; }
    a.xs_cols = g.xs_cols; a.xs_rows = g.xs_rows; a.ring_offset = g.ring_offset; a.n_slots = g.n_slots; a.sa_cap = g.sa_cap;
    return true;
}

static int probe_stage() { static const int v = getenv("B200_STEP_PROBE") ? atoi(getenv("B200_STEP_PROBE")) : -1; return v; }
static int grid_ctas() {
    static const int force = getenv("B200_STEP_CTAS") ? atoi(getenv("B200_STEP_CTAS")) : 0;      // experiments: fixed grid size
    return force > 0 ? force : S().n_sms;
}

// reference ABI (decoder1Predict): one step of the process-global cache, x / mask from the caller, logits out
bool run_step_batch_abi(int nb, int text_offset, const float* d_mask, const float* d_x_in) {
    State& s = S();
    BatchCtx& c = g_bc;
    if (!c.abi_ll) {
        bool ok = dev_alloc(&c.abi_ll, db_ll_words(s.d, s.H), true) && dev_alloc(&c.abi_barrier, (size_t)4, true);
        if (!ok) return false;
        const unsigned one = 1;
        B200_CHECK(cudaMemcpy(c.abi_barrier + 2, &one, sizeof(one), cudaMemcpyHostToDevice));
    }
    DbArgs a{};
    a.W = 1; a.nbw = nb; a.slot_stride = nb; a.win[0] = s.cur_window;
    a.ckv_frag = s.ckv_frag; a.ckv_window_elems = (long)s.ckv_frag_window_elems();
    if (!fill_geometry(a, nb)) return false;
    db_carve_ll(a, c.abi_ll, s.d, s.H);
    a.logits = s.slogits; a.ld_logits = s.V; a.mkv = s.mkv; a.kv_stride = (long)s.bs * N_TEXT_CTX * s.d; a.table = s.table;
    a.mask = d_mask; a.x_in = d_x_in; a.text_offset = text_offset; a.barrier = c.abi_barrier; a.dbg = c.dbg; a.dbg_stage = probe_stage(); a.copy_u = getenv("B200_STEP_COPYU") ? atoi(getenv("B200_STEP_COPYU")) : 13;
    return db_launch(a, grid_ctas(), s.stream);
}

__global__ void batch_init_tokens_kernel(int* tokens, const int* initial, int n, int W, int nb, int eot, int* table) {
    for (int i = threadIdx.x; i < W * nb * DEC_TOK_LD; i += blockDim.x) {
        const int p = i % DEC_TOK_LD;
        tokens[i] = p < n ? initial[p] : eot;
    }
    // every beam of window w reads the prompt's K / V rows from the window's first slot
    for (int i = threadIdx.x; i < W * nb * N_TEXT_CTX; i += blockDim.x) table[i] = (i / (nb * N_TEXT_CTX)) * nb;
}

struct BatchJob { int W, nb, k, n_initial, sample_len, sot_index; const int* windows; };

static DbArgs step_args(const BatchJob& j, bool prompt, int text_offset) {
    State& s = S();
    BatchCtx& c = g_bc;
    DbArgs a{};
    a.W = j.W; a.nbw = prompt ? 1 : j.nb; a.slot_stride = j.nb;
    for (int w = 0; w < j.W; ++w) a.win[w] = j.windows[w];
    a.ckv_frag = s.ckv_frag; a.ckv_window_elems = (long)s.ckv_frag_window_elems();
    fill_geometry(a, a.W * a.nbw);
    db_carve_ll(a, c.ll, s.d, s.H);
    a.logits = c.logits; a.ld_logits = s.V; a.mkv = c.mkv; a.kv_stride = (long)DB_MAX_ROWS * N_TEXT_CTX * s.d; a.table = c.table; a.tokens = c.tokens;
    a.st = prompt ? nullptr : c.st; a.text_offset = text_offset; a.barrier = c.barrier; a.dbg = c.dbg; a.dbg_stage = probe_stage(); a.copy_u = getenv("B200_STEP_COPYU") ? atoi(getenv("B200_STEP_COPYU")) : 13;
    return a;
}

static void launch_batch_sampling(const BatchJob& j, bool shared_logits) {
    State& s = S();
    BatchCtx& c = g_bc;
    SampleBatchArgs b{};
    b.logits = c.logits; b.ld_logits = s.V;
    b.row_stride_w = shared_logits ? 1 : j.nb; b.row_stride_b = shared_logits ? 0 : 1;
    b.tokens = c.tokens; b.table = c.table; b.fin_tokens = c.fin_tokens; b.st = c.st; b.part = c.part; b.cand_lp = c.cand_lp; b.cand_tok = c.cand_tok;
    b.spec = decode_spec(); b.W = j.W; b.nb = j.nb; b.k = j.k; b.slot_stride = j.nb; b.n_text_ctx = N_TEXT_CTX;
    sample_and_update_batch(b, s.stream);
}

static void one_batch_step(const BatchJob& j) {
    const DbArgs a = step_args(j, false, 0);
    db_launch(a, grid_ctas(), S().stream);
    launch_batch_sampling(j, false);
}

constexpr int BATCH_GRAPH_STEPS = 8;
static BatchGraph* batch_graph(const BatchJob& j) {
    State& s = S();
    std::vector<long> key = {j.W, j.nb, j.k, (long)(size_t)g_bc.dbg, grid_ctas()};
    for (int w = 0; w < j.W; ++w) key.push_back(j.windows[w]);
    auto it = g_batch_graphs.find(key);
    if (it != g_batch_graphs.end()) return &it->second;
    if (g_batch_graphs.size() > 32) batch_clear_graphs();
    cudaGraph_t graph = nullptr;
    const long l0 = g_launch_count;
    if (cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    for (int i = 0; i < BATCH_GRAPH_STEPS; ++i) one_batch_step(j);
    BatchGraph g;
    g.launches = g_launch_count - l0;
    g_launch_count = l0;                                               // capture issued nothing; replays are counted per launch
    if (cudaStreamEndCapture(s.stream, &graph) != cudaSuccess || !graph) { cudaGetLastError(); return nullptr; }
    if (cudaGraphInstantiate(&g.exec, graph, 0) != cudaSuccess) { cudaGetLastError(); cudaGraphDestroy(graph); return nullptr; }
    cudaGraphDestroy(graph);
    return &g_batch_graphs.emplace(key, g).first->second;
}

// finalize (decoding.py:411-431 / :320-325) of one window from host copies of its state
static int emit_window(const DecodeState& h, const int* tok, const int* fin, int nb, int n_initial, int eot, int* out_tokens, int* out_lengths,
                       float* out_sum_logprobs, float* out_no_speech) {
    int n_cand = 0;
    auto emit = [&](const int* seq, int len, float score) {
        int* dst = out_tokens + (size_t)n_cand * DEC_TOK_LD;
        for (int i = 0; i < DEC_TOK_LD; ++i) dst[i] = i < len ? seq[i] : eot;
        int l = 0;
        while (n_initial + l < len && seq[n_initial + l] != eot) ++l;  // tokens before the first EOT after sample_begin (:776-779)
        out_lengths[n_cand] = l; out_sum_logprobs[n_cand] = score; ++n_cand;
    };
    if (!h.beam_mode) {
        emit(tok, h.L, h.sum_lp[0]);
    } else {
        for (int f = 0; f < h.n_finished; ++f) emit(fin + (size_t)f * DEC_TOK_LD, h.fin_len[f], h.fin_score[f]);
        if (n_cand < nb) {                                              // not enough finished: add live beams, best first (:418-424)
            std::vector<int> order(nb);
            for (int i = 0; i < nb; ++i) order[i] = i;
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return h.sum_lp[a] > h.sum_lp[b]; });
            for (int i = 0; i < nb && n_cand < nb; ++i) emit(tok + (size_t)order[i] * DEC_TOK_LD, h.L, h.sum_lp[order[i]]);
        }
    }
    for (int i = n_cand; i < nb; ++i) { out_lengths[i] = -1; out_sum_logprobs[i] = -INFINITY; }
    if (out_no_speech) *out_no_speech = h.no_speech_prob;
    return h.step;
}

int decode_windows_batch(const int* windows, int n_windows, const int* initial_tokens, int n_initial, int beam_size, int sample_len,
                         int without_timestamps, int max_initial_timestamp_index, int* out_tokens, int* out_lengths,
                         float* out_sum_logprobs, float* out_no_speech, int* out_steps) {
    State& s = S();
    BatchCtx& c = g_bc;
    if (!ensure_batch_ctx()) return 0;
    const DecodeSpec spec = decode_spec();
    const int nb = beam_size > 0 ? beam_size : 1, cand = nb;
    const int w_max = batch_max_windows(nb);
    int sot_index = -1;
    for (int i = 0; i < n_initial; ++i) if (initial_tokens[i] == spec.sot) sot_index = i;   // tokens.index(sot) (:617)
    cudaStream_t st = s.stream;
    int total_steps = 0;
    for (int w0 = 0; w0 < n_windows; w0 += w_max) {
        BatchJob j{std::min(w_max, n_windows - w0), nb, beam_size > 0 ? nb + 1 : 1, n_initial, sample_len, sot_index, windows + w0};
        // ---- decode state, token histories, slot tables ----
        DecodeState h{};
        h.L = n_initial; h.pos = n_initial - 1; h.sample_begin = n_initial; h.sample_len = sample_len;
        h.beam_mode = beam_size > 0; h.without_timestamps = without_timestamps; h.max_initial_ts = max_initial_timestamp_index;
        h.suppress_blank = 1; h.no_speech_prob = NAN;
        std::vector<DecodeState> hs(j.W, h);
        B200_CHECK(cudaMemcpyAsync(c.st, hs.data(), hs.size() * sizeof(DecodeState), cudaMemcpyHostToDevice, st));
        B200_CHECK(cudaMemcpyAsync(c.d_init, initial_tokens, (size_t)n_initial * sizeof(int), cudaMemcpyHostToDevice, st));
        batch_init_tokens_kernel<<<1, 1024, 0, st>>>(c.tokens, c.d_init, n_initial, j.W, nb, spec.eot, c.table);
        B200_LAUNCH_CHECK();
        B200_CHECK(cudaStreamSynchronize(st));                          // hs / initial_tokens are host temporaries
        {
            // ---- prompt: all beams hold the same tokens (decoding.py:761), so it runs as n_initial one-row-per-window steps into
            //      the window's first cache slot; a causal prefill over n rows IS n steps, and the batched kernel streams each
            //      weight once per position for all windows ----
            StageTimer t(ST_DECODER256);
            for (int p = 0; p < n_initial; ++p) {
                DbArgs a = step_args(j, true, p);
                a.no_vocab = !(p == n_initial - 1 || p == sot_index);
                db_launch(a, grid_ctas(), st);
                if (p == sot_index) no_speech_prob_batch(c.logits, s.V, s.V, spec.no_speech, c.st, j.W, st);   // logits at the sot position (:716-720)
            }
        }
        {
            StageTimer t(ST_SAMPLING);
            launch_batch_sampling(j, true);
        }
        int steps = 1;
        {
            StageTimer t(ST_DECODER1);
            if (steps < sample_len) { one_batch_step(j); ++steps; }     // eager once: sets kernel attributes before any capture
            BatchGraph* graph = steps < sample_len ? batch_graph(j) : nullptr;
            // completion is polled one round behind the issue front, so the GPU never idles while the host synchronises
            cudaEvent_t ev[2];
            for (auto& e : ev) B200_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            int issued = 0, checked = 0;
            auto issue = [&] {
                if (steps < sample_len) {
                    // launches past sample_len / completion are no-ops: every kernel checks DecodeState::done first
                    if (graph) { B200_CHECK(cudaGraphLaunch(graph->exec, st)); g_launch_count += graph->launches; }
                    else for (int i = 0; i < BATCH_GRAPH_STEPS; ++i) one_batch_step(j);
                    steps += BATCH_GRAPH_STEPS;
                }
                B200_CHECK(cudaMemcpyAsync(c.pin_st + (issued & 1) * DB_MAX_WINDOWS, c.st, (size_t)j.W * sizeof(DecodeState), cudaMemcpyDeviceToHost, st));
                B200_CHECK(cudaEventRecord(ev[issued & 1], st));
                ++issued;
            };
            for (;;) {
                while (issued - checked < 2 && steps < sample_len) issue();
                if (checked == issued) break;
                B200_CHECK(cudaEventSynchronize(ev[checked & 1]));
                const DecodeState* ps = c.pin_st + (checked & 1) * DB_MAX_WINDOWS;
                bool all = true;
                for (int w = 0; w < j.W; ++w) all = all && ps[w].done != 0;
                ++checked;
                if (all) break;
            }
            for (auto& e : ev) cudaEventDestroy(e);
        }
        // ---- results ----
        std::vector<int> tok((size_t)j.W * nb * DEC_TOK_LD), fin((size_t)j.W * DEC_MAX_BEAMS * DEC_TOK_LD);
        B200_CHECK(cudaMemcpyAsync(hs.data(), c.st, hs.size() * sizeof(DecodeState), cudaMemcpyDeviceToHost, st));
        B200_CHECK(cudaMemcpyAsync(tok.data(), c.tokens, tok.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
        B200_CHECK(cudaMemcpyAsync(fin.data(), c.fin_tokens, fin.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
        B200_CHECK(cudaStreamSynchronize(st));
        for (int w = 0; w < j.W; ++w) {
            const size_t o = (size_t)(w0 + w);
            const int n = emit_window(hs[w], tok.data() + (size_t)w * nb * DEC_TOK_LD, fin.data() + (size_t)w * DEC_MAX_BEAMS * DEC_TOK_LD, nb, n_initial,
                                      spec.eot, out_tokens + o * cand * DEC_TOK_LD, out_lengths + o * cand, out_sum_logprobs + o * cand,
                                      out_no_speech ? out_no_speech + o : nullptr);
            if (out_steps) out_steps[o] = n;
            total_steps += n;
        }
    }
    return total_steps;
}

// stage timeline of the batched kernel (tools/step_timeline.py): enable allocates + clears the buffer, disable copies it out
int batch_timeline(int enable, unsigned long long* out, int cap_ctas) {
    State& s = S();
    BatchCtx& c = g_bc;
    mega_available();
    const size_t n = (size_t)s.n_sms * DB_DBG_LD;
    if (enable) {
        if (!c.dbg && !dev_alloc(&c.dbg, n)) return 0;
        B200_CHECK(cudaMemset(c.dbg, 0, n * sizeof(unsigned long long)));
        return 1;
    }
    if (!c.dbg) return 0;
    B200_CHECK(cudaDeviceSynchronize());
    const int n_ctas = std::min(cap_ctas, s.n_sms);
    B200_CHECK(cudaMemcpy(out, c.dbg, (size_t)n_ctas * DB_DBG_LD * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    dev_free(&c.dbg);
    return n_ctas;
}

}  // namespace b200
