// Batched device-resident decode loop: every window of a batch advances by one token per launch of decoder_batch_kernel, so
// the decoder weights are read from HBM once per step for ALL windows.  DecodingTask.run /
// _main_loop (whisper/decoding.py:707-816) per window; windows are independent (condition_on_previous_text=False).
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "api_batch.cuh"
#include "decoder_batch.cuh"
#include "sampling.cuh"
#include "state.cuh"

namespace b200 {

// One decode lane: a group of windows that advances together in ONE batched step kernel, with its own KV cache, slot tables,
// LL words, decode states and stream.  Lanes run concurrently, each on its share of the SMs: the step is a latency-bound chain of
// ~33 dependent stages, so independent groups overlap each other's stalls (and one lane's sampling kernel hides behind the other
// lanes' step kernels), while the windows inside a lane share one pass over the decoder weights.
constexpr int MAX_DECODE_LANES = 8;
struct BatchCtx {
    bool ready = false;
    size_t d = 0, Ld = 0, V = 0, rows = 0;
    cudaStream_t stream = nullptr;       // lane 0 runs on the library stream
    bf16* mkv = nullptr;                 // [2Ld][rows][448][d]
    int* table = nullptr;                // [rows][448]
    float* logits = nullptr;             // [rows][V]
    uint2* ll = nullptr;
    unsigned* barrier = nullptr;         // [4]: [1] leavers, [2] launch sequence number
    DecodeState* st = nullptr;           // [DB_MAX_WINDOWS]
    int* tokens = nullptr;               // [rows][DEC_TOK_LD]
    int* fin_tokens = nullptr;           // [DB_MAX_WINDOWS][DEC_MAX_BEAMS][DEC_TOK_LD]
    float* cand_lp = nullptr; int* cand_tok = nullptr;
    SamplePartials* part = nullptr;      // [DB_MAX_WINDOWS]
    int* d_init = nullptr;               // [256] prompt tokens
    DecodeState* pin_st = nullptr;       // pinned host [2][DB_MAX_WINDOWS]
    cudaEvent_t ev[2] = {nullptr, nullptr};
};
static BatchCtx g_lane[MAX_DECODE_LANES];
static unsigned long long* g_dbg = nullptr;             // stage timeline (b200TestStepTimeline); single-lane runs only
// reference-ABI step through the same kernel (decoder1Predict): its own LL words and sequence counter
static uint2* g_abi_ll = nullptr; static unsigned* g_abi_barrier = nullptr;
static bool g_db_model_set = false;

struct BatchGraph { cudaGraphExec_t exec = nullptr; long launches = 0; };
static std::map<std::vector<long>, BatchGraph> g_batch_graphs;
void batch_clear_graphs() {
    for (auto& kv : g_batch_graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    g_batch_graphs.clear();
}

// A bounded wait of the step kernel gave up (protocol bug or a dead peer CTA): log which wait and how far every CTA of every lane got.
static void report_step_fault(unsigned long long f) {
    static bool reported = false;
    if (reported) return;
    reported = true;
    record_error("decoder_batch_kernel gave up in wait %d: CTA %d of %d, thread %d", (int)(f >> 48 & 0x7fff), (int)(f >> 16 & 0xffff),
                 (int)(f >> 32 & 0xffff), (int)(f & 0xffff));
    if (const unsigned long long* x = db_fault_ll_word()) {
        // the long LL-word waits, one line per (CTA, where) with the lowest word waited for
        std::map<std::vector<long>, std::vector<unsigned long long>> rows;
        for (unsigned long long i = 0; i < x[0] && i < (unsigned long long)DB_NOTE_ROWS; ++i) {
            const unsigned long long* r = x + 1 + i * 4;
            long lane = -1, off = -1;
            for (int l = 0; l < MAX_DECODE_LANES; ++l) {
                const long o = g_lane[l].ll ? (long)((const uint2*)r[1] - g_lane[l].ll) : -1;
                if (o >= 0 && o < (long)db_ll_words(S().d, S().H)) { lane = l; off = o; }
            }
            const std::vector<long> key = {lane, (long)(r[0] & 0xffff), (long)(r[0] >> 32 & 0xffff)};
            auto it = rows.find(key);
            if (it == rows.end() || (unsigned long long)off < it->second[0]) rows[key] = {(unsigned long long)off, r[0] >> 16 & 0xffff, r[2], r[3]};
        }
        for (auto& kv : rows)
            fprintf(stderr, "[whisper_b200] long wait: lane %ld CTA %ld where %ld thread %llu: LL word %llu holds payload 0x%08x epoch %u, expected epoch %u\n", kv.first[0],
                    kv.first[1], kv.first[2], kv.second[1], kv.second[0], (unsigned)kv.second[2], (unsigned)(kv.second[2] >> 32), (unsigned)kv.second[3]);
    }
    const unsigned* p = db_fault_progress();
    for (int lane = 0; p && lane < 8; ++lane) {
        std::string line;
        char buf[48];
        for (int c = 0; c < DB_PROGRESS_LD; ++c) {
            const unsigned v = p[(lane * DB_PROGRESS_LD + c) * 2], v1 = p[(lane * DB_PROGRESS_LD + c) * 2 + 1];
            if (!v) continue;
            snprintf(buf, sizeof buf, " %d:%d,%d/%d", c, (int)(v & 0xffff) - 1, (int)(v1 & 0xffff) - 1, (int)(v >> 16) - 1);
            line += buf;
        }
        if (!line.empty()) fprintf(stderr, "[whisper_b200] lane %d progress (CTA:stage of tile group 0,1/producer):%s\n", lane, line.c_str());
    }
}
void batch_set_model(const DbModel& m) { g_db_model_set = true; db_set_model(m); batch_clear_graphs(); }

static void lane_free(BatchCtx& c, bool is_lane0) {
    dev_free(&c.mkv); dev_free(&c.table); dev_free(&c.logits); dev_free(&c.ll); dev_free(&c.barrier); dev_free(&c.st); dev_free(&c.tokens);
    dev_free(&c.fin_tokens); dev_free(&c.cand_lp); dev_free(&c.cand_tok); dev_free(&c.part); dev_free(&c.d_init);
    if (c.pin_st) { cudaFreeHost(c.pin_st); c.pin_st = nullptr; }
    for (auto& e : c.ev) if (e) { cudaEventDestroy(e); e = nullptr; }
    if (!is_lane0 && c.stream) cudaStreamDestroy(c.stream);
    c.stream = nullptr;
    c.ready = false; c.rows = 0;
}
void batch_free() {
    batch_clear_graphs();
    for (int i = 0; i < MAX_DECODE_LANES; ++i) lane_free(g_lane[i], i == 0);
    dev_free(&g_dbg); dev_free(&g_abi_ll); dev_free(&g_abi_barrier);
    g_db_model_set = false;
}

bool batch_available() {
    State& s = S();
    if (!g_db_model_set || !s.dec1_loaded || !s.dec256_loaded) return false;
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

static bool ensure_lane(int i, int rows) {
    BatchCtx& c = g_lane[i];
    State& s = S();
    if (c.ready && c.d == (size_t)s.d && c.Ld == (size_t)s.Ld && c.V == (size_t)s.V && c.rows >= (size_t)rows) return true;
    batch_clear_graphs();
    lane_free(c, i == 0);
    bool ok = true;
    const size_t d = s.d, R = rows;
    if (i == 0) c.stream = s.stream; else ok &= cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking) == cudaSuccess;
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
    ok &= cudaMallocHost((void**)&c.pin_st, 2 * DB_MAX_WINDOWS * sizeof(DecodeState)) == cudaSuccess;
    for (auto& e : c.ev) ok &= cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
    c.d = d; c.Ld = s.Ld; c.V = s.V; c.rows = R;
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
    if (!g_abi_ll) {
        bool ok = dev_alloc(&g_abi_ll, db_ll_words(s.d, s.H), true) && dev_alloc(&g_abi_barrier, (size_t)4, true);
        if (!ok) return false;
        const unsigned one = 1;
        B200_CHECK(cudaMemcpy(g_abi_barrier + 2, &one, sizeof(one), cudaMemcpyHostToDevice));
    }
    DbArgs a{};
    a.W = 1; a.nbw = nb; a.slot_stride = nb; a.win[0] = s.cur_window;
    a.ckv_frag = s.ckv_frag; a.ckv_window_elems = (long)s.ckv_frag_window_elems();
    if (!fill_geometry(a, nb)) return false;
    db_carve_ll(a, g_abi_ll, s.d, s.H);
    a.logits = s.slogits; a.ld_logits = s.V; a.mkv = s.mkv; a.kv_stride = (long)s.bs * N_TEXT_CTX * s.d; a.table = s.table;
    a.mask = d_mask; a.x_in = d_x_in; a.text_offset = text_offset; a.barrier = g_abi_barrier; a.dbg = g_dbg; a.dbg_stage = probe_stage();
    return db_launch(a, grid_ctas(), s.stream);
}

// decoder1StepFused: the reference ABI's cache (s.mkv / s.table, filled by decoder256Predict / earlier steps), token histories
// from the caller; the step kernel embeds the last token of every beam, the sampling kernels leave the candidates
void step_fused_abi(const int* tokens_hist, int n_hist, int sample_begin, int text_offset, int without_timestamps,
                    int max_initial_timestamp_index, float* out_logprob, int* out_token) {
    State& s = S();
    const int nb = s.bs, k = nb + 1;
    if (!ensure_lane(0, nb)) return;
    BatchCtx& c = g_lane[0];
    cudaStream_t st = s.stream;
    const DecodeSpec spec = decode_spec();
    std::vector<int> rows((size_t)nb * DEC_TOK_LD, spec.eot);
    for (int b = 0; b < nb; ++b) memcpy(&rows[(size_t)b * DEC_TOK_LD], tokens_hist + (size_t)b * n_hist, (size_t)n_hist * sizeof(int));
    B200_CHECK(cudaMemcpyAsync(c.tokens, rows.data(), rows.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    DecodeState h{};
    h.L = n_hist; h.pos = text_offset; h.sample_begin = sample_begin; h.sample_len = 1 << 30; h.beam_mode = 1;
    h.without_timestamps = without_timestamps; h.max_initial_ts = max_initial_timestamp_index; h.suppress_blank = 1;
    B200_CHECK(cudaMemcpyAsync(c.st, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    if (!g_abi_ll) {
        bool ok = dev_alloc(&g_abi_ll, db_ll_words(s.d, s.H), true) && dev_alloc(&g_abi_barrier, (size_t)4, true);
        if (!ok) return;
        const unsigned one = 1;
        B200_CHECK(cudaMemcpy(g_abi_barrier + 2, &one, sizeof(one), cudaMemcpyHostToDevice));
    }
    DbArgs a{};
    a.W = 1; a.nbw = nb; a.slot_stride = nb; a.win[0] = s.cur_window;
    a.ckv_frag = s.ckv_frag; a.ckv_window_elems = (long)s.ckv_frag_window_elems();
    if (!fill_geometry(a, nb)) return;
    db_carve_ll(a, g_abi_ll, s.d, s.H);
    a.logits = s.slogits; a.ld_logits = s.V; a.mkv = s.mkv; a.kv_stride = (long)s.bs * N_TEXT_CTX * s.d; a.table = s.table; a.tokens = c.tokens;
    a.text_offset = text_offset; a.barrier = g_abi_barrier; a.dbg_stage = -1;
    db_launch(a, grid_ctas(), st);
    SampleArgs sa{};
    sa.logits = s.slogits; sa.ld_logits = s.V; sa.tokens = c.tokens; sa.st = c.st; sa.spec = spec; sa.nb = nb; sa.k = k;
    sa.cand_lp = c.cand_lp; sa.cand_tok = c.cand_tok; sa.part = c.part;
    sample_partial(sa, st);
    BeamUpdateArgs ba{};
    ba.part = c.part; ba.timestamp_begin = spec.timestamp_begin; ba.update = 0;                     // candidates only
    ba.cand_lp = c.cand_lp; ba.cand_tok = c.cand_tok; ba.nb = nb; ba.k = k; ba.tokens = c.tokens; ba.table = s.table;
    ba.fin_tokens = c.fin_tokens; ba.st = c.st; ba.eot = spec.eot; ba.n_text_ctx = N_TEXT_CTX;
    beam_update(ba, st);
    B200_CHECK(cudaMemcpyAsync(out_logprob, c.cand_lp, (size_t)nb * k * sizeof(float), cudaMemcpyDeviceToHost, st));
    B200_CHECK(cudaMemcpyAsync(out_token, c.cand_tok, (size_t)nb * k * sizeof(int), cudaMemcpyDeviceToHost, st));
    B200_CHECK(cudaStreamSynchronize(st));
}

__global__ void batch_init_tokens_kernel(int* tokens, const int* initial, int n, int W, int nb, int eot, int* table) {
    for (int i = threadIdx.x; i < W * nb * DEC_TOK_LD; i += blockDim.x) {
        const int p = i % DEC_TOK_LD;
        tokens[i] = p < n ? initial[p] : eot;
    }
    // every beam of window w reads the prompt's K / V rows from the window's first slot
    for (int i = threadIdx.x; i < W * nb * N_TEXT_CTX; i += blockDim.x) table[i] = (i / (nb * N_TEXT_CTX)) * nb;
}

struct BatchJob {
    int lane, W, nb, k, n_initial, sample_len, sot_index, n_ctas; const int* windows;
    int w0 = 0, steps = 0, issued = 0, checked = 0; bool done = false; BatchGraph* graph = nullptr;
};

static DbArgs step_args(const BatchJob& j, bool prompt, int text_offset) {
    State& s = S();
    BatchCtx& c = g_lane[j.lane];
    DbArgs a{};
    a.W = j.W; a.nbw = prompt ? 1 : j.nb; a.slot_stride = j.nb; a.lane_id = j.lane;
    for (int w = 0; w < j.W; ++w) a.win[w] = j.windows[w];
    a.ckv_frag = s.ckv_frag; a.ckv_window_elems = (long)s.ckv_frag_window_elems();
    fill_geometry(a, a.W * a.nbw);
    db_carve_ll(a, c.ll, s.d, s.H);
    a.logits = c.logits; a.ld_logits = s.V; a.mkv = c.mkv; a.kv_stride = (long)c.rows * N_TEXT_CTX * s.d; a.table = c.table; a.tokens = c.tokens;
    a.st = prompt ? nullptr : c.st; a.text_offset = text_offset; a.barrier = c.barrier; a.dbg = g_dbg; a.dbg_stage = probe_stage();
    return a;
}

static void launch_batch_sampling(const BatchJob& j, bool shared_logits) {
    State& s = S();
    BatchCtx& c = g_lane[j.lane];
    SampleBatchArgs b{};
    b.logits = c.logits; b.ld_logits = s.V;
    b.row_stride_w = shared_logits ? 1 : j.nb; b.row_stride_b = shared_logits ? 0 : 1;
    b.tokens = c.tokens; b.table = c.table; b.fin_tokens = c.fin_tokens; b.st = c.st; b.part = c.part; b.cand_lp = c.cand_lp; b.cand_tok = c.cand_tok;
    b.spec = decode_spec(); b.W = j.W; b.nb = j.nb; b.k = j.k; b.slot_stride = j.nb; b.n_text_ctx = N_TEXT_CTX;
    b.dbg = g_dbg ? g_dbg + 200 * DB_DBG_LD : nullptr;      // (rows 200.. of the timeline buffer: no launch has that many CTAs)
    sample_and_update_batch(b, c.stream);
}

__global__ void __launch_bounds__(256) batch_noop_kernel() {}
static void one_batch_step(const BatchJob& j) {
    const DbArgs a = step_args(j, false, 0);
    db_launch(a, j.n_ctas, g_lane[j.lane].stream);
    // timing experiments only (the decode is then meaningless): 1 = no sampling kernel, 2 = an empty kernel of the same grid
    static const int experiment = getenv("B200_SAMPLING_EXPERIMENT") ? atoi(getenv("B200_SAMPLING_EXPERIMENT")) : 0;
    if (experiment == 1) return;
    if (experiment == 2) { batch_noop_kernel<<<dim3(SAMPLE_CHUNKS, j.nb, j.W), 256, 0, g_lane[j.lane].stream>>>(); return; }
    launch_batch_sampling(j, false);
}

constexpr int BATCH_GRAPH_STEPS = 8;
static BatchGraph* batch_graph(const BatchJob& j) {
    cudaStream_t st = g_lane[j.lane].stream;
    std::vector<long> key = {j.lane, j.W, j.nb, j.k, (long)(size_t)g_dbg, j.n_ctas};
    for (int w = 0; w < j.W; ++w) key.push_back(j.windows[w]);
    auto it = g_batch_graphs.find(key);
    if (it != g_batch_graphs.end()) return &it->second;
    if (g_batch_graphs.size() > 64) return nullptr;                    // (steps are then launched one by one)
    cudaGraph_t graph = nullptr;
    const long l0 = g_launch_count;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    for (int i = 0; i < BATCH_GRAPH_STEPS; ++i) one_batch_step(j);
    BatchGraph g;
    g.launches = g_launch_count - l0;
    g_launch_count = l0;                                               // capture issued nothing; replays are counted per launch
    if (cudaStreamEndCapture(st, &graph) != cudaSuccess || !graph) { cudaGetLastError(); return nullptr; }
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
    if (!h.beam_mode) {                                                 // GreedyDecoder.finalize (:320-325): every row is a candidate
        for (int b = 0; b < nb; ++b) emit(tok + (size_t)b * DEC_TOK_LD, h.L, h.sum_lp[b]);
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

// issue the next BATCH_GRAPH_STEPS steps of a lane and the read-back of its decode states (no host sync)
static void lane_issue(BatchJob& j) {
    BatchCtx& c = g_lane[j.lane];
    if (j.steps == 1 && j.steps < j.sample_len) {                       // eager once: sets kernel attributes before any capture
        one_batch_step(j); ++j.steps;
        if (j.steps < j.sample_len) j.graph = batch_graph(j);
    }
    if (j.steps < j.sample_len) {
        // launches past sample_len / completion are no-ops: every kernel checks DecodeState::done first
        if (j.graph) { B200_CHECK(cudaGraphLaunch(j.graph->exec, c.stream)); g_launch_count += j.graph->launches; }
        else for (int i = 0; i < BATCH_GRAPH_STEPS; ++i) one_batch_step(j);
        j.steps += BATCH_GRAPH_STEPS;
    }
    B200_CHECK(cudaMemcpyAsync(c.pin_st + (j.issued & 1) * DB_MAX_WINDOWS, c.st, (size_t)j.W * sizeof(DecodeState), cudaMemcpyDeviceToHost, c.stream));
    B200_CHECK(cudaEventRecord(c.ev[j.issued & 1], c.stream));
    ++j.issued;
}

// Windows are spread over up to B200_DECODE_LANES (default 8) concurrent lanes; a lane takes ceil(n / lanes) windows per batched
// step (B200_BATCH_WINDOWS caps it).  Up to 8 windows that is one window per lane - measured faster than one wide batch while
// the step is latency bound - beyond that the lanes' batches grow.
int decode_windows_batch(const int* windows, int n_windows, const int* initial_tokens, int n_initial, int beam_size, int n_group,
                         float temperature, unsigned long long seed, int sample_len,
                         int without_timestamps, int max_initial_timestamp_index, int* out_tokens, int* out_lengths,
                         float* out_sum_logprobs, float* out_no_speech, int* out_steps) {
    State& s = S();
    const DecodeSpec spec = decode_spec();
    const int nb = beam_size > 0 ? beam_size : (n_group > 0 ? n_group : 1), cand = nb;
    static const int max_lanes = [] { const char* e = getenv("B200_DECODE_LANES"); const int v = e ? atoi(e) : MAX_DECODE_LANES; return v < 1 ? 1 : (v > MAX_DECODE_LANES ? MAX_DECODE_LANES : v); }();
    const int w_max = batch_max_windows(nb);
    int sot_index = -1;
    for (int i = 0; i < n_initial; ++i) if (initial_tokens[i] == spec.sot) sot_index = i;   // tokens.index(sot) (:617)
    cudaStream_t main_stream = s.stream;
    cudaEvent_t ev_fork = nullptr, ev_join[MAX_DECODE_LANES] = {nullptr};
    B200_CHECK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    for (auto& e : ev_join) B200_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    int total_steps = 0;
    for (int w0 = 0; w0 < n_windows;) {
        const int left = n_windows - w0;
        const int lanes = std::min(max_lanes, left);
        const int per_lane = std::min(w_max, (left + lanes - 1) / lanes);
        const int n_ctas = grid_ctas() / lanes;
        BatchJob job[MAX_DECODE_LANES];
        int n_jobs = 0, w = w0;
        for (int i = 0; i < lanes && w < n_windows; ++i) {
            const int W = std::min(per_lane, n_windows - w);
            if (!ensure_lane(i, W * nb)) return total_steps;
            job[n_jobs] = BatchJob{i, W, nb, beam_size > 0 ? nb + 1 : 1, n_initial, sample_len, sot_index, n_ctas, windows + w};
            job[n_jobs].w0 = w;
            ++n_jobs; w += W;
        }
        auto fork = [&] {
            B200_CHECK(cudaEventRecord(ev_fork, main_stream));
            for (int i = 1; i < n_jobs; ++i) B200_CHECK(cudaStreamWaitEvent(g_lane[i].stream, ev_fork, 0));
        };
        auto join = [&] {
            for (int i = 1; i < n_jobs; ++i) { B200_CHECK(cudaEventRecord(ev_join[i], g_lane[i].stream)); B200_CHECK(cudaStreamWaitEvent(main_stream, ev_join[i], 0)); }
        };
        // ---- decode state, token histories, slot tables; then the prompt ----
        DecodeState h{};
        h.L = n_initial; h.pos = n_initial - 1; h.sample_begin = n_initial; h.sample_len = sample_len;
        h.beam_mode = beam_size > 0; h.without_timestamps = without_timestamps; h.max_initial_ts = max_initial_timestamp_index;
        h.suppress_blank = 1; h.no_speech_prob = NAN;
        h.temperature = beam_size > 0 ? 0.f : temperature; h.seed_lo = (unsigned)seed; h.seed_hi = (unsigned)(seed >> 32);
        std::vector<DecodeState> hs0(DB_MAX_WINDOWS, h);
        {
            // all beams hold the same tokens (decoding.py:761), so the prompt runs as n_initial one-row-per-window steps into the
            // window's first cache slot: a causal prefill over n rows IS n steps, and the step kernel streams each weight once
            // per position
            StageTimer t(ST_DECODER256);
            fork();
            for (int i = 0; i < n_jobs; ++i) {
                BatchJob& j = job[i];
                BatchCtx& c = g_lane[j.lane];
                for (int q = 0; q < j.W; ++q) hs0[q].stream = (unsigned)j.windows[q];       // every window draws from its own random stream
                B200_CHECK(cudaMemcpyAsync(c.st, hs0.data(), (size_t)j.W * sizeof(DecodeState), cudaMemcpyHostToDevice, c.stream));
                B200_CHECK(cudaStreamSynchronize(c.stream));                            // (hs0 is rewritten for the next lane)
                B200_CHECK(cudaMemcpyAsync(c.d_init, initial_tokens, (size_t)n_initial * sizeof(int), cudaMemcpyHostToDevice, c.stream));
                batch_init_tokens_kernel<<<1, 1024, 0, c.stream>>>(c.tokens, c.d_init, n_initial, j.W, nb, spec.eot, c.table);
                B200_LAUNCH_CHECK();
                for (int p = 0; p < n_initial; ++p) {
                    DbArgs a = step_args(j, true, p);
                    a.no_vocab = !(p == n_initial - 1 || p == sot_index);
                    db_launch(a, j.n_ctas, c.stream);
                    if (p == sot_index) no_speech_prob_batch(c.logits, s.V, s.V, spec.no_speech, c.st, j.W, c.stream);   // logits at the sot position (:716-720)
                }
            }
            join();
        }
        {
            StageTimer t(ST_SAMPLING);
            fork();
            for (int i = 0; i < n_jobs; ++i) { launch_batch_sampling(job[i], true); job[i].steps = 1; }
            join();
        }
        {
            StageTimer t(ST_DECODER1);                                  // on the main stream: fork .. join of all lanes
            fork();
            // every lane advances BATCH_GRAPH_STEPS per round; completion is polled one round behind the issue front, so the GPU
            // never idles while the host synchronises
            bool any = true;
            while (any) {
                for (int i = 0; i < n_jobs; ++i)
                    while (!job[i].done && job[i].steps < sample_len && job[i].issued - job[i].checked < 2) lane_issue(job[i]);
                any = false;
                for (int i = 0; i < n_jobs; ++i) {
                    BatchJob& j = job[i];
                    if (j.done) continue;
                    if (j.checked < j.issued) {
                        BatchCtx& c = g_lane[j.lane];
                        B200_CHECK(cudaEventSynchronize(c.ev[j.checked & 1]));
                        const DecodeState* ps = c.pin_st + (j.checked & 1) * DB_MAX_WINDOWS;
                        bool all = true;
                        for (int q = 0; q < j.W; ++q) all = all && ps[q].done != 0;
                        ++j.checked;
                        if (all || (j.steps >= sample_len && j.checked == j.issued)) j.done = true; else any = true;
                    } else j.done = true;                               // nothing in flight and nothing left to issue
                }
            }
            join();
        }
        // ---- results ----
        for (int i = 0; i < n_jobs; ++i) {
            BatchJob& j = job[i];
            BatchCtx& c = g_lane[j.lane];
            std::vector<DecodeState> hs(j.W);
            std::vector<int> tok((size_t)j.W * nb * DEC_TOK_LD), fin((size_t)j.W * DEC_MAX_BEAMS * DEC_TOK_LD);
            B200_CHECK(cudaMemcpyAsync(hs.data(), c.st, hs.size() * sizeof(DecodeState), cudaMemcpyDeviceToHost, c.stream));
            B200_CHECK(cudaMemcpyAsync(tok.data(), c.tokens, tok.size() * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
            B200_CHECK(cudaMemcpyAsync(fin.data(), c.fin_tokens, fin.size() * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
            B200_CHECK(cudaStreamSynchronize(c.stream));
            if (const unsigned long long f = db_fault_word())
                report_step_fault(f);
            for (int q = 0; q < j.W; ++q) {
                const size_t o = (size_t)(j.w0 + q);
                const int n = emit_window(hs[q], tok.data() + (size_t)q * nb * DEC_TOK_LD, fin.data() + (size_t)q * DEC_MAX_BEAMS * DEC_TOK_LD, nb, n_initial,
                                          spec.eot, out_tokens + o * cand * DEC_TOK_LD, out_lengths + o * cand, out_sum_logprobs + o * cand,
                                          out_no_speech ? out_no_speech + o : nullptr);
                if (out_steps) out_steps[o] = n;
                total_steps += n;
            }
        }
        w0 = w;
    }
    cudaEventDestroy(ev_fork);
    for (auto& e : ev_join) cudaEventDestroy(e);
    return total_steps;
}

// stage timeline of the batched kernel (tools/step_timeline.py): enable allocates + clears the buffer, disable copies it out
int batch_timeline(int enable, unsigned long long* out, int cap_ctas) {
    State& s = S();
    const size_t n = (size_t)256 * DB_DBG_LD;                        // rows 0 .. n_sms - 1: step-kernel CTAs; rows 200 ..: sampling-kernel marks
    if (enable) {
        if (!g_dbg && !dev_alloc(&g_dbg, n)) return 0;
        B200_CHECK(cudaMemset(g_dbg, 0, n * sizeof(unsigned long long)));
        return 1;
    }
    if (!g_dbg) return 0;
    B200_CHECK(cudaDeviceSynchronize());
    const int n_ctas = std::min(cap_ctas, 256);
    B200_CHECK(cudaMemcpy(out, g_dbg, (size_t)n_ctas * DB_DBG_LD * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    dev_free(&g_dbg);
    return n_ctas;
}

}  // namespace b200
