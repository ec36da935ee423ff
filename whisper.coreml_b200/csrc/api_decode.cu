// C ABI Part 2: batched windows and the device-resident decode loop.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <tuple>
#include <vector>

#include "api_batch.cuh"
#include "decoder_mega.cuh"
#include "decoder_step.cuh"
#include "ops.cuh"
#include "sampling.cuh"
#include "gemm.cuh"
#include "state.cuh"
#include "timing.cuh"
#include "whisper_b200.h"

namespace b200 {

struct DecodeCtx {
    DecodeSpec spec;
    std::vector<uint8_t> h_suppress;
    uint8_t* d_suppress = nullptr; size_t suppress_cap = 0;
    DecodeState* st = nullptr;          // device
    int* tokens = nullptr;              // [8][449]
    int* fin_tokens = nullptr;          // [8][449]
    float* cand_lp = nullptr; int* cand_tok = nullptr;   // [8][9]
    SamplePartials* part = nullptr;
    float* ns_logits = nullptr;         // [V] logits at the sot position
    int* pin_done = nullptr;            // pinned host
    bool ready = false;
};
static DecodeCtx g_dc;

DecodeSpec decode_spec() { return g_dc.spec; }

static bool ensure_decode_ctx() {
    DecodeCtx& c = g_dc;
    State& s = S();
    if (c.spec.n_vocab != s.V || c.spec.eot < 0) { record_error("decode: call b200SetDecodeSpec (n_vocab %d) first", s.V); return false; }
    if (c.ready) return true;
    bool ok = true;
    ok &= dev_alloc(&c.st, 1, true);
    ok &= dev_alloc(&c.tokens, (size_t)DEC_MAX_BEAMS * DEC_TOK_LD, true);
    ok &= dev_alloc(&c.fin_tokens, (size_t)DEC_MAX_BEAMS * DEC_TOK_LD, true);
    ok &= dev_alloc(&c.cand_lp, (size_t)DEC_MAX_BEAMS * (DEC_MAX_BEAMS + 1));
    ok &= dev_alloc(&c.cand_tok, (size_t)DEC_MAX_BEAMS * (DEC_MAX_BEAMS + 1));
    ok &= dev_alloc(&c.ns_logits, (size_t)s.V);
    ok &= dev_alloc(&c.part, 1, true);
    if (!c.pin_done) ok &= cudaMallocHost((void**)&c.pin_done, 64) == cudaSuccess;
    c.ready = ok;
    return ok;
}

// x256[r] = tok_emb[tok[r]] + pos_emb[r] for r < n, zero rows after (decoder.py:202,214); mask = causal with
// columns >= n masked (decoder.py:212-213)
__global__ void prefill_inputs_kernel(const bf16* __restrict__ tok_emb, const float* __restrict__ pos_emb,
                                      const int* __restrict__ tokens, int n, int d, float* __restrict__ x, float* __restrict__ mask) {
    const int r = blockIdx.x;
    const int tok = r < n ? tokens[r] : 0;
    for (int c = threadIdx.x; c < d; c += blockDim.x)
        x[(long)r * d + c] = r < n ? __bfloat162float(tok_emb[(long)tok * d + c]) + pos_emb[(long)r * d + c] : 0.f;
    for (int c = threadIdx.x; c < PREFILL_CTX; c += blockDim.x)
        mask[r * PREFILL_CTX + c] = (c <= r && c < n) ? 0.f : -INFINITY;
}
__global__ void init_tokens_kernel(int* tokens, const int* initial, int n, int nb, int eot, int* table) {
    for (int i = threadIdx.x; i < nb * DEC_TOK_LD; i += blockDim.x) {
        const int p = i % DEC_TOK_LD;
        tokens[i] = p < n ? initial[p] : eot;
    }
    for (int i = threadIdx.x; i < nb * N_TEXT_CTX; i += blockDim.x) table[i] = 0;     // every beam reads the prefill from slot 0
}

struct AlignCtx {
    float *tmp = nullptr, *mat = nullptr, *neg = nullptr, *logits = nullptr, *probs = nullptr;
    bf16* rows = nullptr;
    int *path = nullptr, *targets = nullptr, *d_tok = nullptr;
    uint8_t* scratch = nullptr;
    size_t tmp_cap = 0;
    bool ready = false;
};
static AlignCtx g_al;

__global__ void negate_kernel(const float* __restrict__ in, float* __restrict__ out, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = -in[i];
}
__global__ void token_prob_kernel(const float* __restrict__ logits, long ld, int eot, const int* __restrict__ targets,
                                  float* __restrict__ out);

// ---- CUDA graphs of GRAPH_STEPS decoder1 iterations (embed -> blocks -> vocab -> sampling -> beam update) --------
// Every per-step quantity (text_offset, tokens, slot table, completion flag) lives in device memory, so one
// instantiated graph serves every step of every window that uses the same buffers.
constexpr int GRAPH_STEPS = 8;
struct StepGraph { cudaGraphExec_t exec = nullptr; long launches = 0; };
static std::map<std::tuple<int, int, const void*, const void*, const void*>, StepGraph> g_step_graphs;
void decode_clear_graphs() {
    for (auto& kv : g_step_graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    g_step_graphs.clear();
}

bool run_step_mega(int nb, int text_offset, const float* d_mask, const float* d_x_in, const MegaArgs* decode_fields);

static void launch_sampling(int nb, int k, bool shared_logits = false);
static void one_step(int nb, int k) {
    State& s = S();
    DecodeCtx& c = g_dc;
    if (mega_available()) {                                            // embedding .. logits in one persistent launch
        MegaArgs a{};
        a.tokens = c.tokens; a.d_pos = &c.st->pos; a.d_done = &c.st->done;
        // B200_MEGA_SAMPLING=1 runs the sampling tail inside the kernel too; measured slower (61 us vs ~55 us): with one warp
        // per scheduler and no L1 the serial beam update is latency bound there, the two stand-alone kernels are not
        static const bool fused = getenv("B200_MEGA_SAMPLING") && atoi(getenv("B200_MEGA_SAMPLING")) != 0;
        if (fused) { a.do_sampling = 1; a.k = k; a.st = c.st; a.spec = c.spec; a.sp = c.part; a.cand_lp = c.cand_lp; a.cand_tok = c.cand_tok; a.fin_tokens = c.fin_tokens; }
        run_step_mega(nb, 0, nullptr, nullptr, &a);
        if (!fused) launch_sampling(nb, k);
        return;
    }
    step_embed(s.tok_emb, s.pos_emb, c.tokens, DEC_TOK_LD, 0, &c.st->pos, &c.st->done, nb, s.d, s.sx, s.stream);
    run_step(nb, 0, nullptr, true, &c.st->pos, &c.st->done);
    launch_sampling(nb, k);
}

static StepGraph* step_graph(int nb, int k) {
    State& s = S();
    auto key = std::make_tuple(nb * 1000 + s.mega_ctas, k, (const void*)s.ck_ptr(s.cur_window, 0), (const void*)s.mkv, (const void*)s.slogits);
    auto it = g_step_graphs.find(key);
    if (it != g_step_graphs.end()) return &it->second;
    cudaGraph_t graph = nullptr;
    const long l0 = g_launch_count;
    if (cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    for (int i = 0; i < GRAPH_STEPS; ++i) one_step(nb, k);
    StepGraph g;
    g.launches = g_launch_count - l0;
    g_launch_count = l0;                                               // capture issued nothing; replays are counted per launch
    if (cudaStreamEndCapture(s.stream, &graph) != cudaSuccess || !graph) { cudaGetLastError(); return nullptr; }
    if (cudaGraphInstantiate(&g.exec, graph, 0) != cudaSuccess) { cudaGetLastError(); cudaGraphDestroy(graph); return nullptr; }
    cudaGraphDestroy(graph);
    return &g_step_graphs.emplace(key, g).first->second;
}

// shared_logits: every beam reads logits row 0 (the first step after the prompt: all beams hold the same tokens)
static void launch_sampling(int nb, int k, bool shared_logits) {
    State& s = S();
    DecodeCtx& c = g_dc;
    SampleArgs sa{};
    sa.logits = s.slogits; sa.ld_logits = shared_logits ? 0 : s.V; sa.tokens = c.tokens; sa.st = c.st; sa.spec = c.spec; sa.nb = nb; sa.k = k;
    sa.cand_lp = c.cand_lp; sa.cand_tok = c.cand_tok; sa.part = c.part;
    BeamUpdateArgs ba{};
    ba.part = c.part; ba.timestamp_begin = c.spec.timestamp_begin; ba.update = 1;
    ba.cand_lp = c.cand_lp; ba.cand_tok = c.cand_tok; ba.nb = nb; ba.k = k; ba.tokens = c.tokens; ba.table = s.table;
    ba.fin_tokens = c.fin_tokens; ba.st = c.st; ba.eot = c.spec.eot; ba.n_text_ctx = N_TEXT_CTX;
    static const bool split = getenv("B200_SAMPLING_SPLIT") && atoi(getenv("B200_SAMPLING_SPLIT")) != 0;   // two launches (the earlier form)
    if (split) { sample_partial(sa, s.stream); beam_update(ba, s.stream); }
    else sample_and_update(sa, ba, s.stream);
}

}  // namespace b200

using namespace b200;

extern "C" {

void b200SetDecodeSpec(int sot, int eot, int no_timestamps, int timestamp_begin, int no_speech, const int* suppress,
                       int n_suppress, const int* blank, int n_blank) {
    State& s = S();
    DecodeCtx& c = g_dc;
    if (!s.V) { record_error("b200SetDecodeSpec: load the decoder first (n_vocab unknown)"); return; }
    if (eot < 0 || eot >= s.V || timestamp_begin < 0 || timestamp_begin >= s.V) { record_error("b200SetDecodeSpec: eot %d / timestamp_begin %d outside the vocabulary (%d)", eot, timestamp_begin, s.V); return; }
    if (n_blank > 4) { record_error("b200SetDecodeSpec: %d blank tokens (at most 4 are supported)", n_blank); return; }
    if ((timestamp_begin + SAMPLE_TEXT_CHUNKS - 1) / SAMPLE_TEXT_CHUNKS > SAMPLE_CHUNK_TOKENS_MAX || s.V - timestamp_begin > SAMPLE_CHUNK_TOKENS_MAX) {
        record_error("b200SetDecodeSpec: vocabulary chunks of %d text / %d timestamp tokens exceed %d", (timestamp_begin + SAMPLE_TEXT_CHUNKS - 1) / SAMPLE_TEXT_CHUNKS,
                     s.V - timestamp_begin, SAMPLE_CHUNK_TOKENS_MAX);
        return;
    }
    use_device();
    B200_CHECK(cudaDeviceSynchronize());
    decode_clear_graphs(); batch_clear_graphs();                        // captured step graphs hold the old spec (and d_suppress) by value
    c.spec.sot = sot; c.spec.eot = eot; c.spec.no_timestamps = no_timestamps; c.spec.timestamp_begin = timestamp_begin;
    c.spec.no_speech = no_speech; c.spec.n_vocab = s.V;
    for (int i = 0; i < 4; ++i) c.spec.blank[i] = i < n_blank ? blank[i] : -1;
    c.h_suppress.assign((size_t)s.V, 0);
    for (int i = 0; i < n_suppress; ++i)
        if (suppress[i] >= 0 && suppress[i] < s.V) c.h_suppress[suppress[i]] = 1;
    // the flag buffer is kept and overwritten in place (captured graphs hold its address); it only moves when the vocabulary grows
    if ((!c.d_suppress || c.suppress_cap < (size_t)s.V) && !dev_alloc(&c.d_suppress, (size_t)s.V)) return;
    c.suppress_cap = std::max(c.suppress_cap, (size_t)s.V);
    B200_CHECK(cudaMemcpy(c.d_suppress, c.h_suppress.data(), (size_t)s.V, cudaMemcpyHostToDevice));
    c.spec.d_suppress = c.d_suppress;
}

void encoderPredictWindows(const float* d_mel, long total_frames, const int* seeks, int n_windows) {
    encoderPredictWindowsContent(d_mel, total_frames, total_frames, seeks, n_windows);
}

void encoderPredictWindowsContent(const float* d_mel, long total_frames, long content_frames, const int* seeks, int n_windows) {
    State& s = S();
    if (!s.enc_loaded) { record_error("encoderPredictWindows: encoder not loaded"); return; }
    if (n_windows < 1) return;
    use_device();
    B200_CHECK(cudaStreamSynchronize(cudaStreamLegacy));                // d_mel may have been produced on the caller's stream
    if (!ensure_encoder_capacity(n_windows)) return;
    B200_CHECK(cudaMemcpyAsync(s.d_seeks, seeks, (size_t)n_windows * sizeof(int), cudaMemcpyHostToDevice, s.stream));
    {
        StageTimer t(ST_ENCODER);
        run_encoder(d_mel, total_frames, content_frames, n_windows);
    }
    s.cur_window = 0;
    B200_CHECK(cudaStreamSynchronize(s.stream));
}

void crossKVPredictWindows(int n_windows) {
    State& s = S();
    if (!s.ckv_loaded || !s.enc_loaded) { record_error("crossKVPredictWindows: encoder / crossKV not loaded"); return; }
    if (n_windows < 1 || n_windows > s.n_windows) { record_error("crossKVPredictWindows: %d windows requested, %d encoded", n_windows, s.n_windows); return; }
    use_device();
    {
        StageTimer t(ST_CROSSKV);
        run_cross_kv(n_windows);
    }
    B200_CHECK(cudaStreamSynchronize(s.stream));
}

}  // extern "C"

namespace b200 {

// ---- decode lanes --------------------------------------------------------------------------------------------------------
// The persistent step kernel is latency bound (one token step is a chain of ~38 dependent stages), so independent windows
// are decoded CONCURRENTLY, each lane on its own stream with its own KV cache / decode state and its share of the SMs
// (n_sms / lanes CTAs per step kernel).  Lane 0 is the process-global state of the reference ABI; switching lanes swaps
// the lane-specific pointers of State / DecodeCtx with the parked copy, so every kernel-launching helper stays lane-agnostic.
struct Lane {
    cudaStream_t stream = nullptr;
    bf16* mkv = nullptr; int* table = nullptr; float* slogits = nullptr; uint2* mega_ll = nullptr; unsigned* mega_barrier = nullptr;
    DecodeState* st = nullptr; int* tokens = nullptr; int* fin_tokens = nullptr; float* cand_lp = nullptr; int* cand_tok = nullptr;
    SamplePartials* part = nullptr; float* ns_logits = nullptr; int* pin_done = nullptr;
    int cur_window = 0;
    bool ready = false;
};
static Lane g_parked[MAX_LANES];          // g_parked[i] holds lane i's pointers while another lane is active (entry 0 unused until a swap)
static int g_active_lane = 0;

static void lane_swap(Lane& l) {
    State& s = S();
    DecodeCtx& c = g_dc;
    std::swap(l.stream, s.stream); std::swap(l.mkv, s.mkv); std::swap(l.table, s.table); std::swap(l.slogits, s.slogits);
    std::swap(l.mega_ll, s.mega_ll); std::swap(l.mega_barrier, s.mega_barrier); std::swap(l.cur_window, s.cur_window);
    std::swap(l.st, c.st); std::swap(l.tokens, c.tokens); std::swap(l.fin_tokens, c.fin_tokens); std::swap(l.cand_lp, c.cand_lp);
    std::swap(l.cand_tok, c.cand_tok); std::swap(l.part, c.part); std::swap(l.ns_logits, c.ns_logits); std::swap(l.pin_done, c.pin_done);
}
static void activate_lane(int i) {
    if (i == g_active_lane) return;
    lane_swap(g_parked[g_active_lane]);    // park the active lane ...
    lane_swap(g_parked[i]);                // ... and bring lane i in
    g_active_lane = i;
}
size_t mega_ll_words_for(size_t d, size_t H);
static bool ensure_lane(int i) {           // allocate lane i > 0 (the active lane must be 0)
    Lane& l = g_parked[i];
    if (l.ready) return true;
    State& s = S();
    const size_t d = s.d;
    bool ok = cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking) == cudaSuccess;
    ok &= dev_alloc(&l.mkv, (size_t)2 * s.Ld * s.bs * N_TEXT_CTX * d, true);
    ok &= dev_alloc(&l.table, (size_t)STEP_MAX_BEAMS * N_TEXT_CTX, true);
    ok &= dev_alloc(&l.slogits, (size_t)STEP_MAX_BEAMS * s.V);
    ok &= dev_alloc(&l.mega_ll, mega_ll_words_for(d, s.H), true);
    ok &= dev_alloc(&l.mega_barrier, (size_t)4, true);
    if (ok) { const unsigned one = 1; B200_CHECK(cudaMemcpy(l.mega_barrier + 2, &one, sizeof(one), cudaMemcpyHostToDevice)); }
    ok &= dev_alloc(&l.st, 1, true);
    ok &= dev_alloc(&l.tokens, (size_t)DEC_MAX_BEAMS * DEC_TOK_LD, true);
    ok &= dev_alloc(&l.fin_tokens, (size_t)DEC_MAX_BEAMS * DEC_TOK_LD, true);
    ok &= dev_alloc(&l.cand_lp, (size_t)DEC_MAX_BEAMS * (DEC_MAX_BEAMS + 1));
    ok &= dev_alloc(&l.cand_tok, (size_t)DEC_MAX_BEAMS * (DEC_MAX_BEAMS + 1));
    ok &= dev_alloc(&l.ns_logits, (size_t)s.V);
    ok &= dev_alloc(&l.part, 1, true);
    ok &= cudaMallocHost((void**)&l.pin_done, 64) == cudaSuccess;
    l.ready = ok;
    return ok;
}
void decode_free_lanes() {                 // called when the decoder is closed (lane 0 must be active)
    activate_lane(0);
    for (int i = 1; i < MAX_LANES; ++i) {
        Lane& l = g_parked[i];
        if (l.stream) cudaStreamDestroy(l.stream);
        dev_free(&l.mkv); dev_free(&l.table); dev_free(&l.slogits); dev_free(&l.mega_ll); dev_free(&l.mega_barrier); dev_free(&l.st);
        dev_free(&l.tokens); dev_free(&l.fin_tokens); dev_free(&l.cand_lp); dev_free(&l.cand_tok); dev_free(&l.ns_logits); dev_free(&l.part);
        if (l.pin_done) cudaFreeHost(l.pin_done);
        l = Lane{};
    }
}

struct DecodeJob { int n_initial, nb, k, sample_len, sot_index, steps; bool done; StepGraph* graph; int issued = 0, checked = 0; };
// completion is polled one round behind the issue front: round r + 1 is already queued when the host waits for round r's flag,
// so the GPU never idles while the host synchronises and relaunches (rounds issued past the end are no-ops)
static cudaEvent_t g_round_ev[MAX_LANES][2];

// state + prefill + first sampling step of the ACTIVE lane, all on `st` (the prefill workspaces are shared by the lanes)
static void decode_begin(DecodeJob& j, const int* initial_tokens, int beam_size, int without_timestamps, int max_initial_timestamp_index) {
    State& s = S();
    DecodeCtx& c = g_dc;
    cudaStream_t st = s.stream;
    const int d = s.d, nb = j.nb, n_initial = j.n_initial;
    DecodeState h{};
    h.L = n_initial; h.pos = n_initial - 1; h.sample_begin = n_initial; h.sample_len = j.sample_len;
    h.beam_mode = beam_size > 0; h.without_timestamps = without_timestamps; h.max_initial_ts = max_initial_timestamp_index;
    h.suppress_blank = 1; h.no_speech_prob = NAN;
    B200_CHECK(cudaMemcpyAsync(c.st, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    int* d_init = c.fin_tokens;                                         // scratch before any sequence finishes
    B200_CHECK(cudaMemcpyAsync(d_init, initial_tokens, (size_t)n_initial * sizeof(int), cudaMemcpyHostToDevice, st));
    init_tokens_kernel<<<1, 256, 0, st>>>(c.tokens, d_init, n_initial, nb, c.spec.eot, s.table);
    B200_LAUNCH_CHECK();
    const bool by_steps = mega_available();
    if (by_steps) {
        // ---- prompt, all beams hold the same tokens (decoding.py:761): n_initial single-beam token steps of the persistent
        //      kernel into cache slot 0 (every beam's slot table points there).  A causal prefill over n rows IS n steps; the
        //      step kernel streams each weight once per position instead of launching ~50 mostly idle kernels per window.
        StageTimer t(ST_DECODER256);
        const int saved_ctas = s.mega_ctas;
        s.mega_ctas = 0;                                                // the prompt runs alone on the main stream: all SMs
        for (int p = 0; p < n_initial; ++p) {
            MegaArgs a{};
            a.tokens = c.tokens;
            a.no_vocab = !(p == n_initial - 1 || p == j.sot_index);
            run_step_mega(1, p, nullptr, nullptr, &a);
            if (p == j.sot_index) no_speech_prob(s.slogits, s.V, c.spec.no_speech, c.st, st);   // logits of the sot position (:716-720)
        }
        s.mega_ctas = saved_ctas;
    } else {   // ---- prefill once: all beams hold the same initial tokens (decoding.py:761) ----
        StageTimer t(ST_DECODER256);
        prefill_inputs_kernel<<<PREFILL_CTX, 256, 0, st>>>(s.tok_emb, s.pos_emb, d_init, n_initial, d, s.px, s.pmask);
        B200_LAUNCH_CHECK();
        run_prefill(0, false, n_initial);
        init_tokens_kernel<<<1, 256, 0, st>>>(c.tokens, d_init, n_initial, nb, c.spec.eot, s.table);   // run_prefill marked slot 0 only
        B200_LAUNCH_CHECK();
        StepGemv g{};
        g.nb = nb; g.w_frag = s.tok_emb_frag; g.N = s.V; g.K = d; g.x_f32 = s.pout + (size_t)(n_initial - 1) * d; g.ld_x = 0;
        g.out_f32 = s.slogits; g.ld_out = s.V;
        step_gemv(g, st);                                               // logits of the last prompt row, same for every beam
        if (j.sot_index >= 0) {
            g.nb = 1; g.x_f32 = s.pout + (size_t)j.sot_index * d; g.out_f32 = c.ns_logits;
            step_gemv(g, st);
            no_speech_prob(c.ns_logits, s.V, c.spec.no_speech, c.st, st);
        }
    }
    {
        StageTimer t(ST_SAMPLING);
        launch_sampling(nb, j.k, by_steps);
    }
    j.steps = 1; j.done = false; j.graph = nullptr;
}

// issue the next GRAPH_STEPS steps of the ACTIVE lane and the read-back of its completion flag (no host sync)
static void decode_issue(DecodeJob& j) {
    State& s = S();
    DecodeCtx& c = g_dc;
    cudaStream_t st = s.stream;
    if (j.steps == 1 && j.steps < j.sample_len) {                       // eager once: sets kernel attributes before any capture
        one_step(j.nb, j.k); ++j.steps;
        if (j.steps < j.sample_len) j.graph = step_graph(j.nb, j.k);
    }
    if (j.steps < j.sample_len) {
        // launches past sample_len / completion are no-ops: every kernel checks DecodeState::done first
        if (j.graph) { B200_CHECK(cudaGraphLaunch(j.graph->exec, st)); g_launch_count += j.graph->launches; }
        else for (int i = 0; i < GRAPH_STEPS; ++i) one_step(j.nb, j.k);
        j.steps += GRAPH_STEPS;
    }
    B200_CHECK(cudaMemcpyAsync(c.pin_done + (j.issued & 1), &c.st->done, sizeof(int), cudaMemcpyDeviceToHost, st));
    cudaEvent_t& ev = g_round_ev[g_active_lane][j.issued & 1];
    if (!ev) B200_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    B200_CHECK(cudaEventRecord(ev, st));
    ++j.issued;
}

// finalize (decoding.py:411-431 / :320-325) of the ACTIVE lane; its stream must be idle
static int decode_finish(const DecodeJob& j, int* out_tokens, int* out_lengths, float* out_sum_logprobs, float* out_no_speech) {
    State& s = S();
    DecodeCtx& c = g_dc;
    cudaStream_t st = s.stream;
    const int nb = j.nb, n_initial = j.n_initial;
    DecodeState h{};
    std::vector<int> tok((size_t)DEC_MAX_BEAMS * DEC_TOK_LD), fin((size_t)DEC_MAX_BEAMS * DEC_TOK_LD);
    B200_CHECK(cudaMemcpyAsync(&h, c.st, sizeof(h), cudaMemcpyDeviceToHost, st));
    B200_CHECK(cudaMemcpyAsync(tok.data(), c.tokens, tok.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
    B200_CHECK(cudaMemcpyAsync(fin.data(), c.fin_tokens, fin.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
    B200_CHECK(cudaStreamSynchronize(st));
    const int eot = c.spec.eot;
    int n_cand = 0;
    auto emit = [&](const int* seq, int len, float score) {
        int* dst = out_tokens + (size_t)n_cand * DEC_TOK_LD;
        for (int i = 0; i < DEC_TOK_LD; ++i) dst[i] = i < len ? seq[i] : eot;
        int l = 0;
        while (n_initial + l < len && seq[n_initial + l] != eot) ++l;  // tokens before the first EOT after sample_begin (:776-779)
        out_lengths[n_cand] = l; out_sum_logprobs[n_cand] = score; ++n_cand;
    };
    if (!h.beam_mode) {
        emit(tok.data(), h.L, h.sum_lp[0]);
    } else {
        for (int f = 0; f < h.n_finished; ++f) emit(fin.data() + (size_t)f * DEC_TOK_LD, h.fin_len[f], h.fin_score[f]);
        if (n_cand < nb) {                                              // not enough finished: add live beams, best first (:418-424)
            std::vector<int> order(nb);
            for (int i = 0; i < nb; ++i) order[i] = i;
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return h.sum_lp[a] > h.sum_lp[b]; });
            for (int i = 0; i < nb && n_cand < nb; ++i) emit(tok.data() + (size_t)order[i] * DEC_TOK_LD, h.L, h.sum_lp[order[i]]);
        }
    }
    for (int i = n_cand; i < nb; ++i) { out_lengths[i] = -1; out_sum_logprobs[i] = -INFINITY; }
    if (out_no_speech) *out_no_speech = h.no_speech_prob;
    return h.step;
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200DecodeWindows(const int* windows, int n_windows, const int* initial_tokens, int n_initial, int beam_size, int sample_len,
                      int without_timestamps, int max_initial_timestamp_index, int* out_tokens, int* out_lengths,
                      float* out_sum_logprobs, float* out_no_speech, int* out_steps) {
    State& s = S();
    DecodeCtx& c = g_dc;
    if (!s.dec1_loaded || !s.dec256_loaded || !s.ckv_loaded) { record_error("b200DecodeWindows: decoder256 / decoder1 / crossKV not loaded"); return 0; }
    const int nb = beam_size > 0 ? beam_size : 1;
    if (nb > s.bs) { record_error("b200DecodeWindows: %d beams but loadDecoder256 reserved %d cache slots", nb, s.bs); return 0; }
    if (n_initial < 1 || n_initial > PREFILL_CTX) { record_error("b200DecodeWindows: n_initial %d outside [1, 256]", n_initial); return 0; }
    if (n_windows < 1 || sample_len < 1) return 0;
    for (int w = 0; w < n_windows; ++w)
        if (windows[w] < 0 || windows[w] >= (s.ckv_cap > 0 ? s.ckv_cap : 1)) { record_error("b200DecodeWindows: window %d outside [0, %d)", windows[w], s.ckv_cap); return 0; }
    use_device();
    activate_lane(0);
    if (c.spec.n_vocab != s.V || c.spec.eot < 0) { record_error("decode: call b200SetDecodeSpec (n_vocab %d) first", s.V); return 0; }
    if (batch_available())                                              // every window of a batch advances in ONE step kernel
        return decode_windows_batch(windows, n_windows, initial_tokens, n_initial, beam_size, sample_len, without_timestamps,
                                    max_initial_timestamp_index, out_tokens, out_lengths, out_sum_logprobs, out_no_speech, out_steps);
    if (!ensure_decode_ctx()) return 0;
    // lanes: only the persistent step kernel can share the GPU between decodes; B200_DECODE_LANES=1 turns the overlap off
    static const int max_lanes = [] { const char* e = getenv("B200_DECODE_LANES"); const int v = e ? atoi(e) : MAX_LANES; return v < 1 ? 1 : (v > MAX_LANES ? MAX_LANES : v); }();
    int lanes = mega_available() ? std::min(max_lanes, n_windows) : 1;
    for (int i = 1; i < lanes; ++i) if (!ensure_lane(i)) { lanes = 1; break; }
    const int cand = beam_size > 0 ? nb : 1;
    int sot_index = -1;
    for (int i = 0; i < n_initial; ++i) if (initial_tokens[i] == c.spec.sot) sot_index = i;   // tokens.index(sot) (:617)
    cudaStream_t main_stream = s.stream;
    cudaEvent_t ev_fork = nullptr, ev_join[MAX_LANES] = {nullptr};
    if (lanes > 1) { cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming); for (int i = 1; i < lanes; ++i) cudaEventCreateWithFlags(&ev_join[i], cudaEventDisableTiming); }
    int total_steps = 0;
    for (int w0 = 0; w0 < n_windows; w0 += lanes) {
        const int n = std::min(lanes, n_windows - w0);
        DecodeJob job[MAX_LANES];
        // ---- prefill of every lane on the main stream (shared workspaces), then fork ----
        for (int i = 0; i < n; ++i) {
            activate_lane(i);
            cudaStream_t own = S().stream;
            S().stream = main_stream;
            S().cur_window = windows[w0 + i];
            job[i] = DecodeJob{n_initial, nb, beam_size > 0 ? nb + 1 : 1, sample_len, sot_index, 0, false, nullptr};
            decode_begin(job[i], initial_tokens, beam_size, without_timestamps, max_initial_timestamp_index);
            S().stream = own;
        }
        activate_lane(0);
        s.mega_ctas = n > 1 ? s.n_sms / n : 0;
        {
            StageTimer t(ST_DECODER1);                                  // on the main stream: fork .. join of all lanes
            if (n > 1) { B200_CHECK(cudaEventRecord(ev_fork, main_stream)); }
            for (int i = 1; i < n; ++i) { activate_lane(i); B200_CHECK(cudaStreamWaitEvent(S().stream, ev_fork, 0)); }
            // ---- steps: every lane advances GRAPH_STEPS per round; completion is polled once per round ----
            bool any = true;
            while (any) {
                for (int i = 0; i < n; ++i)
                    while (!job[i].done && job[i].steps < sample_len && job[i].issued - job[i].checked < 2) { activate_lane(i); decode_issue(job[i]); }
                any = false;
                for (int i = 0; i < n; ++i) {
                    if (job[i].done) continue;
                    activate_lane(i);
                    if (job[i].checked < job[i].issued) {
                        B200_CHECK(cudaEventSynchronize(g_round_ev[i][job[i].checked & 1]));
                        const int flag = g_dc.pin_done[job[i].checked & 1];
                        ++job[i].checked;
                        if (flag != 0 || (job[i].steps >= sample_len && job[i].checked == job[i].issued)) job[i].done = true; else any = true;
                    } else job[i].done = true;                          // nothing in flight and nothing left to issue
                }
            }
            for (int i = 1; i < n; ++i) { activate_lane(i); B200_CHECK(cudaEventRecord(ev_join[i], S().stream)); B200_CHECK(cudaStreamWaitEvent(main_stream, ev_join[i], 0)); }
            activate_lane(0);
        }
        s.mega_ctas = 0;
        for (int i = 0; i < n; ++i) {
            activate_lane(i);
            const size_t o = (size_t)(w0 + i);
            const int steps = decode_finish(job[i], out_tokens + o * cand * DEC_TOK_LD, out_lengths + o * cand, out_sum_logprobs + o * cand,
                                            out_no_speech ? out_no_speech + o : nullptr);
            if (out_steps) out_steps[o] = steps;
            total_steps += steps;
        }
        activate_lane(0);
    }
    if (ev_fork) cudaEventDestroy(ev_fork);
    for (int i = 1; i < MAX_LANES; ++i) if (ev_join[i]) cudaEventDestroy(ev_join[i]);
    return total_steps;
}

int b200DecodeWindow(const int* initial_tokens, int n_initial, int beam_size, int sample_len, int without_timestamps,
                     int max_initial_timestamp_index, int* out_tokens, int* out_lengths, float* out_sum_logprobs,
                     float* out_no_speech) {
    activate_lane(0);
    const int w = S().cur_window;
    return b200DecodeWindows(&w, 1, initial_tokens, n_initial, beam_size, sample_len, without_timestamps, max_initial_timestamp_index,
                             out_tokens, out_lengths, out_sum_logprobs, out_no_speech, nullptr);
}

void decoder1StepFused(const int* tokens_hist, int n_hist, int sample_begin, int text_offset, int without_timestamps,
                       int max_initial_timestamp_index, float* out_logprob, int* out_token) {
    State& s = S();
    DecodeCtx& c = g_dc;
    if (!s.dec1_loaded || !s.dec256_loaded || !s.ckv_loaded) { record_error("decoder1StepFused: decoders not loaded"); return; }
    if (n_hist < 1 || n_hist > N_TEXT_CTX || text_offset != n_hist - 1) { record_error("decoder1StepFused: n_hist %d / text_offset %d", n_hist, text_offset); return; }
    use_device();
    if (!ensure_decode_ctx()) return;
    cudaStream_t st = s.stream;
    const int nb = s.bs, k = nb + 1;
    std::vector<int> rows((size_t)DEC_MAX_BEAMS * DEC_TOK_LD, c.spec.eot);
    for (int b = 0; b < nb; ++b) memcpy(&rows[(size_t)b * DEC_TOK_LD], tokens_hist + (size_t)b * n_hist, (size_t)n_hist * sizeof(int));
    B200_CHECK(cudaMemcpyAsync(c.tokens, rows.data(), rows.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    DecodeState h{};
    h.L = n_hist; h.pos = text_offset; h.sample_begin = sample_begin; h.sample_len = 1 << 30; h.beam_mode = 1;
    h.without_timestamps = without_timestamps; h.max_initial_ts = max_initial_timestamp_index; h.suppress_blank = 1;
    B200_CHECK(cudaMemcpyAsync(c.st, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    step_embed(s.tok_emb, s.pos_emb, c.tokens, DEC_TOK_LD, text_offset, nullptr, nullptr, nb, s.d, s.sx, st);
    run_step(nb, text_offset, nullptr, true, nullptr, nullptr);
    SampleArgs sa{};
    sa.logits = s.slogits; sa.ld_logits = s.V; sa.tokens = c.tokens; sa.st = c.st; sa.spec = c.spec; sa.nb = nb; sa.k = k;
    sa.cand_lp = c.cand_lp; sa.cand_tok = c.cand_tok; sa.part = c.part;
    sample_partial(sa, st);
    BeamUpdateArgs ba{};
    ba.part = c.part; ba.timestamp_begin = c.spec.timestamp_begin; ba.update = 0;
    ba.cand_lp = c.cand_lp; ba.cand_tok = c.cand_tok; ba.nb = nb; ba.k = k; ba.tokens = c.tokens; ba.table = s.table;
    ba.fin_tokens = c.fin_tokens; ba.st = c.st; ba.eot = c.spec.eot; ba.n_text_ctx = N_TEXT_CTX;
    beam_update(ba, st);
    B200_CHECK(cudaMemcpyAsync(out_logprob, c.cand_lp, (size_t)nb * k * sizeof(float), cudaMemcpyDeviceToHost, st));
    B200_CHECK(cudaMemcpyAsync(out_token, c.cand_tok, (size_t)nb * k * sizeof(int), cudaMemcpyDeviceToHost, st));
    B200_CHECK(cudaStreamSynchronize(st));
}

int b200AlignTokens(const int* tokens, int n_tokens, int n_skip, int num_frames, int medfilt_width, int* out_i, int* out_j,
                    float* out_matrix, float* out_text_token_probs) {
    State& s = S();
    AlignCtx& a = g_al;
    if (!s.dec256_loaded || !s.ckv_loaded) { record_error("b200AlignTokens: decoder256 / crossKV not loaded"); return 0; }
    if (s.n_align < 1) { record_error("b200AlignTokens: loadDecoder256 was called with n_alignment_head = 0"); return 0; }
    const int F = num_frames / 2, n_rows = n_tokens - 1 - n_skip, n_text = n_tokens - n_skip - 2;
    if (n_tokens > PREFILL_CTX || n_rows < 1 || F < 1 || F > N_AUDIO_CTX) { record_error("b200AlignTokens: n_tokens %d n_skip %d num_frames %d", n_tokens, n_skip, num_frames); return 0; }
    use_device();
    cudaStream_t st = s.stream;
    const size_t need = (size_t)2 * s.n_align * PREFILL_CTX * N_AUDIO_CTX;
    if (!a.ready || a.tmp_cap < need) {
        bool ok = true;
        ok &= dev_alloc(&a.tmp, need); a.tmp_cap = need;
        ok &= dev_alloc(&a.mat, (size_t)PREFILL_CTX * N_AUDIO_CTX); ok &= dev_alloc(&a.neg, (size_t)PREFILL_CTX * N_AUDIO_CTX);
        ok &= dev_alloc(&a.logits, (size_t)PREFILL_CTX * ((s.V + 7) / 8 * 8)); ok &= dev_alloc(&a.probs, (size_t)PREFILL_CTX);
        ok &= dev_alloc(&a.rows, (size_t)PREFILL_CTX * s.d); ok &= dev_alloc(&a.path, (size_t)2 * (PREFILL_CTX + N_AUDIO_CTX) + 1);
        ok &= dev_alloc(&a.targets, (size_t)PREFILL_CTX); ok &= dev_alloc(&a.d_tok, (size_t)PREFILL_CTX);
        ok &= dev_alloc(&a.scratch, dtw_scratch_bytes(PREFILL_CTX, N_AUDIO_CTX));
        a.ready = ok;
        if (!ok) return 0;
    }
    StageTimer timer(ST_ALIGN);
    B200_CHECK(cudaMemcpyAsync(a.d_tok, tokens, (size_t)n_tokens * sizeof(int), cudaMemcpyHostToDevice, st));
    prefill_inputs_kernel<<<PREFILL_CTX, 256, 0, st>>>(s.tok_emb, s.pos_emb, a.d_tok, n_tokens, s.d, s.px, s.pmask);
    B200_LAUNCH_CHECK();
    run_prefill(0, true, n_tokens);                                     // model(tokens[None]) (timing.py:185, model.py:110-119)
    alignment_matrix_dev(s.pchw, s.n_align, PREFILL_CTX, n_tokens, F, n_skip, medfilt_width, a.tmp, a.mat, false, st);
    negate_kernel<<<cdiv(n_rows * F, 256), 256, 0, st>>>(a.mat, a.neg, (long)n_rows * F);   // dtw(-matrix) (timing.py:205)
    B200_LAUNCH_CHECK();
    const int cap = n_rows + F;
    B200_CHECK(cudaMemsetAsync(a.path + 2 * cap, 0, sizeof(int), st));
    dtw_dev(a.neg, n_rows, F, a.path, a.path + cap, a.path + 2 * cap, a.scratch, st);
    if (out_text_token_probs && n_text > 0) {                           // token_probs over [:eot] (timing.py:187-190)
        f32_to_bf16(s.pout + (size_t)n_skip * s.d, a.rows, (long)n_text * s.d, st);
        GemmParams g = gemm_plain(a.rows, s.tok_emb, a.logits, n_text, s.V, s.d);
        g.c_fp32 = 1; g.ldc = (s.V + 7) / 8 * 8;                        // padded rows keep the epilogue stores vectorised
        gemm_tcgen05(g, st);
        B200_CHECK(cudaMemcpyAsync(a.targets, tokens + n_skip + 1, (size_t)n_text * sizeof(int), cudaMemcpyHostToDevice, st));
        token_prob_kernel<<<n_text, 1024, 0, st>>>(a.logits, (s.V + 7) / 8 * 8, g_dc.spec.eot >= 0 ? g_dc.spec.eot : s.V, a.targets, a.probs);
        B200_LAUNCH_CHECK();
        B200_CHECK(cudaMemcpyAsync(out_text_token_probs, a.probs, (size_t)n_text * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    int len = 0;
    B200_CHECK(cudaMemcpyAsync(&len, a.path + 2 * cap, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (out_matrix) B200_CHECK(cudaMemcpyAsync(out_matrix, a.mat, (size_t)n_rows * F * sizeof(float), cudaMemcpyDeviceToHost, st));
    B200_CHECK(cudaStreamSynchronize(st));
    if (len > 0 && len <= cap) {
        B200_CHECK(cudaMemcpy(out_i, a.path, (size_t)len * sizeof(int), cudaMemcpyDeviceToHost));
        B200_CHECK(cudaMemcpy(out_j, a.path + cap, (size_t)len * sizeof(int), cudaMemcpyDeviceToHost));
    }
    return len;
}

void b200GetStageTimes(float* out_ms7, int reset) { stage_times(out_ms7, reset != 0); }

}  // extern "C"
