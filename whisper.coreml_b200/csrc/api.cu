// C ABI of libwhisper_b200.so, Part 1: the plugin surface of wangchou/whisper.coreml
// (coreml/coreml.h:5-31, bodies coreml/coreml.mm) re-implemented on sm_100a.
#include <string.h>
#include <map>

#include <stdlib.h>

#include "api_batch.cuh"
#include "gemm.cuh"
#include "ops.cuh"
#include "state.cuh"
#include "whisper_b200.h"

namespace b200 {

int take_errors(char*, int);
void gemm_clear_map_cache();
void attention_clear_map_cache();
void decode_clear_graphs();
void encoder_clear_graphs();

State& S() { static State s; return s; }

void use_device() {
    State& s = S();
    B200_CHECK(cudaSetDevice(s.device));
    if (!s.stream) B200_CHECK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
}

// =================================================================================================
// encoder (whisper/encoder.py:103-136)
// =================================================================================================
bool ensure_encoder_capacity(int W) {
    State& s = S();
    if (W <= s.w_cap) return true;
    const size_t M = (size_t)W * N_AUDIO_CTX, d = s.d;
    bool ok = true;
    ok &= dev_alloc(&s.melrows, (size_t)W * 3002 * s.cpad, true);
    ok &= dev_alloc(&s.h1, (size_t)W * 3002 * d, true);               // frame rows 0 / 3001 stay zero
    ok &= dev_alloc(&s.x, M * d);
    ok &= dev_alloc(&s.y, M * d);
    ok &= dev_alloc(&s.qkv, M * 3 * d);
    ok &= dev_alloc(&s.att, M * d);
    ok &= dev_alloc(&s.hid, M * 4 * d);
    bf16* new_xa = nullptr;
    ok &= dev_alloc(&new_xa, M * d);
    if (ok && s.xa && s.w_cap > 0)
        B200_CHECK(cudaMemcpy(new_xa, s.xa, (size_t)s.w_cap * N_AUDIO_CTX * d * sizeof(bf16), cudaMemcpyDeviceToDevice));
    dev_free(&s.xa);
    s.xa = new_xa;
    ok &= dev_alloc(&s.d_seeks, (size_t)W);
    gemm_clear_map_cache(); attention_clear_map_cache(); decode_clear_graphs(); encoder_clear_graphs();
    if (ok) s.w_cap = W;
    return ok;
}

static GemmParams lin(const bf16* A, int M, int K, const bf16* W, const float* bias, int N, void* C, bool c_fp32) {
    GemmParams p = gemm_plain(A, W, C, M, N, K);
    p.bias = bias; p.c_fp32 = c_fp32 ? 1 : 0;
    return p;
}

// Everything behind mel_to_rows reads and writes the encoder's own workspaces only, so the ~290 launches of a batch of W windows
// are captured once per W (after one eager pass that creates the tensor maps and sets the kernel attributes) and replayed as a
// CUDA graph: the kernels are 9 - 60 us long and the gaps between dependent launches add up (B200_ENCODER_GRAPH=0 turns it off).
struct EncGraph { cudaGraphExec_t exec = nullptr; long launches = 0; bool seen = false; };
static std::map<int, EncGraph> g_enc_graphs;
void encoder_clear_graphs() {
    for (auto& kv : g_enc_graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    g_enc_graphs.clear();
}
static void encoder_layers(int W);

void run_encoder(const float* d_mel, long total_frames, long valid_frames, int W) {
    State& s = S();
    cudaStream_t st = s.stream;
    mel_to_rows(d_mel, total_frames, valid_frames, s.d_seeks, W, s.n_mels, s.cpad, s.melrows, st);
    static const bool use_graph = !(getenv("B200_ENCODER_GRAPH") && atoi(getenv("B200_ENCODER_GRAPH")) == 0);
    EncGraph& g = g_enc_graphs[W];
    if (use_graph && g.exec) {
        B200_CHECK(cudaGraphLaunch(g.exec, st));
        g_launch_count += g.launches;
    } else if (use_graph && g.seen) {
        cudaGraph_t graph = nullptr;
        const long l0 = g_launch_count;
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            encoder_layers(W);
            g.launches = g_launch_count - l0;
            if (cudaStreamEndCapture(st, &graph) == cudaSuccess && graph && cudaGraphInstantiate(&g.exec, graph, 0) == cudaSuccess) {
                g_launch_count = l0;
                B200_CHECK(cudaGraphLaunch(g.exec, st));
                g_launch_count += g.launches;
            } else {
                cudaGetLastError();
                g.exec = nullptr;
                g_launch_count = l0;
                encoder_layers(W);
            }
            if (graph) cudaGraphDestroy(graph);
        } else {
            cudaGetLastError();
            encoder_layers(W);
        }
    } else {
        encoder_layers(W);
        g.seen = true;
    }
    s.n_windows = W;
}

static void encoder_layers(int W) {
    State& s = S();
    const int d = s.d, M = W * N_AUDIO_CTX;
    cudaStream_t st = s.stream;
    {   // conv1 (k3, p1) + GELU: three accumulating passes over row-shifted views (encoder.py:124)
        GemmParams p{};
        for (int k = 0; k < 3; ++k) p.A[k] = s.melrows + (size_t)k * s.cpad;
        p.num_a_maps = 3; p.kblocks_per_map = s.cpad / 64; p.a_inner = s.cpad; p.a_row_stride = s.cpad;
        p.a_batch_stride = (long)3002 * s.cpad; p.rows_per_batch = N_FRAMES; p.batch = W;
        p.B = s.conv1_w; p.ldb = 3 * s.cpad; p.N = d; p.K = 3 * s.cpad; p.bias = s.conv1_b; p.gelu = 1;
        p.C = s.h1; p.c_fp32 = 0; p.ldc = d; p.c_batch_rows = 3002; p.c_row0 = 1; p.add_rows = 1;
        gemm_tcgen05(p, st);
    }
    {   // conv2 (k3, s2, p1) + GELU + positional embedding (encoder.py:125-128)
        GemmParams p{};
        for (int k = 0; k < 3; ++k) p.A[k] = s.h1 + (size_t)k * d;
        p.num_a_maps = 3; p.kblocks_per_map = d / 64; p.a_inner = d; p.a_row_stride = 2L * d;
        p.a_batch_stride = (long)3002 * d; p.rows_per_batch = N_AUDIO_CTX; p.batch = W;
        p.B = s.conv2_w; p.ldb = 3L * d; p.N = d; p.K = 3 * d; p.bias = s.conv2_b; p.gelu = 1;
        p.add = s.pos; p.add_rows = N_AUDIO_CTX; p.ld_add = d;
        p.C = s.x; p.c_fp32 = 1; p.ldc = d; p.c_batch_rows = N_AUDIO_CTX; p.c_row0 = 0;
        gemm_tcgen05(p, st);
    }
    for (int l = 0; l < s.Le; ++l) {                                   // encoder.py:61-80
        const EncLayer& L = s.enc_layers[l];
        layernorm(s.x, L.attn_ln_w, L.attn_ln_b, 1e-7f, s.y, nullptr, M, d, st);
        gemm_tcgen05(lin(s.y, M, d, L.qkv_w, L.qkv_b, 3 * d, s.qkv, false), st);
        AttnParams a{};
        a.Q = s.qkv; a.K = s.qkv + d; a.V = s.qkv + 2 * d;
        a.ldq = a.ldk = a.ldv = 3L * d; a.q_head_stride = a.k_head_stride = a.v_head_stride = 64;
        a.q_batch_stride = a.k_batch_stride = a.v_batch_stride = (long)N_AUDIO_CTX * 3 * d;
        a.O = s.att; a.ldo = d; a.o_head_stride = 64; a.o_batch_stride = (long)N_AUDIO_CTX * d;
        a.n_q = a.n_k = N_AUDIO_CTX; a.n_head = s.H; a.batch = W;
        attention_tc(a, st);
        GemmParams po = lin(s.att, M, d, L.out_w, L.out_b, d, s.x, true);
        po.add = s.x; po.add_rows = M; po.ld_add = d;
        gemm_tcgen05(po, st);
        layernorm(s.x, L.mlp_ln_w, L.mlp_ln_b, 1e-7f, s.y, nullptr, M, d, st);
        GemmParams p1 = lin(s.y, M, d, L.mlp1_w, L.mlp1_b, 4 * d, s.hid, false);
        p1.gelu = 1;
        gemm_tcgen05(p1, st);
        GemmParams p2 = lin(s.hid, M, 4 * d, L.mlp2_w, L.mlp2_b, d, s.x, true);
        p2.add = s.x; p2.add_rows = M; p2.ld_add = d;
        gemm_tcgen05(p2, st);
    }
    layernorm(s.x, s.ln_post_w, s.ln_post_b, 1e-7f, s.xa, nullptr, M, d, st);   // encoder.py:133-134
}

// =================================================================================================
// crossKV (whisper/decoder.py:172-187): K without bias, V with bias, stored head-major per window
// =================================================================================================
void run_cross_kv(int W) {
    State& s = S();
    if (W > s.ckv_cap) {
        if (!dev_alloc(&s.ckv, (size_t)W * s.ckv_window_elems())) return;
        if (!dev_alloc(&s.ckv_frag, (size_t)W * s.ckv_frag_window_elems(), true)) return;   // pad keys 1500..1503 stay zero
        s.ckv_cap = W;
        gemm_clear_map_cache(); attention_clear_map_cache(); decode_clear_graphs(); encoder_clear_graphs();
    }
    const int d = s.d;
    GemmParams p{};
    p.A[0] = s.xa; p.num_a_maps = 1; p.kblocks_per_map = d / 64; p.a_inner = d; p.a_row_stride = d;
    p.a_batch_stride = (long)N_AUDIO_CTX * d; p.rows_per_batch = N_AUDIO_CTX; p.batch = W;
    p.B = s.ckv_wt; p.ldb = d; p.N = 2 * s.Ld * d; p.K = d; p.bias = s.ckv_b;
    p.C = s.ckv; p.c_fp32 = 0; p.ldc = 64; p.c_split = 1; p.c_split_stride = (long)N_AUDIO_CTX * 64;
    p.c_batch_rows = 2 * s.Ld * s.H * N_AUDIO_CTX; p.c_row0 = 0; p.add_rows = 1;
    p.C2 = s.ckv_frag; p.c2_batch_stride = (long)s.ckv_frag_window_elems(); p.c2_heads = s.H; p.c2_keys = CROSS_KEYS_PAD;
    gemm_tcgen05(p, s.stream);
}

// =================================================================================================
// decoder256 prefill (whisper/decoder.py:261-329 with qk_mask.shape[0] == 256; coreml.mm:279-327)
// =================================================================================================
struct PermuteArgs { int src[STEP_MAX_BEAMS]; };
__global__ void permute_table_kernel(int* table, PermuteArgs pa, int bs, int n) {
    __shared__ int stage[STEP_MAX_BEAMS * N_TEXT_CTX];
    for (int i = threadIdx.x; i < bs * N_TEXT_CTX; i += blockDim.x) stage[i] = table[i];
    __syncthreads();
    for (int i = threadIdx.x; i < bs * n; i += blockDim.x) {
        const int b = i / n, p = i % n;
        table[b * N_TEXT_CTX + p] = stage[pa.src[b] * N_TEXT_CTX + p];
    }
}

__global__ void fill_table_kernel(int* table, int beam, int n, int value) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) table[beam * N_TEXT_CTX + i] = value;
}

// rows < 256: only the first `rows` positions are computed.  The reference computes and stores all 256 rows, pad rows
// included (decoder.py:214, coreml.mm:313-326), but rows >= n_ctx of x / CHW are sliced away (decoder.py:236-237) and their
// K/V rows are masked by text_offset until the step that overwrites them, so the device-resident decode loop skips them.
void run_prefill(int beam_idx, bool want_chw, int rows) {
    State& s = S();
    const int d = s.d, M = rows < 1 ? 1 : (rows > PREFILL_CTX ? PREFILL_CTX : rows);
    cudaStream_t st = s.stream;
    for (int l = 0; l < s.Ld; ++l) {
        const DecLayer& L = s.dec_layers[l];
        layernorm(s.px, L.attn_ln_w, L.attn_ln_b, 1e-5f, s.py, nullptr, M, d, st);
        gemm_tcgen05(lin(s.py, M, d, L.qkv.w, L.qkv.b, 3 * d, s.pqkv, false), st);
        // the 256 K/V rows go to slot `beam_idx` of the 448-row cache (coreml.mm:313-326); pad rows included
        copy_rows_bf16(s.pqkv + d, 3L * d, s.mk_ptr(l) + (size_t)beam_idx * N_TEXT_CTX * d, d, M, d, st);
        copy_rows_bf16(s.pqkv + 2 * d, 3L * d, s.mv_ptr(l) + (size_t)beam_idx * N_TEXT_CTX * d, d, M, d, st);
        AttnParams a{};
        a.Q = s.pqkv; a.K = s.pqkv + d; a.V = s.pqkv + 2 * d; a.ldq = a.ldk = a.ldv = 3L * d;
        a.q_head_stride = a.k_head_stride = a.v_head_stride = 64;
        a.O = s.patt; a.ldo = d; a.o_head_stride = 64;
        a.n_q = a.n_k = M; a.n_head = s.H; a.batch = 1; a.mask = s.pmask; a.ld_mask = PREFILL_CTX;
        attention_simt(a, st);
        GemmParams po = lin(s.patt, M, d, L.attn_out.w, L.attn_out.b, d, s.px, true);
        po.add = s.px; po.add_rows = M; po.ld_add = d;
        gemm_tcgen05(po, st);
        layernorm(s.px, L.cross_ln_w, L.cross_ln_b, 1e-5f, s.py, nullptr, M, d, st);
        gemm_tcgen05(lin(s.py, M, d, L.cross_q.w, L.cross_q.b, d, s.pq, false), st);
        AttnParams c{};
        c.Q = s.pq; c.ldq = d; c.q_head_stride = 64;
        c.K = s.ck_ptr(s.cur_window, l); c.V = s.cv_ptr(s.cur_window, l);
        c.ldk = c.ldv = 64; c.k_head_stride = c.v_head_stride = (long)N_AUDIO_CTX * 64;
        c.O = s.patt; c.ldo = d; c.o_head_stride = 64;
        c.n_q = M; c.n_k = N_AUDIO_CTX; c.n_head = s.H; c.batch = 1;
        // no mask: the 256 x 1500 cross-attention runs on the tensor cores (tcgen05 flash attention); the raw QK of the alignment
        // heads (decoder.py:306-308) comes from a small kernel of its own
        attention_tc(c, st);
        if (want_chw && s.n_align > 0) {
            c.qk_dump = s.pchw; c.dump_slot = s.d_dump_slot + l * s.H; c.dump_ld = N_AUDIO_CTX;
            c.dump_slot_stride = (long)PREFILL_CTX * N_AUDIO_CTX;
            attention_qk_dump(c, st);
        }
        GemmParams pc = lin(s.patt, M, d, L.cross_out.w, L.cross_out.b, d, s.px, true);
        pc.add = s.px; pc.add_rows = M; pc.ld_add = d;
        gemm_tcgen05(pc, st);
        layernorm(s.px, L.mlp_ln_w, L.mlp_ln_b, 1e-5f, s.py, nullptr, M, d, st);
        GemmParams p1 = lin(s.py, M, d, L.mlp1.w, L.mlp1.b, 4 * d, s.phid, false);
        p1.gelu = 1;
        gemm_tcgen05(p1, st);
        GemmParams p2 = lin(s.phid, M, 4 * d, L.mlp2.w, L.mlp2.b, d, s.px, true);
        p2.add = s.px; p2.add_rows = M; p2.ld_add = d;
        gemm_tcgen05(p2, st);
    }
    layernorm(s.px, s.ln_w, s.ln_b, 1e-5f, nullptr, s.pout, M, d, st);  // decoder.py:316
    fill_table_kernel<<<1, 256, 0, st>>>(s.table, beam_idx, M, beam_idx);
    B200_LAUNCH_CHECK();
}

// =================================================================================================
// decoder1 step (whisper/decoder.py:241-257, 261-327; coreml.mm:404-444): descriptor of the batched persistent step kernel
// =================================================================================================
static void build_step_model() {
    State& s = S();
    if (!s.dec1_loaded || !s.dec256_loaded || !s.mkv) return;
    if (s.Ld > DB_MAX_LAYERS) return;
    DbModel m{};
    m.d = s.d; m.H = s.H; m.Ld = s.Ld; m.V = s.V; m.n_tiles_vocab = (s.V + 15) / 16;
    m.tok_emb = s.tok_emb; m.tok_emb_frag = s.tok_emb_frag; m.pos_emb = s.pos_emb; m.ln_w = s.ln_w; m.ln_b = s.ln_b;
    for (int l = 0; l < s.Ld; ++l) {
        const DecLayer& L = s.dec_layers[l];
        DbLayer& o = m.layers[l];
        o.qkv = L.qkv.frag; o.attn_out = L.attn_out.frag; o.cross_q = L.cross_q.frag; o.cross_out = L.cross_out.frag;
        o.mlp1 = L.mlp1.frag; o.mlp2 = L.mlp2.frag;
        o.qkv_b = L.qkv.b; o.attn_out_b = L.attn_out.b; o.cross_q_b = L.cross_q.b; o.cross_out_b = L.cross_out.b;
        o.mlp1_b = L.mlp1.b; o.mlp2_b = L.mlp2.b;
        o.ln1_w = L.attn_ln_w; o.ln1_b = L.attn_ln_b; o.ln2_w = L.cross_ln_w; o.ln2_b = L.cross_ln_b; o.ln3_w = L.mlp_ln_w; o.ln3_b = L.mlp_ln_b;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&s.n_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&s.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    batch_set_model(m);
}

// =================================================================================================
// load / close helpers
// =================================================================================================
static bool check_dims(const WeightFile& w, int idx, int expect, const char* what) {
    const int* dims = w.i32_host("dims");
    if (!dims) return false;
    if (dims[idx] != expect) { record_error("%s: %s = %d in file, %d requested", w.path().c_str(), what, dims[idx], expect); return false; }
    return true;
}

static void build_dump_slots() {
    State& s = S();
    if (!s.Ld || !s.H) return;
    std::vector<int> slots((size_t)s.Ld * s.H, -1);
    int n = 0;
    if (s.align_heads.empty()) {                                        // model.py:55-58: last half of the layers
        for (int l = s.Ld / 2; l < s.Ld; ++l)
            for (int h = 0; h < s.H; ++h)
                if (n < s.n_align) slots[l * s.H + h] = n++;
    } else {
        for (size_t i = 0; i + 1 < s.align_heads.size(); i += 2) {
            const int l = s.align_heads[i], h = s.align_heads[i + 1];
            if (l >= 0 && l < s.Ld && h >= 0 && h < s.H && n < s.n_align) slots[l * s.H + h] = n++;
        }
    }
    if (!s.d_dump_slot) dev_alloc(&s.d_dump_slot, slots.size());
    if (s.d_dump_slot) B200_CHECK(cudaMemcpy(s.d_dump_slot, slots.data(), slots.size() * sizeof(int), cudaMemcpyHostToDevice));
}

static bool load_decoder_weights(const char* path, int n_layer, int n_state) {
    State& s = S();
    if (s.dec_w.loaded()) { ++s.dec_w.refcount; return true; }
    if (!s.dec_w.load(path)) return false;
    const WeightFile& w = s.dec_w;
    if (!check_dims(w, 9, n_layer, "n_text_layer") || !check_dims(w, 7, n_state, "n_text_state")) { s.dec_w.unload(); return false; }
    s.dec_w.refcount = 1;
    s.V = w.i32_host("dims")[5];
    s.tok_emb = w.b16("tok_emb.w"); s.tok_emb_frag = w.b16("tok_emb.frag"); s.pos_emb = w.f32("pos_emb");
    s.ln_w = w.f32("ln.w"); s.ln_b = w.f32("ln.b");
    s.dec_layers.resize(n_layer);
    auto getlin = [&](const std::string& n) { return DecLinear{w.b16(n + ".w"), w.b16(n + ".frag"), w.f32(n + ".b")}; };
    for (int l = 0; l < n_layer; ++l) {
        const std::string p = "l" + std::to_string(l) + ".";
        DecLayer& L = s.dec_layers[l];
        L.attn_ln_w = w.f32(p + "attn_ln.w"); L.attn_ln_b = w.f32(p + "attn_ln.b");
        L.cross_ln_w = w.f32(p + "cross_attn_ln.w"); L.cross_ln_b = w.f32(p + "cross_attn_ln.b");
        L.mlp_ln_w = w.f32(p + "mlp_ln.w"); L.mlp_ln_b = w.f32(p + "mlp_ln.b");
        L.qkv = getlin(p + "qkv"); L.attn_out = getlin(p + "attn_out"); L.cross_q = getlin(p + "cross_q");
        L.cross_out = getlin(p + "cross_out"); L.mlp1 = getlin(p + "mlp1"); L.mlp2 = getlin(p + "mlp2");
    }
    return true;
}
static void release_decoder_weights() {
    State& s = S();
    if (s.dec_w.loaded() && --s.dec_w.refcount <= 0) { s.dec_w.unload(); s.dec_layers.clear(); }
}

}  // namespace b200

using namespace b200;

// =================================================================================================
// extern "C": Part 1
// =================================================================================================
extern "C" {

void loadEncoder(const char* modelFolderPath, int n_layer, int n_state, int n_mels) {
    State& s = S();
    if (s.enc_loaded) return;                                           // idempotent (coreml.mm:43-45)
    use_device();
    const std::string path = std::string(modelFolderPath) + "/Encoder.b2w";
    if (!s.enc_w.load(path)) return;
    const WeightFile& w = s.enc_w;
    if (!check_dims(w, 4, n_layer, "n_audio_layer") || !check_dims(w, 2, n_state, "n_audio_state") ||
        !check_dims(w, 0, n_mels, "n_mels")) { s.enc_w.unload(); return; }
    if (n_state % 128 != 0) { record_error("loadEncoder: n_state %d must be a multiple of 128", n_state); s.enc_w.unload(); return; }
    s.d = n_state; s.H = n_state / 64; s.Le = n_layer; s.n_mels = n_mels; s.cpad = (n_mels + 63) / 64 * 64;
    s.conv1_w = w.b16("conv1.w"); s.conv1_b = w.f32("conv1.b"); s.conv2_w = w.b16("conv2.w"); s.conv2_b = w.f32("conv2.b");
    s.pos = w.f32("pos"); s.ln_post_w = w.f32("ln_post.w"); s.ln_post_b = w.f32("ln_post.b");
    s.enc_layers.resize(n_layer);
    for (int l = 0; l < n_layer; ++l) {
        const std::string p = "l" + std::to_string(l) + ".";
        EncLayer& L = s.enc_layers[l];
        L.attn_ln_w = w.f32(p + "attn_ln.w"); L.attn_ln_b = w.f32(p + "attn_ln.b");
        L.qkv_w = w.b16(p + "qkv.w"); L.qkv_b = w.f32(p + "qkv.b"); L.out_w = w.b16(p + "out.w"); L.out_b = w.f32(p + "out.b");
        L.mlp_ln_w = w.f32(p + "mlp_ln.w"); L.mlp_ln_b = w.f32(p + "mlp_ln.b");
        L.mlp1_w = w.b16(p + "mlp1.w"); L.mlp1_b = w.f32(p + "mlp1.b"); L.mlp2_w = w.b16(p + "mlp2.w"); L.mlp2_b = w.f32(p + "mlp2.b");
    }
    dev_alloc(&s.mel_stage, (size_t)n_mels * N_FRAMES);
    s.w_cap = 0;
    if (!ensure_encoder_capacity(1)) return;
    s.enc_loaded = true;
}

void closeEncoder() {
    State& s = S();
    if (!s.enc_loaded) return;
    use_device();
    B200_CHECK(cudaDeviceSynchronize());
    dev_free(&s.melrows); dev_free(&s.h1); dev_free(&s.x); dev_free(&s.y); dev_free(&s.qkv); dev_free(&s.att);
    dev_free(&s.hid); dev_free(&s.xa); dev_free(&s.mel_stage); dev_free(&s.d_seeks);
    s.enc_w.unload(); s.enc_layers.clear(); s.w_cap = 0; s.n_windows = 0; s.enc_loaded = false;
    gemm_clear_map_cache(); attention_clear_map_cache(); decode_clear_graphs(); encoder_clear_graphs();
}

void encoderPredict(float* melSegment) {
    State& s = S();
    if (!s.enc_loaded) { record_error("encoderPredict: encoder not loaded"); return; }
    use_device();
    B200_CHECK(cudaMemcpyAsync(s.mel_stage, melSegment, (size_t)s.n_mels * N_FRAMES * sizeof(float), cudaMemcpyHostToDevice, s.stream));
    B200_CHECK(cudaMemsetAsync(s.d_seeks, 0, sizeof(int), s.stream));
    run_encoder(s.mel_stage, N_FRAMES, N_FRAMES, 1);
    s.cur_window = 0;
    B200_CHECK(cudaStreamSynchronize(s.stream));
}

void loadCrossKV(const char* modelPath, int n_layer, int n_state) {
    State& s = S();
    if (s.ckv_loaded) return;
    use_device();
    if (!s.ckv_w.load(modelPath)) return;
    if (!check_dims(s.ckv_w, 9, n_layer, "n_text_layer") || !check_dims(s.ckv_w, 7, n_state, "n_text_state")) { s.ckv_w.unload(); return; }
    if (s.d && s.d != n_state) { record_error("loadCrossKV: n_state %d != encoder %d", n_state, s.d); s.ckv_w.unload(); return; }
    s.d = n_state; s.H = n_state / 64; s.Ld = n_layer;                  // n_head recomputed (coreml.mm:139)
    s.ckv_wt = s.ckv_w.b16("ckv.w"); s.ckv_b = s.ckv_w.f32("ckv.b");
    s.ckv_cap = 0;
    if (!dev_alloc(&s.ckv, s.ckv_window_elems()) || !dev_alloc(&s.ckv_frag, s.ckv_frag_window_elems(), true)) return;
    s.ckv_cap = 1;
    s.ckv_loaded = true;
}

void closeCrossKV() {
    State& s = S();
    if (!s.ckv_loaded) return;
    use_device();
    B200_CHECK(cudaDeviceSynchronize());
    dev_free(&s.ckv); dev_free(&s.ckv_frag); s.ckv_cap = 0; s.ckv_w.unload(); s.ckv_loaded = false;
    gemm_clear_map_cache(); attention_clear_map_cache(); decode_clear_graphs(); encoder_clear_graphs();
}

void crossKVPredict() {
    State& s = S();
    if (!s.ckv_loaded || !s.enc_loaded) { record_error("crossKVPredict: encoder / crossKV not loaded"); return; }
    use_device();
    run_cross_kv(1);
    B200_CHECK(cudaStreamSynchronize(s.stream));
}

void loadDecoder256(const char* modelPath, int n_layer, int n_state, int n_head, int n_alignment_head, int beam_size) {
    State& s = S();
    if (s.dec256_loaded) return;
    use_device();
    (void)n_head;                                                       // recomputed as n_state / 64 (coreml.mm:139)
    if (beam_size < 1 || beam_size > STEP_MAX_BEAMS) { record_error("loadDecoder256: beam_size %d outside [1, %d]", beam_size, STEP_MAX_BEAMS); return; }
    if (!load_decoder_weights(modelPath, n_layer, n_state)) return;
    s.d = n_state; s.H = n_state / 64; s.Ld = n_layer; s.bs = beam_size; s.n_align = n_alignment_head > 0 ? n_alignment_head : 0;
    const size_t d = s.d, M = PREFILL_CTX;
    bool ok = true;
    ok &= dev_alloc(&s.mkv, (size_t)2 * s.Ld * s.bs * N_TEXT_CTX * d, true);    // coreml.mm:231-233
    ok &= dev_alloc(&s.table, (size_t)STEP_MAX_BEAMS * N_TEXT_CTX, true);
    ok &= dev_alloc(&s.px, M * d); ok &= dev_alloc(&s.pout, M * d); ok &= dev_alloc(&s.pmask, M * M);
    ok &= dev_alloc(&s.pchw, (size_t)(s.n_align > 0 ? s.n_align : 1) * M * N_AUDIO_CTX);
    ok &= dev_alloc(&s.py, M * d); ok &= dev_alloc(&s.pqkv, M * 3 * d); ok &= dev_alloc(&s.patt, M * d);
    ok &= dev_alloc(&s.phid, M * 4 * d); ok &= dev_alloc(&s.pq, M * d);
    dev_free(&s.d_dump_slot);
    build_dump_slots();
    if (!ok) return;
    s.dec256_loaded = true;
    build_step_model();
}

void closeDecoder256() {
    State& s = S();
    if (!s.dec256_loaded) return;
    use_device();
    B200_CHECK(cudaDeviceSynchronize());
    batch_free();
    dev_free(&s.mkv); dev_free(&s.table); dev_free(&s.px); dev_free(&s.pout); dev_free(&s.pmask); dev_free(&s.pchw);
    dev_free(&s.py); dev_free(&s.pqkv); dev_free(&s.patt); dev_free(&s.phid); dev_free(&s.pq); dev_free(&s.d_dump_slot);
    release_decoder_weights();
    s.align_heads.clear();             // a model loaded next starts from the default alignment heads (model.py:55-58), not this one's
    s.dec256_loaded = false;
    gemm_clear_map_cache(); attention_clear_map_cache(); decode_clear_graphs(); encoder_clear_graphs();
}

void decoder256Predict(float* x, float* qk_mask, float* out_x, float* out_cross_head_weights, int beam_idx) {
    State& s = S();
    if (!s.dec256_loaded || !s.ckv_loaded) { record_error("decoder256Predict: decoder256 / crossKV not loaded"); return; }
    if (beam_idx < 0 || beam_idx >= s.bs) { record_error("decoder256Predict: beam_idx %d outside [0, %d)", beam_idx, s.bs); return; }
    use_device();
    const size_t d = s.d, M = PREFILL_CTX;
    B200_CHECK(cudaMemcpyAsync(s.px, x, M * d * sizeof(float), cudaMemcpyHostToDevice, s.stream));
    B200_CHECK(cudaMemcpyAsync(s.pmask, qk_mask, M * M * sizeof(float), cudaMemcpyHostToDevice, s.stream));
    run_prefill(beam_idx, out_cross_head_weights != nullptr, PREFILL_CTX);
    B200_CHECK(cudaMemcpyAsync(out_x, s.pout, M * d * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
    if (out_cross_head_weights && s.n_align > 0)
        B200_CHECK(cudaMemcpyAsync(out_cross_head_weights, s.pchw, (size_t)s.n_align * M * N_AUDIO_CTX * sizeof(float),
                                   cudaMemcpyDeviceToHost, s.stream));
    B200_CHECK(cudaStreamSynchronize(s.stream));
}

void loadDecoder1(const char* modelPath, int n_layer, int n_state, int n_head, int n_vocab) {
    State& s = S();
    if (s.dec1_loaded) return;
    use_device();
    (void)n_head;                                                       // coreml.mm:389
    if (!load_decoder_weights(modelPath, n_layer, n_state)) return;
    if (s.V != n_vocab) { record_error("loadDecoder1: n_vocab %d in file, %d requested", s.V, n_vocab); release_decoder_weights(); return; }
    s.d = n_state; s.H = n_state / 64; s.Ld = n_layer;
    const size_t d = s.d, B = STEP_MAX_BEAMS;
    bool ok = true;
    ok &= dev_alloc(&s.slogits, B * (size_t)s.V); ok &= dev_alloc(&s.smask, (size_t)512); ok &= dev_alloc(&s.sxin, B * d);
    if (!ok) return;
    s.dec1_loaded = true;
    build_step_model();
}

void closeDecoder1() {
    State& s = S();
    if (!s.dec1_loaded) return;
    use_device();
    B200_CHECK(cudaDeviceSynchronize());
    batch_free();
    dev_free(&s.slogits); dev_free(&s.smask); dev_free(&s.sxin);
    decode_clear_graphs();
    release_decoder_weights();
    s.dec1_loaded = false;
}

// cache[:, b, :text_offset] = cache[:, indices[b], :text_offset] (coreml.mm:251-277, decoding.py:189-204).
// The cache rows never move: the permutation is applied to the (beam, position) -> slot table that
// every reader goes through, so the call costs 2 KB of traffic instead of 4 * 2Ld * bs * t * d bytes.
void rearrange_mkv(int* indices, int text_offset) {
    State& s = S();
    if (!s.dec256_loaded) { record_error("rearrange_mkv: decoder not loaded"); return; }
    if (text_offset < 0 || text_offset > N_TEXT_CTX) { record_error("rearrange_mkv: text_offset %d", text_offset); return; }
    use_device();
    int h_idx[STEP_MAX_BEAMS];
    for (int b = 0; b < s.bs; ++b) {
        h_idx[b] = indices[b];
        if (h_idx[b] < 0 || h_idx[b] >= s.bs) { record_error("rearrange_mkv: index %d outside [0, %d)", h_idx[b], s.bs); return; }
    }
    PermuteArgs pa;
    for (int b = 0; b < STEP_MAX_BEAMS; ++b) pa.src[b] = b < s.bs ? h_idx[b] : b;
    permute_table_kernel<<<1, 256, 0, s.stream>>>(s.table, pa, s.bs, text_offset);
    B200_LAUNCH_CHECK();
    B200_CHECK(cudaStreamSynchronize(s.stream));
}

void decoder1Predict(float* x, float* qk_mask, int text_offset, float* out_x) {
    State& s = S();
    if (!s.dec1_loaded || !s.dec256_loaded || !s.ckv_loaded) { record_error("decoder1Predict: decoder1 / decoder256 / crossKV not loaded"); return; }
    if (text_offset < 1 || text_offset >= N_TEXT_CTX) { record_error("decoder1Predict: text_offset %d outside [1, 448)", text_offset); return; }
    use_device();
    const size_t d = s.d;
    const int nb = s.bs;
    if (!batch_available()) { record_error("decoder1Predict: the step kernel does not support these dimensions (n_state %d, %d layers)", s.d, s.Ld); return; }
    B200_CHECK(cudaMemcpyAsync(s.sxin, x, nb * d * sizeof(float), cudaMemcpyHostToDevice, s.stream));
    B200_CHECK(cudaMemcpyAsync(s.smask, qk_mask, (size_t)(nb == 1 ? 450 : 449) * sizeof(float), cudaMemcpyHostToDevice, s.stream));
    run_step_batch_abi(nb, text_offset, s.smask, s.sxin);
    B200_CHECK(cudaMemcpyAsync(out_x, s.slogits, (size_t)nb * s.V * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
    B200_CHECK(cudaStreamSynchronize(s.stream));
}

// =================================================================================================
// Part 2 (core): errors, device, alignment heads, counters
// =================================================================================================
int b200LastError(char* buf, int buf_len) { return b200::take_errors(buf, buf_len); }
long b200KernelLaunchCount() { return b200::g_launch_count; }
void b200SetDevice(int device) { S().device = device; }
void b200SetAlignmentHeads(const int* layer_head_pairs, int n) {
    State& s = S();
    s.align_heads.assign(layer_head_pairs, layer_head_pairs + 2 * (size_t)n);
    if (s.dec256_loaded) { use_device(); build_dump_slots(); }
}
void b200SelectWindow(int w) {
    State& s = S();
    if (w < 0 || w >= (s.ckv_cap > 0 ? s.ckv_cap : 1)) { record_error("b200SelectWindow: %d outside [0, %d)", w, s.ckv_cap); return; }
    s.cur_window = w;
}

}  // extern "C"
