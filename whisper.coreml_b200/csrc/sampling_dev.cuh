// Device bodies of the sampling kernels, shared by the stand-alone kernels (sampling.cu) and the persistent
// decoder step kernel (decoder_batch.cu).
#pragma once
#include "sampling.cuh"

namespace b200 {

struct Rules {
    int at_begin, last_ts, penult_ts, last_stamp, lim;
};

__device__ __forceinline__ bool is_masked(int v, const DecodeSpec& sp, const DecodeState& st, const Rules& r) {
    if (sp.d_suppress[v]) return true;                                              // SuppressTokens (:460-465)
    if (r.at_begin && st.suppress_blank &&
        (v == sp.eot || v == sp.blank[0] || v == sp.blank[1] || v == sp.blank[2] || v == sp.blank[3]))
        return true;                                                                // SuppressBlank (:450-457)
    if (st.without_timestamps) return false;
    const int tb = sp.timestamp_begin;                                              // ApplyTimestampRules (:468-523)
    if (v == sp.no_timestamps) return true;
    if (r.last_ts) {
        if (r.penult_ts) { if (v >= tb) return true; }
        else if (v < sp.eot) return true;
    }
    if (r.last_stamp >= 0 && v >= tb && v < r.lim) return true;
    if (r.at_begin) {
        if (v < tb) return true;
        if (st.max_initial_ts >= 0 && v > tb + st.max_initial_ts) return true;
    }
    return false;
}

__device__ __forceinline__ void online_add(float& m, float& s, float x) {
    if (x == -INFINITY) return;
    if (x > m) { s = s * __expf(m - x) + 1.f; m = x; }
    else s += __expf(x - m);
}
__device__ __forceinline__ void online_merge(float& m, float& s, float m2, float s2) {
    if (m2 == -INFINITY) return;
    if (m == -INFINITY) { m = m2; s = s2; return; }
    const float M = fmaxf(m, m2);
    s = s * __expf(m - M) + s2 * __expf(m2 - M);
    m = M;
}

// cycle probes for experiments with the in-kernel sampling tail (compile with -DB200_PROBES): thread 0 of block 0 stores
// clock64() at g_probe[idx]; tools/step_timeline.py prints them
#ifdef B200_PROBES
__device__ unsigned long long* g_probe = nullptr;
__device__ __forceinline__ void probe(int idx, int tid) { if (g_probe && tid == 0 && blockIdx.x == 0) g_probe[idx] = (unsigned long long)clock64(); }
#else
__device__ __forceinline__ void probe(int, int) {}
#endif

// ---- Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3"): a counter-based generator, so every
// (decode stream, row, step, token) draws its own number with no state to carry between kernels ----------------------------
__device__ __forceinline__ uint32_t philox_first(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0;
}
// Gumbel(0, 1) noise for token v: argmax_v (logit_v / T + G_v) is a draw from Categorical(logits / T)
__device__ __forceinline__ float gumbel_noise(const DecodeState& st, int row, int v) {
    const uint32_t r = philox_first((uint32_t)v, (uint32_t)st.step, (uint32_t)row, st.stream, st.seed_lo, st.seed_hi);
    const float u = ((float)(r >> 8) + 0.5f) * (1.0f / 16777216.0f);      // (0, 1), 24 bits
    return -__logf(-__logf(u));
}

constexpr int SP_CHUNK_MAX = SAMPLE_CHUNK_TOKENS_MAX;                     // >= largest chunk (1799 text / 1502 timestamp tokens)

// One (chunk, beam) unit; `tid` in [0, SP_THREADS); sync() is a barrier over exactly those SP_THREADS threads (256 in the
// stand-alone kernel, the 128 consumer threads inside the persistent step kernel).
template <int SP_THREADS, class Sync>
__device__ __forceinline__ void sample_partial_body(const SampleArgs& a, int chunk, int b, int tid, Sync sync) {
    constexpr int SP_PER_THREAD = SP_CHUNK_MAX / SP_THREADS, SP_WARPS = SP_THREADS / 32;
    __shared__ Rules rules;
    __shared__ int s_last;
    __shared__ float red_m[8], red_s[8];
    const DecodeState& st = *a.st;
    if (st.done) return;
    const int lane = tid & 31, warp = tid >> 5;
    const DecodeSpec& sp = a.spec;
    const int V = sp.n_vocab, tb = sp.timestamp_begin;
    const int* seq = a.tokens + b * DEC_TOK_LD + st.sample_begin;
    const int n = st.L - st.sample_begin;
    probe(300, tid);
    if (tid == 0) s_last = -1;
    sync();
    int last = -1;
    for (int j = tid; j < n; j += SP_THREADS) if (seq[j] >= tb) last = j;      // position of the last timestamp token
    last = max(last, __shfl_xor_sync(0xffffffffu, last, 16)); last = max(last, __shfl_xor_sync(0xffffffffu, last, 8));
    last = max(last, __shfl_xor_sync(0xffffffffu, last, 4)); last = max(last, __shfl_xor_sync(0xffffffffu, last, 2));
    last = max(last, __shfl_xor_sync(0xffffffffu, last, 1));
    if (lane == 0 && last >= 0) atomicMax(&s_last, last);
    sync();
    if (tid == 0) {
        Rules r;
        r.at_begin = n == 0;
        r.last_ts = n >= 1 && seq[n - 1] >= tb;
        r.penult_ts = n < 2 || seq[n - 2] >= tb;
        r.last_stamp = s_last >= 0 ? seq[s_last] : -1;
        r.lim = (r.last_ts && !r.penult_ts) ? r.last_stamp : r.last_stamp + 1;
        rules = r;
    }
    sync();
    probe(301, tid);
    const Rules r = rules;
    int lo, hi;
    if (chunk < SAMPLE_TEXT_CHUNKS) {
        const int per = (tb + SAMPLE_TEXT_CHUNKS - 1) / SAMPLE_TEXT_CHUNKS;
        lo = chunk * per; hi = min(tb, lo + per);
    } else { lo = tb; hi = V; }
    const float* x = a.logits + (long)b * a.ld_logits;
    // Three passes so that nothing in the per-token work is a dependent chain: (1) all loads, (2) the filters with the
    // thread-uniform state hoisted into registers, (3) max, then the sum of exponentials.
    float val[SP_PER_THREAD];
    unsigned supm = 0;
#pragma unroll
    for (int e = 0; e < SP_PER_THREAD; ++e) {
        const int v = lo + e * SP_THREADS + tid;
        val[e] = -INFINITY;
        if (v < hi) { val[e] = x[v]; supm |= (unsigned)(sp.d_suppress[v] != 0) << e; }
    }
    {
        const int eot = sp.eot, no_ts = sp.no_timestamps, b0 = sp.blank[0], b1 = sp.blank[1], b2 = sp.blank[2], b3 = sp.blank[3];
        const bool blank_rule = r.at_begin && st.suppress_blank, ts_rules = !st.without_timestamps;
        const int max_init = st.max_initial_ts;
#pragma unroll
        for (int e = 0; e < SP_PER_THREAD; ++e) {
            const int v = lo + e * SP_THREADS + tid;
            bool masked = v >= hi || ((supm >> e) & 1u);                                             // SuppressTokens (:460-465)
            masked |= blank_rule && (v == eot || v == b0 || v == b1 || v == b2 || v == b3);           // SuppressBlank (:450-457)
            if (ts_rules) {                                                                           // ApplyTimestampRules (:468-523)
                masked |= v == no_ts;
                if (r.last_ts) masked |= r.penult_ts ? v >= tb : v < eot;
                masked |= r.last_stamp >= 0 && v >= tb && v < r.lim;
                if (r.at_begin) masked |= v < tb || (max_init >= 0 && v > tb + max_init);
            }
            if (masked) val[e] = -INFINITY;
        }
    }
    float m = -INFINITY, s = 0.f;
#pragma unroll
    for (int e = 0; e < SP_PER_THREAD; ++e) m = fmaxf(m, val[e]);
    if (m != -INFINITY) {
#pragma unroll
        for (int e = 0; e < SP_PER_THREAD; ++e) s += __expf(val[e] - m);                             // exp(-inf) = 0 for the masked tokens
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) online_merge(m, s, __shfl_xor_sync(0xffffffffu, m, o), __shfl_xor_sync(0xffffffffu, s, o));
    probe(302, tid);
    if (lane == 0) { red_m[warp] = m; red_s[warp] = s; }
    sync();
    if (tid == 0) {
        for (int w = 1; w < SP_WARPS; ++w) online_merge(m, s, red_m[w], red_s[w]);
        a.part->m[b][chunk] = m; a.part->s[b][chunk] = s;
    }
    // temperature > 0 (greedy rows only): the selection key becomes logit / T + Gumbel noise; the log-softmax partial above stays
    // that of the unperturbed logits (decoding.py:307-313), and the winner's plain logit travels along (topx)
    const float temp = st.temperature;
    float plain[SP_PER_THREAD];
    if (temp > 0.f) {
        const float inv_t = 1.f / temp;
#pragma unroll
        for (int e = 0; e < SP_PER_THREAD; ++e) {
            plain[e] = val[e];
            const int v = lo + e * SP_THREADS + tid;
            if (v < hi && val[e] != -INFINITY) val[e] = fmaf(val[e], inv_t, gumbel_noise(st, b, v));
        }
    }
    // chunk-local top-k (ties -> lowest index) in two levels: every warp extracts its own top-k from the register-resident
    // values with shuffles only, then warp 0 merges the <= 8 * k survivors - one block barrier instead of two per rank
    __shared__ float wv[8][SAMPLE_MAX_K]; __shared__ int wi[8][SAMPLE_MAX_K]; __shared__ float wx[8];
    probe(303, tid);
    unsigned taken = 0;
    for (int c = 0; c < a.k; ++c) {
        float bv = -INFINITY; int bi = 0x7fffffff;
#pragma unroll
        for (int e = 0; e < SP_PER_THREAD; ++e) {
            const int v = lo + e * SP_THREADS + tid;
            if (v < hi && !((taken >> e) & 1u) && (val[e] > bv || (val[e] == bv && v < bi))) { bv = val[e]; bi = v; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { wv[warp][c] = bv; wi[warp][c] = bi; }
        if (bi != 0x7fffffff && (bi - lo) % SP_THREADS == tid) {
            taken |= 1u << ((bi - lo) / SP_THREADS);         // a picked slot never competes again
            if (temp > 0.f && c == 0) {                       // the owner of the warp's winner publishes its plain logit
                float px = -INFINITY;
#pragma unroll
                for (int e = 0; e < SP_PER_THREAD; ++e) if (e == (bi - lo) / SP_THREADS) px = plain[e];
                wx[warp] = px;
            }
        }
    }
    sync();
    probe(304, tid);
    if (warp == 0) {
        const int n = SP_WARPS * a.k;                      // <= 72 survivors, <= 3 per lane
        float cv[3]; int ci[3];
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            const int q = lane + 32 * e;
            cv[e] = -INFINITY; ci[e] = 0x7fffffff;
            if (q < n) { cv[e] = wv[q / a.k][q % a.k]; ci[e] = wi[q / a.k][q % a.k]; }
        }
        unsigned tk = 0;
        for (int c = 0; c < a.k; ++c) {
            float bv = -INFINITY; int bi = 0x7fffffff;
#pragma unroll
            for (int e = 0; e < 3; ++e)
                if (!((tk >> e) & 1u) && (cv[e] > bv || (cv[e] == bv && ci[e] < bi))) { bv = cv[e]; bi = ci[e]; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) { a.part->topv[b][chunk][c] = bv; a.part->topi[b][chunk][c] = bi; }
            if (bi != 0x7fffffff) {
#pragma unroll
                for (int e = 0; e < 3; ++e)
                    if (ci[e] == bi) {
                        tk |= 1u << e;                            // token indices are unique
                        if (temp > 0.f && c == 0) a.part->topx[b][chunk] = wx[(lane + 32 * e) / a.k];      // the winner's warp
                    }
            }
        }
    }
    probe(305, tid);
}


// `stage`: DEC_MAX_BEAMS * DEC_TOK_LD ints of shared scratch; `tid` in [0, NT); sync() as above.  Nothing here may live in
// local memory (inside the persistent kernel a stack access is an L2 round trip): the sort keys are in shared memory.
// PRE: the caller has already copied the token rows into `stage` and the slot table into `stage2` (the persistent step kernel
// stages them with cp.async; without an L1 a plain global load per element would be an L2 round trip each).
template <int NT, bool PRE, class Sync>
__device__ __forceinline__ void beam_update_body(const BeamUpdateArgs& a, int* stage, int* stage2, int tid, Sync sync) {
    const int* tok_src = PRE ? stage : a.tokens;
    __shared__ float sc[DEC_MAX_BEAMS * SAMPLE_MAX_K]; __shared__ int id[DEC_MAX_BEAMS * SAMPLE_MAX_K];
    __shared__ int nsrc[DEC_MAX_BEAMS], ntok[DEC_MAX_BEAMS];
    __shared__ float nsum[DEC_MAX_BEAMS];
    __shared__ int fin_src[DEC_MAX_BEAMS]; __shared__ float fin_sc[DEC_MAX_BEAMS];
    __shared__ int s_nfin_new, s_done;
    __shared__ float c_lp[DEC_MAX_BEAMS * SAMPLE_MAX_K]; __shared__ int c_tok[DEC_MAX_BEAMS * SAMPLE_MAX_K];
    DecodeState& st = *a.st;
    if (st.done) return;
    const int warp = tid >> 5, lane = tid & 31, nb = a.nb, L = st.L;
    probe(310, tid);
    // ---- warp b: log-softmax normaliser and top-k of beam b from the chunk partials ----------------------
    for (int bw = warp; bw < nb; bw += NT / 32) {
        const int warp = bw;                                 // one warp per beam
        const SamplePartials& P = *a.part;
        float m = -INFINITY, s = 0.f;
        if (lane < SAMPLE_CHUNKS) { m = __ldcg(&P.m[warp][lane]); s = __ldcg(&P.s[warp][lane]); }   // other CTAs' partials: read at L2 (never a stale L1 line)
        float mt = lane < SAMPLE_TEXT_CHUNKS ? m : -INFINITY, stx = lane < SAMPLE_TEXT_CHUNKS ? s : 0.f;   // text group
        float mq = lane == SAMPLE_TEXT_CHUNKS ? m : -INFINITY, sq = lane == SAMPLE_TEXT_CHUNKS ? s : 0.f;  // timestamp group
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            online_merge(mt, stx, __shfl_xor_sync(0xffffffffu, mt, o), __shfl_xor_sync(0xffffffffu, stx, o));
            online_merge(mq, sq, __shfl_xor_sync(0xffffffffu, mq, o), __shfl_xor_sync(0xffffffffu, sq, o));
        }
        const float lse_q = mq == -INFINITY ? -INFINITY : mq + logf(sq);
        float ma = mt, sa = stx;
        online_merge(ma, sa, mq, sq);
        const float lse_all = ma == -INFINITY ? -INFINITY : ma + logf(sa);
        // "if sum of probability over timestamps is above any other token, sample timestamp" (decoding.py:525-532)
        const bool mask_text = !st.without_timestamps && (lse_q - lse_all) > (mt - lse_all);
        const float lse = mask_text ? lse_q : lse_all;
        // candidates: lane c holds the k survivors of chunk c, read ONCE into registers (no index arithmetic)
        float cv[9]; int ci[9];
        const bool mine = lane < SAMPLE_CHUNKS && !(mask_text && lane < SAMPLE_TEXT_CHUNKS);
#pragma unroll
        for (int e = 0; e < 9; ++e) {
            cv[e] = -INFINITY; ci[e] = 0x7fffffff;
            if (mine && e < a.k) { cv[e] = __ldcg(&P.topv[warp][lane][e]); ci[e] = __ldcg(&P.topi[warp][lane][e]); }
        }
        const float px = (st.temperature > 0.f && mine) ? __ldcg(&P.topx[warp][lane]) : -INFINITY;   // plain logit of the chunk's winner
        unsigned taken = 0;
        for (int c = 0; c < a.k; ++c) {
            float bv = -INFINITY; int bi = 0x7fffffff;
#pragma unroll
            for (int e = 0; e < 9; ++e)
                if (!((taken >> e) & 1u) && (cv[e] > bv || (cv[e] == bv && ci[e] < bi))) { bv = cv[e]; bi = ci[e]; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (st.temperature > 0.f) {                      // sampled token: its log-probability comes from the unperturbed logit
                const unsigned who = __ballot_sync(0xffffffffu, bi != 0x7fffffff && ci[0] == bi);
                bv = who ? __shfl_sync(0xffffffffu, px, __ffs(who) - 1) : -INFINITY;
            }
            if (bi != 0x7fffffff) {
#pragma unroll
                for (int e = 0; e < 9; ++e) if (ci[e] == bi) taken |= 1u << e;       // token indices are unique across chunks
            }
            if (lane == 0) {
                c_lp[warp * a.k + c] = bv - lse; c_tok[warp * a.k + c] = bi;
                a.cand_lp[warp * a.k + c] = bv - lse; a.cand_tok[warp * a.k + c] = bi;
            }
        }
    }
    sync();
    probe(311, tid);
    if (!a.update) return;
    if (st.beam_mode) {
        // stable descending sort of the <= 72 cumulative scores by rank counting: thread i places candidate i (the Python
        // `sorted(..., reverse=True)` of decoding.py:377 keeps equal keys in their original order)
        __shared__ float key[DEC_MAX_BEAMS * SAMPLE_MAX_K];
        const int nsb = st.step == 0 ? 1 : nb, n = nsb * a.k;
        float mykey = 0.f;
        if (tid < n) {
            mykey = st.sum_lp[tid / a.k] + c_lp[tid];
            if (!(mykey == mykey)) mykey = -INFINITY;        // NaN (-inf - -inf) sorts last
            key[tid] = mykey;
        }
        sync();
        if (tid < n) {
            int rank = 0;
            for (int q = 0; q < n; ++q) { const float o = key[q]; rank += (o > mykey || (o == mykey && q < tid)) ? 1 : 0; }
            sc[rank] = mykey; id[rank] = tid;
        }
        sync();
    }
    if (tid == 0) {
        int nfin_new = 0, done = 0;
        if (!st.beam_mode) {                             // GreedyDecoder.update (:303-318): every row (best_of samples at T > 0) on its own
            done = 1;
            for (int b = 0; b < nb; ++b) {
                const int tok = c_tok[b * a.k], last = tok_src[b * DEC_TOK_LD + L - 1];
                nsrc[b] = b;
                ntok[b] = last == a.eot ? a.eot : tok;
                nsum[b] = st.sum_lp[b] + (last == a.eot ? 0.f : c_lp[b * a.k]);
                done = done && ntok[b] == a.eot;
            }
        } else {
            const int nsb = st.step == 0 ? 1 : nb;      // identical prefixes at step 0 collapse to one key set (:366-373)
            const int n = nsb * a.k;                     // sc / id were ranked by all threads above (stable, descending, :377)
            int cnt = 0;
            for (int i = 0; i < n && cnt < nb; ++i) {
                const int j = id[i] / a.k, tok = c_tok[id[i]];
                if (tok == a.eot) { if (nfin_new < DEC_MAX_BEAMS) { fin_src[nfin_new] = j; fin_sc[nfin_new] = sc[i]; ++nfin_new; } }
                else { nsrc[cnt] = j; ntok[cnt] = tok; nsum[cnt] = sc[i]; ++cnt; }
            }
            for (; cnt < nb; ++cnt) { nsrc[cnt] = nsrc[cnt > 0 ? cnt - 1 : 0]; ntok[cnt] = ntok[cnt > 0 ? cnt - 1 : 0]; nsum[cnt] = -INFINITY; }
        }
        s_nfin_new = nfin_new; s_done = done;
    }
    sync();
    probe(312, tid);
    // finished pool: at most max_candidates (= nb, patience 1) sequences, best first (:396-402)
    const int max_cand = nb;
    int nfin = st.n_finished;
    for (int f = 0; f < s_nfin_new && nfin < max_cand; ++f, ++nfin) {
        const int* src = tok_src + fin_src[f] * DEC_TOK_LD;
        int* dst = a.fin_tokens + nfin * DEC_TOK_LD;
        for (int i = tid; i < L; i += NT) dst[i] = src[i];
        if (tid == 0) { dst[L] = a.eot; st.fin_len[nfin] = L + 1; st.fin_score[nfin] = fin_sc[f]; }
    }
    sync();
    probe(313, tid);
    // permute token histories and KV slot tables by source beam, append the new tokens
    if (!PRE) {
        for (int i = tid; i < nb * DEC_TOK_LD; i += NT) stage[i] = a.tokens[i];
        sync();
    }
    for (int i = tid; i < nb * L; i += NT) {
        const int bb = i / L, p = i % L;
        a.tokens[bb * DEC_TOK_LD + p] = stage[nsrc[bb] * DEC_TOK_LD + p];
    }
    const int* tab_src = PRE ? stage2 : stage;
    if (!PRE) {
        sync();
        for (int i = tid; i < nb * 448; i += NT) stage[i] = a.table[i];
        sync();
    }
    for (int i = tid; i < nb * L; i += NT) {
        const int bb = i / L, p = i % L;
        if (p < 448) a.table[bb * 448 + p] = tab_src[nsrc[bb] * 448 + p];
    }
    if (tid < nb) { a.tokens[tid * DEC_TOK_LD + L] = ntok[tid]; st.sum_lp[tid] = nsum[tid]; }
    probe(314, tid);
    sync();
    if (tid == 0) {
        st.n_finished = nfin;
        st.L = L + 1; st.pos = L; st.step += 1;
        int done = s_done;
        if (st.beam_mode && nfin >= max_cand) done = 1;
        if (L + 1 > a.n_text_ctx) done = 1;                              // :732
        if (st.step >= st.sample_len) done = 1;
        st.done = done;
    }
}


}  // namespace b200
