// Generate loop: KV-cached (bf16 tier: latent-cached) greedy / sampled decode steps, one CUDA graph per branch replayed per
// step, on-device EOS bookkeeping [model/decoder.py:77-122].
#include "engine_internal.h"

// ------------------------------------------------------------------------------------------------ decode step
struct DecState {
    int64_t* cur_tok; int* step; int* done_step; int* block_counter; unsigned* call_ctr; int* seen;     // step/done/counter: [MAX_BRANCH]
};
static size_t dec_state_bytes(int B) { return (size_t)B * 8 + (3 * MAX_BRANCH + 4) * 4 + (size_t)B * 4; }
static DecState dec_state(texocr_handle* h, int B) {
    DecState d;
    char* p = (char*)h->dec_state.p;
    d.cur_tok = (int64_t*)p;
    int* ip = (int*)(p + (size_t)B * 8);
    d.step = ip; d.done_step = ip + MAX_BRANCH; d.block_counter = ip + 2 * MAX_BRANCH; d.call_ctr = (unsigned*)(ip + 3 * MAX_BRANCH);
    d.seen = ip + 3 * MAX_BRANCH + 4;
    return d;
}

// One greedy step for rows [row0, row0+rows) of a batch of B (a "branch": every row is independent, so the batch is cut
// into sub-batches whose step graphs run concurrently and hide each other's launch / dependency latency).
// `t_host` is only used for the profiler's byte accounting (-1: unknown).
static int enqueue_decode_step(texocr_handle* h, int B, int row0, int rows, int branch, int tcap, int eos, const int* d_enc_off,
                               int max_s, double sum_s, int t_host, cudaStream_t st) {
    const texocr_config& c = h->cfg;
    const int L = c.dec_layers;
    DecState ds = dec_state(h, B);
    RowCtx rc{rows, KC_DEC_GEMM, KC_DEC_ROW, h->dec_ln_g, h->dec_ln_b, row0};
    int* step = ds.step + branch;
    const size_t e = h->esz;
    int r;
    float* sbuf = rowf(h->s, rc, 256);
    float* xbuf = rowf(h->x, rc, 256);
    void* xnbuf = rowa(h, h->xn, rc, 256);
    // the step's input (x = embedding, xn = LN(x)) was written by enqueue_first_embed (step 0) or by the previous step's token kernel
    const double tkeys = t_host >= 0 ? (double)(t_host + 1) : 0.0;
    char* qb = (char*)rowa(h, h->qkv, rc, 1536);
    for (int l = 0; l < L; ++l) {
        // ---- causal self-attention, absorbed form: the cache holds this layer's LayerNorm'd inputs (256 per position), Q' = xn . Wqk^T,
        // C_h = softmax(Q'_h . Z^T / 8) . Z over positions 0..t (t = this step's own row), y = C . Wvo^T + bo -> GLU -> + residual
        if (h->self_abs_active) {
            void* qa = rowa(h, h->qabs, rc, 2048);
            void* ca = rowa(h, h->cabs, rc, 2048);
            GemmArgs gq = mk_gemm(xnbuf, 256, h->dec_self[l].wqk, 256, qa, 2048, rows, 2048, 256, EPI_STORE, h->dt, h->dt, nullptr, nullptr, 0);
            LAUNCH(KC_DEC_GEMM_Q, 1, (double)rows * 256 * e + 2048.0 * 256 * e + (double)rows * 2048 * e, gemm_flops(gq), run_gemm(h, gq, st));
            AttnAbsArgs ab{};
            ab.q = qa; ab.ldq = 2048; ab.latent = (char*)h->latcache.p + ((size_t)l * B + row0) * tcap * 256 * e; ab.latent_rows = (long)rows * tcap;
            ab.znew = xnbuf; ab.ldz = 256; ab.tcap = tcap; ab.step = step; ab.o = ca; ab.ldo = 2048; ab.batch = rows;
            if (h->attn_trace_on && h->attn_trace.p && 2 * l + 1 < 8) {
                ab.trace = h->attn_trace.as<unsigned long long>() + (size_t)branch * 3 * 2048; ab.trace_step = step; ab.trace_k = 2 * l;
                ab.dbg = h->attn_trace.as<unsigned long long>() + (size_t)MAX_BRANCH * 3 * 2048;
            }
            if (!(h->dbg_skip & 1))
                LAUNCH(KC_DEC_ATTN_SELF, 1, (double)rows * tkeys * 256 * e, 4.0 * rows * tkeys * 2048,
                       (g_attn_seq ? launch_attn_seq(ab, h->num_sms * h->attn_ctas_per_sm, st) : launch_attn_abs(ab, h->num_sms * h->attn_ctas_per_sm, st)));
            if ((r = sub_abs_out(h, rc, h->dec_self[l], ca, st))) return r;
        } else {
        // ---- causal self-attention over the KV cache
        {
            GemmArgs gq = mk_gemm(xnbuf, 256, h->dec_self[l].wqkv, 256, qb, 1536, rows, 1536, 256, EPI_STORE, h->dt, h->dt, nullptr, nullptr, 0);
            LAUNCH(KC_DEC_GEMM, 1, gemm_bytes(gq, e), gemm_flops(gq), run_gemm(h, gq, st));
        }
        // KV cache, head-major: [layer][sequence][head][key][K 64 | V 64] -- every (sequence, head) is one contiguous stream
        AttnDecodeArgs ad{};
        char* kv = (char*)h->kvcache.p + ((size_t)l * B + row0) * tcap * 1024 * e;
        ad.q = qb; ad.ldq = 1536; ad.knew = qb + 512 * e; ad.vnew = qb + 1024 * e; ad.ldnew = 1536;
        ad.kcache = kv; ad.vcache = kv + 64 * e; ad.ldkv = 128; ad.batch_stride = (int64_t)tcap * 1024; ad.head_stride = (int64_t)tcap * 128;
        ad.step = step; ad.o = rowa(h, h->o, rc, 512); ad.ldo = 512; ad.batch = rows; ad.dt = h->dt;
        const KvLayout lay_self{kv, (long)rows * 8 * tcap, 128, 128, 0, 0, 64, tcap, 8 * tcap};
        if (h->attn_trace_on && h->attn_trace.p && 2 * l + 1 < 8) {
            ad.trace = h->attn_trace.as<unsigned long long>() + (size_t)branch * 3 * 2048; ad.trace_step = step; ad.trace_k = 2 * l;
        }
        if (h->dbg_skip & 1) {}
        else if ((h->use_tma_attn == 1 || h->use_tma_attn == 2) && attn_decode_tma_supported(ad))
            LAUNCH(KC_DEC_ATTN_SELF, 1, (double)rows * tkeys * 1024 * e, 4.0 * rows * tkeys * 512,
                   launch_attn_decode_tma(ad, lay_self, h->num_sms * h->attn_ctas_per_sm, st));
        else
            LAUNCH(KC_DEC_ATTN_SELF, 1, (double)rows * tkeys * 1024 * e, 4.0 * rows * tkeys * 512, launch_attn_decode(ad, tcap, st));
        if ((r = sub_attn_out(h, rc, h->dec_self[l], st))) return r;
        }
        // ---- cross-attention, absorbed form: Q' = LN.LN(s) . Wqk^T (8 x 256 per row), C_h = softmax(Q'_h . enc^T / 8) . enc over the
        // bf16 encoder memory itself, y = C . Wvo^T + bo -> GLU -> + residual
        if (h->dec_enc) {
            if ((r = sub_norm(h, rc, false, nullptr, nullptr, nullptr, nullptr, st))) return r;
            void* qa = rowa(h, h->qabs, rc, 2048);
            void* ca = rowa(h, h->cabs, rc, 2048);
            GemmArgs gc = mk_gemm(xnbuf, 256, h->dec_cross[l].wqk, 256, qa, 2048, rows, 2048, 256, EPI_STORE, h->dt, h->dt, nullptr, nullptr, 0);
            LAUNCH(KC_DEC_GEMM_Q, 1, (double)rows * 256 * e + 2048.0 * 256 * e + (double)rows * 2048 * e, gemm_flops(gc), run_gemm(h, gc, st));
            AttnAbsArgs ab{};
            ab.q = qa; ab.ldq = 2048; ab.latent = h->dec_enc; ab.latent_rows = h->crosskv_rows; ab.k_off = d_enc_off + row0; ab.o = ca; ab.ldo = 2048; ab.batch = rows;
            // equal memory lengths (B x max_s tokens in total <=> every sequence has max_s): sequence b's rows start at b * max_s
            if ((long)B * max_s == (long)h->crosskv_rows) {
                ab.uni_nk = max_s;
                ab.latent = (const char*)h->dec_enc + (size_t)row0 * max_s * 256 * e; ab.latent_rows = (long)rows * max_s;
            }
            if (h->attn_trace_on && h->attn_trace.p && 2 * l + 1 < 8) {
                ab.trace = h->attn_trace.as<unsigned long long>() + (size_t)branch * 3 * 2048; ab.trace_step = step; ab.trace_k = 2 * l + 1;
                ab.dbg = h->attn_trace.as<unsigned long long>() + (size_t)MAX_BRANCH * 3 * 2048 + 8;
            }
            if (!(h->dbg_skip & 2))
                LAUNCH(KC_DEC_ATTN_CROSS, 1, sum_s * rows / B * 256 * e, 4.0 * sum_s * rows / B * 2048,
                       (g_attn_seq ? launch_attn_seq(ab, h->num_sms * h->attn_ctas_per_sm, st) : launch_attn_abs(ab, h->num_sms * h->attn_ctas_per_sm, st)));
            if ((r = sub_abs_out(h, rc, h->dec_cross[l], ca, st))) return r;
        } else {
        // ---- cross-attention over the (pre-projected) encoder memory; q goes to the first 512 columns of this branch's qkv rows
        {
            if ((r = sub_norm(h, rc, false, nullptr, nullptr, nullptr, nullptr, st))) return r;
            GemmArgs gc = mk_gemm(xnbuf, 256, h->dec_cross[l].wq, 256, qb, 512, rows, 512, 256, EPI_STORE, h->dt, h->dt, nullptr, nullptr, 0);
            LAUNCH(KC_DEC_GEMM, 1, gemm_bytes(gc, e), gemm_flops(gc), run_gemm(h, gc, st));
        }
        // encoder-memory K/V, head-major: [layer][head][token][K 64 | V 64]
        AttnDecodeArgs ac{};
        const long ntok_all = h->crosskv_rows;
        char* ckv = (char*)h->crosskv_hm.p + (size_t)l * 8 * ntok_all * 128 * e;
        ac.q = qb; ac.ldq = 512; ac.kcache = ckv; ac.vcache = ckv + 64 * e; ac.ldkv = 128; ac.head_stride = ntok_all * 128;
        ac.k_off = d_enc_off + row0; ac.o = rowa(h, h->o, rc, 512); ac.ldo = 512; ac.batch = rows; ac.dt = h->dt;
        const KvLayout lay_cross{ckv, 8 * ntok_all, 128, 128, 0, 0, 64, (int)ntok_all, 0};
        if (h->attn_trace_on && h->attn_trace.p && 2 * l + 1 < 8) {
            ac.trace = h->attn_trace.as<unsigned long long>() + (size_t)branch * 3 * 2048; ac.trace_step = step; ac.trace_k = 2 * l + 1;
        }
        if (h->dbg_skip & 2) {}
        else if ((h->use_tma_attn == 1 || h->use_tma_attn == 3) && attn_decode_tma_supported(ac))
            LAUNCH(KC_DEC_ATTN_CROSS, 1, sum_s * rows / B * 1024 * e, 4.0 * sum_s * rows / B * 512,
                   launch_attn_decode_tma(ac, lay_cross, h->num_sms * h->attn_ctas_per_sm, st));
        else
            LAUNCH(KC_DEC_ATTN_CROSS, 1, sum_s * rows / B * 1024 * e, 4.0 * sum_s * rows / B * 512, launch_attn_decode(ac, max_s, st));
        if ((r = sub_attn_out(h, rc, h->dec_cross[l], st))) return r;
        }
        // ---- GeGLU MLP
        if ((r = sub_norm(h, rc, false, nullptr, nullptr, nullptr, nullptr, st))) return r;
        if ((r = sub_mlp(h, rc, h->dec_mlp[l], st))) return r;
        if ((r = sub_norm(h, rc, l == L - 1, h->dec_norm_g, h->dec_norm_b, nullptr, xnbuf, st))) return r;
    }
    float* lg = h->logits.as<float>() + (size_t)row0 * c.vocab_size;
    // greedy bf16 tier: the vocabulary GEMM reduces every 32-column tile to (max, first index) in its epilogue and the token
    // kernel finishes the argmax over those partials -- the logits never reach HBM (model/decoder.py:60,103 takes
    // logits[:, -1] of a full (B, T, V) tensor).  Sampling and the keep_logits debug option need the row of logits itself.
    const int nparts = (c.vocab_size + 31) / 32;
    const bool fused_amax = h->dt == DT_BF16 && h->use_tcgen05 && h->samp_temp <= 0.0 && !h->keep_logits && h->amax_part.p;
    float2* parts = fused_amax ? h->amax_part.as<float2>() + (size_t)row0 * nparts : nullptr;
    if (fused_amax) {
        GemmArgs gl = mk_gemm(xnbuf, 256, h->w_logits, 256, parts, nparts, rows, c.vocab_size, 256, EPI_ARGMAX, h->dt, DT_F32, h->b_logits, nullptr, 0);
        LAUNCH(KC_DEC_GEMM_LOGITS, 1, (double)rows * 256 * e + (double)c.vocab_size * 256 * e + (double)rows * nparts * 8, gemm_flops(gl), run_gemm(h, gl, st));
    } else {
        GemmArgs gl = mk_gemm(xnbuf, 256, h->w_logits, 256, lg, c.vocab_size, rows, c.vocab_size, 256, EPI_STORE, h->dt, DT_F32, h->b_logits, nullptr, 0);
        LAUNCH(KC_DEC_GEMM_LOGITS, 1, gemm_bytes(gl, e), gemm_flops(gl), run_gemm(h, gl, st));
    }
    ArgmaxArgs aa{};
    aa.logits = lg; aa.partials = parts; aa.nparts = nparts; aa.B = rows; aa.V = c.vocab_size; aa.out_ids = h->out_ids.as<int64_t>() + (size_t)row0 * tcap; aa.out_ld = tcap;
    aa.cur_tok = ds.cur_tok + row0; aa.step = step; aa.seen_eos = ds.seen + row0; aa.done_step = ds.done_step + branch;
    aa.block_counter = ds.block_counter + branch; aa.eos = eos;
    aa.tok_emb = h->tok_emb; aa.pos_emb = h->pos_emb; aa.emb_dt = h->dt; aa.emb_max_pos = tcap;
    aa.emb_x = xbuf; aa.emb_xn = xnbuf; aa.emb_g = h->dec_ln_g; aa.emb_b = h->dec_ln_b;
    if (h->samp_temp > 0.0) {
        aa.topk = sampling_k(h); aa.inv_temp = (float)(1.0 / h->samp_temp); aa.seed = h->samp_seed; aa.row_base = row0; aa.call_ctr = ds.call_ctr;
    }
    LAUNCH(KC_DEC_ARGMAX, 1, fused_amax ? (double)rows * nparts * 8 : (double)rows * c.vocab_size * 4, 0.0, launch_argmax_step(aa, st));
    return 0;
}

// Input of the first decode step of a branch (later steps get theirs from the token kernel of the step before).
static int enqueue_first_embed(texocr_handle* h, int B, int row0, int rows, int branch, cudaStream_t st) {
    const texocr_config& c = h->cfg;
    DecState ds = dec_state(h, B);
    RowCtx rc{rows, KC_DEC_GEMM, KC_DEC_ROW, h->dec_ln_g, h->dec_ln_b, row0};
    LAUNCH(KC_DEC_ROW, 1, (double)rows * 256 * (8 + 4 + h->esz), 0.0,
           launch_embed_ln(ds.cur_tok + row0, ds.step + branch, 1, rows, h->tok_emb, h->pos_emb, c.vocab_size, h->dec_ln_g, h->dec_ln_b,
                           rowf(h->x, rc, 256), rowa(h, h->xn, rc, 256), h->dt, st));
    return 0;
}

// k of the reference's top-k filter: int((1 - threshold) * vocab) in double arithmetic, as Python evaluates it (utils.py:87)
int sampling_k(const texocr_handle* h) { return (int)((1.0 - h->samp_threshold) * (double)h->cfg.vocab_size); }

struct BranchPlan { int n; int row0[MAX_BRANCH]; int rows[MAX_BRANCH]; };
static BranchPlan plan_branches(texocr_handle* h, int B) {
    // one 128-row GEMM tile per branch (4 branches at B = 512: measured 85.4 ms per generate vs 88.6 with 3, 89.3 with 6 and 94.2 with 8
    // at the end of round 2; round 1, before the 8-warp epilogues and the late PDL release, preferred ~86 rows), at most 8
    int n = h->decode_branches > 0 ? h->decode_branches : std::min(8, std::max(1, (B + 64) / 128));
    n = std::max(1, std::min(std::min(n, MAX_BRANCH), B));
    BranchPlan p;
    p.n = n;
    const int per = (B + n - 1) / n;
    for (int i = 0; i < n; ++i) { p.row0[i] = std::min(B, i * per); p.rows[i] = std::min(B, (i + 1) * per) - p.row0[i]; }
    while (p.n > 1 && p.rows[p.n - 1] <= 0) --p.n;
    return p;
}

// all branches of one decode step: sequentially on `st` (eager / profiling) or forked onto the branch streams (graph capture)
static int enqueue_all_branches(texocr_handle* h, const BranchPlan& bp, bool fork, int B, int tcap, int eos, const int* d_enc_off,
                                int max_s, double sum_s, int t_host, cudaStream_t st) {
    int r;
    if (!fork || bp.n == 1) {
        for (int i = 0; i < bp.n; ++i)
            if ((r = enqueue_decode_step(h, B, bp.row0[i], bp.rows[i], i, tcap, eos, d_enc_off, max_s, sum_s, t_host, st))) return r;
        return 0;
    }
    CK(cudaEventRecord(h->fork_ev, st));
    for (int i = 1; i < bp.n; ++i) {
        CK(cudaStreamWaitEvent(h->branch_stream[i], h->fork_ev, 0));
        if ((r = enqueue_decode_step(h, B, bp.row0[i], bp.rows[i], i, tcap, eos, d_enc_off, max_s, sum_s, t_host, h->branch_stream[i]))) return r;
        CK(cudaEventRecord(h->join_ev[i], h->branch_stream[i]));
    }
    if ((r = enqueue_decode_step(h, B, bp.row0[0], bp.rows[0], 0, tcap, eos, d_enc_off, max_s, sum_s, t_host, st))) return r;
    for (int i = 1; i < bp.n; ++i) CK(cudaStreamWaitEvent(st, h->join_ev[i], 0));
    return 0;
}

// enc memory must already be projected into h->crosskv; d_enc_off = per-row token offsets (device, B+1)
int run_generate(texocr_handle* h, const int64_t* d_start, int eos, const int* d_enc_off, int max_s, double sum_s, int B,
                        int max_len, int64_t* out_ids, int32_t* n_steps, cudaStream_t st) {
    const texocr_config& c = h->cfg;
    if (max_len <= 0) return fail(h, TEXOCR_ERR_ARG, "max_len must be positive");
    if (max_len > c.max_length)
        return fail(h, TEXOCR_ERR_ARG, "max_len %d > config max_length %d: the KV cache is position-indexed; the reference's "
                    "sliding-window regime (model/decoder.py:99-100) is not implemented", max_len, c.max_length);
    const int tcap = max_len;
    int r;
    if ((r = ensure_rows(h, B))) return r;
    const bool absorb = h->dec_enc != nullptr;
    if (absorb || h->self_abs_active) {
        ENSURE(h->qabs, (size_t)B * 2048 * h->esz);
        ENSURE(h->cabs, (size_t)B * 2048 * h->esz);
    }
    if (!absorb) {   // the decode loop streams the memory K/V per (sequence, head): re-lay the GEMM output head-major, once
        const int ntok = h->crosskv_rows;
        ENSURE(h->crosskv_hm, (size_t)ntok * c.dec_layers * 1024 * h->esz);
        LAUNCH(KC_MISC, 1, (double)ntok * c.dec_layers * 1024 * h->esz * 2, 0.0,
               launch_crosskv_head_major(h->crosskv.p, h->crosskv_hm.p, ntok, c.dec_layers, h->dt, st));
    }
    ENSURE(h->logits, (size_t)B * c.vocab_size * 4);
    if (h->dt == DT_BF16 && h->use_tcgen05) ENSURE(h->amax_part, (size_t)B * ((c.vocab_size + 31) / 32) * 8);
    if (h->self_abs_active) ENSURE(h->latcache, (size_t)c.dec_layers * B * tcap * 256 * h->esz);
    else ENSURE(h->kvcache, (size_t)c.dec_layers * B * tcap * 1024 * h->esz);
    const void* kv_key = h->self_abs_active ? h->latcache.p : h->kvcache.p;
    ENSURE(h->dec_state, dec_state_bytes(B));
    ENSURE(h->out_ids, (size_t)B * tcap * 8);
    if (h->attn_trace_on) {
        ENSURE(h->attn_trace, (size_t)MAX_BRANCH * 3 * 2048 * 8 + 512);
        CK(cudaMemsetAsync((char*)h->attn_trace.p + (size_t)MAX_BRANCH * 3 * 2048 * 8, 0, 512, st));
        for (int i = 0; i < MAX_BRANCH; ++i) {
            char* base = (char*)h->attn_trace.p + (size_t)i * 3 * 2048 * 8;
            CK(cudaMemsetAsync(base, 0xff, 2 * 2048 * 8, st));
            CK(cudaMemsetAsync(base + 2 * 2048 * 8, 0, 2048 * 8, st));
        }
    }
    if (!h->h_poll) CK(cudaMallocHost(&h->h_poll, 4 * MAX_BRANCH * 4));
    DecState ds = dec_state(h, B);
    CK(cudaMemsetAsync((char*)h->dec_state.p + (size_t)B * 8, 0, dec_state_bytes(B) - (size_t)B * 8, st));
    CK(cudaMemcpyAsync(ds.cur_tok, d_start, (size_t)B * 8, cudaMemcpyDeviceToDevice, st));
    if (h->samp_temp > 0.0) {      // every sampled generate call draws from a fresh Philox sub-stream
        h->h_poll[3 * MAX_BRANCH] = (int)h->samp_calls++;
        CK(cudaMemcpyAsync(ds.call_ctr, &h->h_poll[3 * MAX_BRANCH], 4, cudaMemcpyHostToDevice, st));
    }
    const BranchPlan bp = plan_branches(h, B);
    for (int i = 0; i < bp.n; ++i) {
        if (!h->branch_stream[i]) {
            CK(cudaStreamCreateWithFlags(&h->branch_stream[i], cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&h->join_ev[i], cudaEventDisableTiming));
        }
    }
    h->own_stream2 = h->branch_stream[0];
    if (!h->fork_ev) CK(cudaEventCreateWithFlags(&h->fork_ev, cudaEventDisableTiming));

    const bool graph_ok = h->use_graph && !h->prof_on;
    // Stream of branch i: its own non-blocking stream (branch 0 included when there are several branches), so the
    // branches run as independent, phase-shifted pipelines that only meet again at the end of the call.
    cudaStream_t bst[MAX_BRANCH];
    const bool own = graph_ok && bp.n > 1;
    for (int i = 0; i < bp.n; ++i) bst[i] = !own ? st : (i == 0 ? h->own_stream2 : h->branch_stream[i]);
    const void* ckv_key = absorb ? h->dec_enc : h->crosskv_hm.p;
    const int samp_key = h->samp_temp > 0.0 ? sampling_k(h) : 0;      // seed and temperature are compared in full (gkey.samp_seed / samp_temp)
    const uint64_t seed_key = h->samp_temp > 0.0 ? h->samp_seed : 0;
    const double temp_key = h->samp_temp > 0.0 ? h->samp_temp : 0.0;
    if (graph_ok) {
        const bool hit = h->graph_exec && h->gkey.B == B && h->gkey.tcap == tcap && h->gkey.eos == eos && h->gkey.max_s == max_s &&
                         h->gkey.samp == samp_key && h->gkey.samp_seed == seed_key && h->gkey.samp_temp == temp_key && h->gkey.enc_off == d_enc_off &&
                         h->gkey.ntok == h->crosskv_rows && h->gkey.kv == kv_key && h->gkey.ckv == ckv_key && h->gkey.x == h->x.p && h->gkey.nb == bp.n;
        if (!hit) {
            std::lock_guard<std::recursive_mutex> lk(g_dev_mu);
            drop_graphs(h);
            const int64_t before = h->launches;
            for (int i = 0; i < bp.n; ++i) {
                CK(cudaStreamBeginCapture(bst[i], cudaStreamCaptureModeRelaxed));
                r = enqueue_decode_step(h, B, bp.row0[i], bp.rows[i], i, tcap, eos, d_enc_off, max_s, sum_s, -1, bst[i]);
                cudaError_t ce = cudaStreamEndCapture(bst[i], &h->bgraph[i]);
                if (r) return r;
                CK(ce);
                CK(cudaGraphInstantiate(&h->bgraph_exec[i], h->bgraph[i], 0));
            }
            h->graph = h->bgraph[0]; h->graph_exec = h->bgraph_exec[0];
            h->gkey.kernels = (int)(h->launches - before) / bp.n;
            h->launches = before;        // capture does not execute
            h->gkey.samp = samp_key; h->gkey.samp_seed = seed_key; h->gkey.samp_temp = temp_key; h->gkey.enc_off = d_enc_off;
            h->gkey.B = B; h->gkey.tcap = tcap; h->gkey.eos = eos; h->gkey.max_s = max_s; h->gkey.nb = bp.n;
            h->gkey.kv = const_cast<void*>(kv_key); h->gkey.ckv = const_cast<void*>(ckv_key); h->gkey.x = h->x.p; h->gkey.ntok = h->crosskv_rows;
        }
        if (own) {      // fork: every branch stream waits for the work already queued on st, then starts with its phase shift
            CK(cudaEventRecord(h->fork_ev, st));
            for (int i = 0; i < bp.n; ++i) {
                CK(cudaStreamWaitEvent(bst[i], h->fork_ev, 0));
                if (i > 0 && h->stagger_us > 0) CK(launch_delay((long)i * h->stagger_us * 1000L, bst[i]));
            }
        }
    }
    for (int s2 = 0; s2 < 2; ++s2)
        for (int i = 0; i < bp.n; ++i)
            if (!h->poll_ev[s2][i]) CK(cudaEventCreateWithFlags(&h->poll_ev[s2][i], cudaEventDisableTiming | cudaEventBlockingSync));
    for (int i = 0; i < bp.n; ++i)
        if ((r = enqueue_first_embed(h, B, bp.row0[i], bp.rows[i], i, bst[i]))) return r;
    // Host runs ahead of the device by at most 2*POLL steps; an early exit costs at most that many extra steps.
    const int POLL = 16;
    int issued = 0, polls = 0;
    bool stop = false;
    for (int t = 0; t < max_len && !stop; ++t) {
        if (graph_ok) {
            for (int i = 0; i < bp.n; ++i) { CK(cudaGraphLaunch(h->bgraph_exec[i], bst[i])); h->launches += h->gkey.kernels; }
        } else if ((r = enqueue_all_branches(h, bp, false, B, tcap, eos, d_enc_off, max_s, sum_s, t, st))) return r;
        ++issued;
        if (eos >= 0 && issued % POLL == 0 && t + 1 < max_len) {
            const int slot = polls & 1;
            if (polls >= 1) {      // wait for the PREVIOUS poll (issued POLL steps ago), keeps the queues non-empty
                bool all = true;
                for (int i = 0; i < bp.n; ++i) {
                    CK(cudaEventSynchronize(h->poll_ev[slot ^ 1][i]));
                    all = all && h->h_poll[(slot ^ 1) * MAX_BRANCH + i] > 0;
                }
                if (all) stop = true;
            }
            for (int i = 0; i < bp.n; ++i) {
                CK(cudaMemcpyAsync(&h->h_poll[slot * MAX_BRANCH + i], ds.done_step + i, 4, cudaMemcpyDeviceToHost, bst[i]));
                CK(cudaEventRecord(h->poll_ev[slot][i], bst[i]));
            }
            ++polls;
        }
    }
    if (own) {      // join
        for (int i = 0; i < bp.n; ++i) {
            CK(cudaEventRecord(h->join_ev[i], bst[i]));
            CK(cudaStreamWaitEvent(st, h->join_ev[i], 0));
        }
    }
    CK(cudaMemcpyAsync(&h->h_poll[2 * MAX_BRANCH], ds.done_step, MAX_BRANCH * 4, cudaMemcpyDeviceToHost, st));
    if ((r = from_device(h, out_ids, h->out_ids.p, (size_t)B * tcap * 8, st))) return r;
    // The host waits on blocking-sync events (the thread sleeps instead of spinning): with several batches in flight per GPU and
    // several ranks per box there are more waiting host threads than cores.
    if (!h->done_ev) CK(cudaEventCreateWithFlags(&h->done_ev, cudaEventDisableTiming | cudaEventBlockingSync));
    CK(cudaEventRecord(h->done_ev, st));
    CK(cudaEventSynchronize(h->done_ev));
    // every row holds an EOS once every branch has seen one in all of its rows: the LAST branch to finish decides
    int done = 0;
    bool all_done = true;
    for (int i = 0; i < bp.n; ++i) {
        const int d = h->h_poll[2 * MAX_BRANCH + i];
        all_done = all_done && d > 0;
        done = std::max(done, d);
    }
    *n_steps = all_done ? done : max_len;
    return 0;
}
