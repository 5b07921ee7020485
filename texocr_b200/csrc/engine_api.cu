// C ABI of libtexocr_b200.so (include/texocr.h).  What runs where (reference file:line in brackets):
//   texocr_encode            encoder [model/encoder.py:128-152, model/resnet.py]
//   texocr_decoder_logits    teacher-forced Transformer.forward [model/decoder.py:41-67]
//   texocr_decoder_generate  KV-cached generate loop over a given memory [model/decoder.py:77-122]
//   texocr_generate          encoder + loop [model/ocr_model.py:46-66]
// There is no CPU fallback anywhere: every entry point needs the sm_100 device the handle was created on.
#include "engine_internal.h"

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" {

const char* texocr_last_error(const texocr_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int texocr_create(const texocr_config* cfg, int device, texocr_handle** out) {
    texocr_handle* h = nullptr;
    if (!cfg || !out) return fail(h, TEXOCR_ERR_ARG, "null argument");
    *out = nullptr;
    if (cfg->abi_version != TEXOCR_ABI_VERSION) return fail(h, TEXOCR_ERR_ARG, "ABI version mismatch: header %d, library %d", cfg->abi_version, TEXOCR_ABI_VERSION);
    if (cfg->vocab_size <= 0 || cfg->vocab_size % 4) return fail(h, TEXOCR_ERR_ARG, "vocab_size must be a positive multiple of 4");
    if (cfg->max_length <= 0 || cfg->enc_layers <= 0 || cfg->dec_layers <= 0) return fail(h, TEXOCR_ERR_ARG, "bad layer counts / max_length");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return fail(h, TEXOCR_ERR_NODEVICE, "no CUDA device: texocr_b200 has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(h, TEXOCR_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(h, TEXOCR_ERR_NODEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    CK(cudaSetDevice(device));
    h = new texocr_handle();
    h->num_sms = prop.multiProcessorCount;
    h->cfg = *cfg; h->device = device;
    h->dt = cfg->precision == TEXOCR_BF16 ? DT_BF16 : DT_F32;
    h->esz = h->dt == DT_BF16 ? 2 : 4;
    cudaError_t e = cudaEventCreateWithFlags(&h->geom_ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->hop_in, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->hop_out, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete h; return fail(nullptr, TEXOCR_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(e)); }
    *out = h;
    return 0;
}

void texocr_destroy(texocr_handle* h) {
    if (!h) return;
    std::lock_guard<std::recursive_mutex> lk(g_dev_mu);
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    drop_graphs(h);
    for (int s2 = 0; s2 < 2; ++s2) for (int i = 0; i < 16; ++i) if (h->poll_ev[s2][i]) cudaEventDestroy(h->poll_ev[s2][i]);
    for (void* p : h->weight_allocs) cudaFree(p);
    DevBuf* bufs[] = {&h->geom, &h->img_stage, &h->raw1, &h->act2, &h->actA, &h->actB, &h->rawMid, &h->actMid, &h->rawMid2, &h->actMid2,
                      &h->raw3, &h->rawDs, &h->gn_partial, &h->gn_stats[0], &h->gn_stats[1], &h->gn_stats[2], &h->gn_stats[3],
                      &h->proj_out, &h->patch_cols, &h->backbone_a, &h->col, &h->x, &h->s, &h->xn, &h->qkv, &h->o, &h->hid, &h->logits,
                      &h->enc_out, &h->enc_a, &h->crosskv, &h->crosskv_hm, &h->kvcache, &h->ids_stage, &h->mask_stage, &h->enc_stage, &h->tgt_stage,
                      &h->row_loss, &h->scalars, &h->dec_state, &h->out_ids, &h->attn_trace, &h->amax_part, &h->qabs, &h->cabs, &h->latcache, &h->prep_meta, &h->prep_in, &h->prep_out};
    for (DevBuf* b : bufs) if (b->p) cudaFree(b->p);
    if (h->h_geom) cudaFreeHost(h->h_geom);
    if (h->h_poll) cudaFreeHost(h->h_poll);
    if (h->h_bos) cudaFreeHost(h->h_bos);
    if (h->geom_ev) cudaEventDestroy(h->geom_ev);
    if (h->done_ev) cudaEventDestroy(h->done_ev);
    if (h->hop_in) cudaEventDestroy(h->hop_in);
    if (h->hop_out) cudaEventDestroy(h->hop_out);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    for (int i = 0; i < 16; ++i) {
        if (h->branch_stream[i]) cudaStreamDestroy(h->branch_stream[i]);
        if (h->join_ev[i]) cudaEventDestroy(h->join_ev[i]);
    }
    if (h->fork_ev) cudaEventDestroy(h->fork_ev);
    for (auto& p : h->prof) { cudaEventDestroy(p.e0); cudaEventDestroy(p.e1); }
    for (auto e : h->ev_pool) cudaEventDestroy(e);
    delete h;
}

int texocr_set_weight(texocr_handle* h, const char* name, const float* data, int32_t ndim, const int64_t* shape) {
    if (!h || !name || !data || ndim < 0 || ndim > 8) return fail(h, TEXOCR_ERR_ARG, "bad argument to texocr_set_weight");
    if (h->finalized) return fail(h, TEXOCR_ERR_STATE, "weights already finalised; create a new handle to load other weights");
    CK(cudaSetDevice(h->device));
    HostTensor t;
    t.shape.assign(shape, shape + ndim);
    const int64_t n = t.numel();
    if (n <= 0 || n > (int64_t)1 << 31) return fail(h, TEXOCR_ERR_ARG, "bad shape for '%s'", name);
    t.data.resize((size_t)n);
    if (is_device_ptr(data)) CK(cudaMemcpy(t.data.data(), data, (size_t)n * 4, cudaMemcpyDeviceToHost));
    else memcpy(t.data.data(), data, (size_t)n * 4);
    h->sd[name] = std::move(t);
    return 0;
}

int texocr_finalize_weights(texocr_handle* h) {
    if (!h) return TEXOCR_ERR_ARG;
    if (h->finalized) return fail(h, TEXOCR_ERR_STATE, "weights already finalised");
    CK(cudaSetDevice(h->device));
    return finalize_weights(h);
}

// The legacy / per-thread default streams cannot be captured into a CUDA graph, so work submitted on them hops to
// the handle's own non-blocking stream: it waits for everything already queued on the caller's stream, and the
// caller's stream waits for it on exit -- stream-ordering as seen by the caller is unchanged.
struct StreamHop {
    texocr_handle* h; cudaStream_t user, work; bool hop;
    StreamHop(texocr_handle* h_, void* stream) : h(h_), user((cudaStream_t)stream), work((cudaStream_t)stream), hop(false) {
        if (user == nullptr || user == cudaStreamLegacy || user == cudaStreamPerThread) {
            hop = true;
            work = h->own_stream;
            cudaEventRecord(h->hop_in, user);
            cudaStreamWaitEvent(work, h->hop_in, 0);
        }
    }
    ~StreamHop() {
        if (hop) {
            cudaEventRecord(h->hop_out, work);
            cudaStreamWaitEvent(user, h->hop_out, 0);
        }
    }
};

#define ENTRY_CHECKS()                                                                                     \
    if (!h) return TEXOCR_ERR_ARG;                                                                         \
    if (!h->finalized) return fail(h, TEXOCR_ERR_STATE, "weights not finalised (texocr_finalize_weights)"); \
    CK(cudaSetDevice(h->device));                                                                          \
    StreamHop hop__(h, stream);                                                                            \
    cudaStream_t st = hop__.work

static long total_pixels(const int32_t* hw, int B) {
    long n = 0;
    for (int b = 0; b < B; ++b) n += (long)hw[2 * b] * hw[2 * b + 1];
    return n;
}

int texocr_encode(texocr_handle* h, const float* images, const int32_t* hw, int32_t batch, float* enc_out, void* stream) {
    ENTRY_CHECKS();
    if (!images || !hw || !enc_out) return fail(h, TEXOCR_ERR_ARG, "null argument");
    EncGeom g;
    int r;
    if ((r = plan_geometry(h, hw, batch, g, st))) return r;
    const void* d_img = nullptr;
    if ((r = to_device(h, images, (size_t)total_pixels(hw, batch) * 4, h->img_stage, &d_img, st))) return r;
    if ((r = run_encoder(h, (const float*)d_img, g, st))) return r;
    if ((r = from_device(h, enc_out, h->enc_out.p, (size_t)g.ntok * 256 * 4, st))) return r;
    if (!is_device_ptr(enc_out)) CK(cudaStreamSynchronize(st));
    return 0;
}

static int memory_offsets(texocr_handle* h, const int32_t* enc_len, int B, std::vector<int>& off, int* max_s) {
    off.assign(B + 1, 0);
    *max_s = 0;
    for (int b = 0; b < B; ++b) {
        if (enc_len[b] <= 0) return fail(h, TEXOCR_ERR_ARG, "enc_len[%d] must be positive", b);
        off[b + 1] = off[b] + enc_len[b];
        *max_s = std::max(*max_s, (int)enc_len[b]);
    }
    return 0;
}

int texocr_decoder_logits(texocr_handle* h, const int64_t* ids, const uint8_t* mask, const float* enc, const int32_t* enc_len,
                          int32_t batch, int32_t T, float* logits_out, void* stream) {
    ENTRY_CHECKS();
    const texocr_config& c = h->cfg;
    if (!ids || !enc || !enc_len || !logits_out || batch <= 0 || T <= 0) return fail(h, TEXOCR_ERR_ARG, "bad argument");
    if (T > c.max_length) return fail(h, TEXOCR_ERR_ARG, "T %d exceeds the positional table (max_length %d)", T, c.max_length);
    const int B = batch, L = c.dec_layers;
    const long R = (long)B * T;
    std::vector<int> enc_off;
    int max_s, r;
    if ((r = memory_offsets(h, enc_len, B, enc_off, &max_s))) return r;
    const int ntok = enc_off[B];
    std::vector<int> v;
    for (int b = 0; b <= B; ++b) v.push_back(b * T);
    v.insert(v.end(), enc_off.begin(), enc_off.end());
    if ((r = upload_ints(h, v, st))) return r;
    const int* d_row_off = h->geom.as<int>();
    const int* d_enc_off = d_row_off + (B + 1);
    const void *d_ids, *d_mask = nullptr, *d_enc;
    if ((r = to_device(h, ids, (size_t)R * 8, h->ids_stage, &d_ids, st))) return r;
    if (mask && (r = to_device(h, mask, (size_t)R, h->mask_stage, &d_mask, st))) return r;
    if ((r = to_device(h, enc, (size_t)ntok * 256 * 4, h->enc_stage, &d_enc, st))) return r;
    if ((r = ensure_rows(h, R))) return r;
    float* d_logits = logits_out;
    if (!is_device_ptr(logits_out)) { ENSURE(h->logits, (size_t)R * c.vocab_size * 4); d_logits = h->logits.as<float>(); }
    if ((r = run_crosskv(h, (const float*)d_enc, nullptr, ntok, st))) return r;

    RowCtx rc{(int)R, KC_TF_GEMM, KC_TF_ROW, h->dec_ln_g, h->dec_ln_b};
    LAUNCH(KC_TF_ROW, 1, (double)R * 256 * 12, 0.0,
           launch_embed_ln((const int64_t*)d_ids, nullptr, T, (int)R, h->tok_emb, h->pos_emb, c.vocab_size, h->dec_ln_g, h->dec_ln_b,
                           h->x.as<float>(), h->xn.p, h->dt, st));
    for (int l = 0; l < L; ++l) {
        GemmArgs gq = mk_gemm(h->xn.p, 256, h->dec_self[l].wqkv, 256, h->qkv.p, 1536, (int)R, 1536, 256, EPI_STORE, h->dt, h->dt, nullptr, nullptr, 0);
        LAUNCH(KC_TF_GEMM, 1, gemm_bytes(gq, h->esz), gemm_flops(gq), run_gemm(h, gq, st));
        AttnVarlenArgs av{};
        const char* base = (const char*)h->qkv.p;
        av.q = base; av.k = base + 512 * h->esz; av.v = base + 1024 * h->esz; av.ldq = av.ldk = av.ldv = 1536;
        av.o = h->o.p; av.ldo = 512; av.q_off = d_row_off; av.k_off = d_row_off; av.batch = B; av.max_q = T; av.causal = 1; av.dt = h->dt;
        av.q_mask = (const uint8_t*)d_mask; av.k_mask = (const uint8_t*)d_mask;
        LAUNCH(KC_TF_ATTN, 1, (double)R * 2048 * h->esz, 2.0 * B * (double)T * T * 512, launch_attn_varlen(av, st));
        if ((r = sub_attn_out(h, rc, h->dec_self[l], st))) return r;
        if ((r = sub_norm(h, rc, false, nullptr, nullptr, nullptr, nullptr, st))) return r;
        GemmArgs gc = mk_gemm(h->xn.p, 256, h->dec_cross[l].wq, 256, h->qkv.p, 512, (int)R, 512, 256, EPI_STORE, h->dt, h->dt, nullptr, nullptr, 0);
        LAUNCH(KC_TF_GEMM, 1, gemm_bytes(gc, h->esz), gemm_flops(gc), run_gemm(h, gc, st));
        AttnVarlenArgs ac{};
        const char* ckv = (const char*)h->crosskv.p + (size_t)l * 1024 * h->esz;
        ac.q = h->qkv.p; ac.ldq = 512; ac.k = ckv; ac.v = ckv + 512 * h->esz; ac.ldk = ac.ldv = L * 1024;
        ac.o = h->o.p; ac.ldo = 512; ac.q_off = d_row_off; ac.k_off = d_enc_off; ac.batch = B; ac.max_q = T; ac.causal = 0; ac.dt = h->dt;
        ac.q_mask = (const uint8_t*)d_mask;      // enc_mask is never passed by the reference: keys unmasked (model/attention.py:138-141)
        LAUNCH(KC_TF_ATTN, 1, (double)R * 512 * h->esz + (double)ntok * 1024 * h->esz, 4.0 * T * (double)ntok * 512, launch_attn_varlen(ac, st));
        if ((r = sub_attn_out(h, rc, h->dec_cross[l], st))) return r;
        if ((r = sub_norm(h, rc, false, nullptr, nullptr, nullptr, nullptr, st))) return r;
        if ((r = sub_mlp(h, rc, h->dec_mlp[l], st))) return r;
        if ((r = sub_norm(h, rc, l == L - 1, h->dec_norm_g, h->dec_norm_b, nullptr, h->xn.p, st))) return r;
    }
    GemmArgs gl = mk_gemm(h->xn.p, 256, h->w_logits, 256, d_logits, c.vocab_size, (int)R, c.vocab_size, 256, EPI_STORE, h->dt, DT_F32, h->b_logits, nullptr, 0);
    LAUNCH(KC_TF_GEMM, 1, gemm_bytes(gl, h->esz), gemm_flops(gl), run_gemm(h, gl, st));
    if (d_logits != logits_out) {
        if ((r = from_device(h, logits_out, d_logits, (size_t)R * c.vocab_size * 4, st))) return r;
        CK(cudaStreamSynchronize(st));
    }
    return 0;
}

int texocr_decoder_generate(texocr_handle* h, const int64_t* start_tokens, int32_t eos_tok, const float* enc, const int32_t* enc_len,
                            int32_t batch, int32_t max_len, int64_t* out_ids, int32_t* n_steps, void* stream) {
    ENTRY_CHECKS();
    if (!start_tokens || !enc || !enc_len || !out_ids || !n_steps || batch <= 0) return fail(h, TEXOCR_ERR_ARG, "bad argument");
    std::vector<int> enc_off;
    int max_s, r;
    if ((r = memory_offsets(h, enc_len, batch, enc_off, &max_s))) return r;
    const int ntok = enc_off[batch];
    if ((r = upload_ints(h, enc_off, st))) return r;
    const void *d_enc, *d_start;
    if ((r = to_device(h, enc, (size_t)ntok * 256 * 4, h->enc_stage, &d_enc, st))) return r;
    if ((r = to_device(h, start_tokens, (size_t)batch * 8, h->ids_stage, &d_start, st))) return r;
    if ((r = run_crosskv(h, (const float*)d_enc, nullptr, ntok, st, true))) return r;
    return run_generate(h, (const int64_t*)d_start, eos_tok, h->geom.as<int>(), max_s, (double)ntok, batch, max_len, out_ids, n_steps, st);
}

int texocr_generate(texocr_handle* h, const float* images, const int32_t* hw, int32_t batch, int32_t max_len, int64_t* out_ids,
                    int32_t* n_steps, void* stream) {
    ENTRY_CHECKS();
    if (!images || !hw || !out_ids || !n_steps) return fail(h, TEXOCR_ERR_ARG, "null argument");
    EncGeom g;
    int r;
    if (h->poison && (r = poison_workspaces(h, st))) return r;
    if ((r = plan_geometry(h, hw, batch, g, st))) return r;
    const void* d_img = nullptr;
    if ((r = to_device(h, images, (size_t)total_pixels(hw, batch) * 4, h->img_stage, &d_img, st))) return r;
    // start column = BOS for every row (model/ocr_model.py:57): pinned, constant content, copied on the work stream
    if (h->h_bos_cap < (size_t)batch) {
        std::lock_guard<std::recursive_mutex> lk(g_dev_mu);
        if (h->h_bos) { CK(cudaStreamSynchronize(st)); CK(cudaFreeHost(h->h_bos)); h->h_bos = nullptr; }
        h->h_bos_cap = std::max((size_t)batch, (size_t)1024);
        CK(cudaMallocHost(&h->h_bos, h->h_bos_cap * 8));
        for (size_t i = 0; i < h->h_bos_cap; ++i) h->h_bos[i] = (int64_t)h->cfg.bos_token;
    }
    ENSURE(h->ids_stage, (size_t)batch * 8);
    CK(cudaMemcpyAsync(h->ids_stage.p, h->h_bos, (size_t)batch * 8, cudaMemcpyHostToDevice, st));
    if ((r = run_encoder(h, (const float*)d_img, g, st))) return r;
    if ((r = run_crosskv(h, h->enc_out.as<float>(), h->dt == DT_F32 ? nullptr : h->enc_a.p, g.ntok, st, true))) return r;
    return run_generate(h, h->ids_stage.as<int64_t>(), h->no_early_exit ? -1 : h->cfg.eos_token, g.d_tok_off, g.max_tok, (double)g.ntok, batch, max_len, out_ids, n_steps, st);
}

int texocr_preprocess_u8(texocr_handle* h, const uint8_t* pixels, const int32_t* hwc, int32_t batch, int32_t pad_multiple,
                         float* out_images, int32_t* out_hw, void* stream) {
    if (!h) return TEXOCR_ERR_ARG;
    if (!pixels || !hwc || !out_images || !out_hw || batch <= 0 || pad_multiple < 1) return fail(h, TEXOCR_ERR_ARG, "bad argument to texocr_preprocess_u8");
    CK(cudaSetDevice(h->device));
    StreamHop hop__(h, stream);
    cudaStream_t st = hop__.work;
    // meta block: long in_off[B] | long out_off[B] | int hwc[3B] | int out_hw[2B]
    std::vector<long> offs((size_t)2 * batch);
    std::vector<int> ohw((size_t)2 * batch);
    long in_bytes = 0, out_elems = 0, max_out = 0;
    for (int b = 0; b < batch; ++b) {
        const int H = hwc[3 * b], W = hwc[3 * b + 1], C = hwc[3 * b + 2];
        if (H <= 0 || W <= 0) return fail(h, TEXOCR_ERR_ARG, "image %d: bad size %d x %d", b, H, W);
        if (C != 1 && C != 3) return fail(h, TEXOCR_ERR_ARG, "image %d: %d channels; Grayscale (torchvision) takes 1 or 3", b, C);
        const int Hp = (H + pad_multiple - 1) / pad_multiple * pad_multiple, Wp = (W + pad_multiple - 1) / pad_multiple * pad_multiple;
        offs[b] = in_bytes; offs[batch + b] = out_elems;
        ohw[2 * b] = Hp; ohw[2 * b + 1] = Wp;
        in_bytes += (long)H * W * C; out_elems += (long)Hp * Wp;
        max_out = std::max(max_out, (long)Hp * Wp);
    }
    const size_t meta_bytes = (size_t)batch * (2 * sizeof(long) + 5 * sizeof(int));
    ENSURE(h->prep_meta, meta_bytes);
    char* mp = (char*)h->prep_meta.p;
    CK(cudaMemcpyAsync(mp, offs.data(), (size_t)2 * batch * sizeof(long), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(mp + (size_t)2 * batch * sizeof(long), hwc, (size_t)3 * batch * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(mp + (size_t)2 * batch * sizeof(long) + (size_t)3 * batch * sizeof(int), ohw.data(), (size_t)2 * batch * sizeof(int), cudaMemcpyHostToDevice, st));
    int r;
    const void* d_in = nullptr;
    if ((r = to_device(h, pixels, (size_t)in_bytes, h->prep_in, &d_in, st))) return r;
    float* d_out = out_images;
    const bool out_dev = is_device_ptr(out_images);
    if (!out_dev) { ENSURE(h->prep_out, (size_t)out_elems * 4); d_out = h->prep_out.as<float>(); }
    const long* d_off = (const long*)mp;
    const int* d_hwc = (const int*)(mp + (size_t)2 * batch * sizeof(long));
    LAUNCH(KC_MISC, 1, (double)in_bytes + (double)out_elems * 4, 0.0,
           launch_preprocess_u8((const uint8_t*)d_in, d_off, d_hwc, d_off + batch, d_hwc + 3 * batch, d_out, batch, max_out, st));
    if (!out_dev) { if ((r = from_device(h, out_images, d_out, (size_t)out_elems * 4, st))) return r; }
    CK(cudaStreamSynchronize(st));        // the host vectors above are staged from pageable memory
    memcpy(out_hw, ohw.data(), (size_t)2 * batch * sizeof(int));
    return 0;
}

int texocr_cross_entropy(texocr_handle* h, const float* logits, const int64_t* targets, int64_t rows, float* loss_out, void* stream) {
    ENTRY_CHECKS();
    if (!logits || !targets || !loss_out || rows <= 0) return fail(h, TEXOCR_ERR_ARG, "bad argument");
    const int V = h->cfg.vocab_size;
    int r;
    const void *d_logits, *d_tgt;
    if ((r = to_device(h, logits, (size_t)rows * V * 4, h->logits, &d_logits, st))) return r;
    if ((r = to_device(h, targets, (size_t)rows * 8, h->tgt_stage, &d_tgt, st))) return r;
    ENSURE(h->row_loss, (size_t)rows * 4);
    ENSURE(h->scalars, 64);
    float* d_loss = is_device_ptr(loss_out) ? loss_out : h->scalars.as<float>();
    LAUNCH(KC_TF_ROW, 2, (double)rows * V * 4, 0.0, launch_cross_entropy((const float*)d_logits, (const int64_t*)d_tgt, rows, V, h->row_loss.as<float>(), d_loss, st));
    if (d_loss != loss_out) {
        CK(cudaMemcpyAsync(loss_out, d_loss, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return 0;
}

int64_t texocr_kernel_launches(const texocr_handle* h) { return h ? h->launches : 0; }

int texocr_profile_enable(texocr_handle* h, int32_t on) {
    if (!h) return TEXOCR_ERR_ARG;
    h->prof_on = on != 0;
    return 0;
}

int texocr_profile_read(texocr_handle* h, texocr_profile_row* rows, int32_t cap) {
    if (!h || !rows) return TEXOCR_ERR_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    for (auto& p : h->prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) {
            h->prof_ms[p.cls] += ms; h->prof_bytes[p.cls] += p.bytes; h->prof_flops[p.cls] += p.flops; h->prof_n[p.cls] += 1;
        }
        h->ev_pool.push_back(p.e0); h->ev_pool.push_back(p.e1);
    }
    h->prof.clear();
    int n = 0;
    for (int k = 0; k < KC_COUNT && n < cap; ++k) {
        if (!h->prof_n[k]) continue;
        memset(&rows[n], 0, sizeof rows[n]);
        strncpy(rows[n].name, kclass_name[k], sizeof rows[n].name - 1);
        rows[n].launches = h->prof_n[k]; rows[n].ms = h->prof_ms[k]; rows[n].bytes = h->prof_bytes[k]; rows[n].flops = h->prof_flops[k];
        ++n;
    }
    for (int k = 0; k < KC_COUNT; ++k) { h->prof_ms[k] = h->prof_bytes[k] = h->prof_flops[k] = 0; h->prof_n[k] = 0; }
    return n;
}

int texocr_set_sampling(texocr_handle* h, double temp, double threshold, uint64_t seed) {
    if (!h) return TEXOCR_ERR_ARG;
    if (temp > 0.0) {
        if (!(threshold >= 0.0 && threshold < 1.0)) return fail(h, TEXOCR_ERR_ARG, "sampling threshold must be in [0, 1)");
        if (h->cfg.vocab_size > 1024) return fail(h, TEXOCR_ERR_ARG, "sampling supports vocab_size <= 1024");
        const int k = (int)((1.0 - threshold) * (double)h->cfg.vocab_size);
        if (k < 1) return fail(h, TEXOCR_ERR_ARG, "top-k filter keeps k = %d logits: the reference's softmax would be all-NaN", k);
    }
    h->samp_temp = temp > 0.0 ? temp : 0.0; h->samp_threshold = threshold; h->samp_seed = seed; h->samp_calls = 0;
    drop_graphs(h);
    return 0;
}

int texocr_debug_sample_step(texocr_handle* h, const float* logits, int32_t rows, int32_t step, uint32_t call, int64_t* out_ids) {
    if (!h || !logits || !out_ids || rows <= 0 || step < 0) return TEXOCR_ERR_ARG;
    CK(cudaSetDevice(h->device));
    const int V = h->cfg.vocab_size;
    char* scratch = nullptr;
    const size_t nb = (size_t)rows * 8 * 2 + (size_t)rows * 4 + 64;
    CK(cudaMalloc(&scratch, nb));
    CK(cudaMemset(scratch, 0, nb));
    ArgmaxArgs aa{};
    aa.logits = logits; aa.B = rows; aa.V = V; aa.out_ids = (int64_t*)scratch; aa.out_ld = 1;
    aa.cur_tok = (int64_t*)(scratch + (size_t)rows * 8);
    int* ip = (int*)(scratch + (size_t)rows * 16);
    aa.step = ip; aa.done_step = ip + 1; aa.block_counter = ip + 2; aa.call_ctr = (unsigned*)(ip + 3); aa.seen_eos = ip + 8;
    aa.eos = -1;
    if (h->samp_temp > 0.0) { aa.topk = sampling_k(h); aa.inv_temp = (float)(1.0 / h->samp_temp); aa.seed = h->samp_seed; aa.row_base = 0; }
    const int64_t off = -(int64_t)step;        // the kernel writes out_ids[row * out_ld + step]
    aa.out_ids += off;
    CK(cudaMemcpy(ip, &step, 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ip + 3, &call, 4, cudaMemcpyHostToDevice));
    cudaError_t e = launch_argmax_step(aa, nullptr);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(out_ids, scratch, (size_t)rows * 8, is_device_ptr(out_ids) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost);
    cudaFree(scratch);
    if (e != cudaSuccess) return fail_cuda(h, e, "debug_sample_step", __LINE__, __FILE__);
    return 0;
}

int texocr_set_option(texocr_handle* h, const char* name, int64_t value) {
    if (!h || !name) return TEXOCR_ERR_ARG;
    if (!strcmp(name, "cuda_graph")) { h->use_graph = value != 0; return 0; }
    if (!strcmp(name, "stagger_us")) { h->stagger_us = (int)value; return 0; }
    if (!strcmp(name, "attn_seq")) { g_attn_seq = (int)value; drop_graphs(h); return 0; }
    if (!strcmp(name, "attn_abs_minb")) { g_attn_abs_minb = (int)value; drop_graphs(h); return 0; }
    if (!strcmp(name, "absorb_two_stage")) { h->absorb_two_stage = (int)value; drop_graphs(h); return 0; }
    if (!strcmp(name, "self_absorb")) { h->self_absorb = (int)value; drop_graphs(h); return 0; }
    if (!strcmp(name, "cross_absorb")) { h->cross_absorb = (int)value; drop_graphs(h); return 0; }
    if (!strcmp(name, "attn_trace")) { h->attn_trace_on = value != 0; drop_graphs(h); return 0; }
    if (!strcmp(name, "attn_full_tail")) { g_attn_full_tail = (int)value; drop_graphs(h); return 0; }
    if (!strcmp(name, "no_early_exit")) { h->no_early_exit = value != 0; return 0; }
    if (!strcmp(name, "keep_logits")) { h->keep_logits = value != 0; drop_graphs(h); return 0; }
    if (!strcmp(name, "poison")) { h->poison = value != 0; return 0; }
    if (!strcmp(name, "attn_ctas_per_sm")) { h->attn_ctas_per_sm = (int)std::max<int64_t>(1, std::min<int64_t>(8, value)); drop_graphs(h); return 0; }
    if (!strcmp(name, "dbg_skip")) { h->dbg_skip = (int)value; drop_graphs(h); return 0; }
    if (!strcmp(name, "pdl_mid")) { g_texocr_pdl_mid = (int)value; drop_graphs(h); return 0; }
    if (!strcmp(name, "pdl")) {
        g_texocr_pdl = (int)value;
        drop_graphs(h);
        return 0;
    }
    if (!strcmp(name, "tma_attention")) {
        h->use_tma_attn = (int)value;
        drop_graphs(h);
        return 0;
    }
    if (!strcmp(name, "decode_branches")) {
        if (value < 0 || value > MAX_BRANCH) return fail(h, TEXOCR_ERR_ARG, "decode_branches must be in [0, %d] (0 = automatic)", MAX_BRANCH);
        h->decode_branches = (int)value;
        return 0;
    }
    if (!strcmp(name, "gemm_persistent")) { g_tc_persistent = (int)value; return 0; }
    if (!strcmp(name, "gemm_min_ctas")) { g_tc_min_ctas = (int)value; drop_graphs(h); return 0; }
    if (!strcmp(name, "gemm_persistent_stages")) { g_tc_persistent_stages = (int)value; return 0; }
    if (!strcmp(name, "gemm_bn256")) { g_tc_bn256 = (int)value; return 0; }
    if (!strcmp(name, "gemm_epi_warps")) { g_tc_epi_warps = value == 4 ? 4 : 8; drop_graphs(h); return 0; }
    if (!strcmp(name, "gn_fused")) { h->gn_fused = (int)value; return 0; }
    if (!strcmp(name, "im2col_tma")) { h->use_im2col_tma = value != 0; return 0; }
    if (!strcmp(name, "conv_gather")) { h->use_conv_gather = value != 0; return 0; }
    if (!strcmp(name, "stem_tc")) { h->use_stem_tc = value != 0; return 0; }
    if (!strcmp(name, "tcgen05")) {
        h->use_tcgen05 = value != 0;
        drop_graphs(h);
        return 0;
    }
    return fail(h, TEXOCR_ERR_ARG, "unknown option '%s'", name);
}

int64_t texocr_debug_read(texocr_handle* h, const char* name, float* out, int64_t cap_elems) {
    if (!h || !name || !out) return TEXOCR_ERR_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    if (!strcmp(name, "backbone")) {
        const int64_t n = (int64_t)h->last_backbone_pixels * 1024;
        if (n <= 0) return fail(h, TEXOCR_ERR_STATE, "no backbone activation recorded");
        if (n > cap_elems) return fail(h, TEXOCR_ERR_ARG, "buffer too small: need %lld floats", (long long)n);
        // the last block wrote into whichever ping-pong buffer is current: 12 blocks -> pingpong[1] (actB)
        const float* src = h->actB.as<float>();
        CK(cudaMemcpy(out, src, (size_t)n * 4, is_device_ptr(out) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost));
        return n;
    }
    {   // raw workspace taps (byte-exact copies reinterpreted as float32 words): name -> buffer
        struct { const char* n; DevBuf* b; } taps[] = {{"logits", &h->logits}, {"kvcache", &h->kvcache}, {"x", &h->x}, {"s", &h->s},
                                                       {"xn", &h->xn}, {"qkv", &h->qkv}, {"o", &h->o}, {"hid", &h->hid},
                                                       {"crosskv_hm", &h->crosskv_hm}, {"enc_out", &h->enc_out}, {"attn_trace", &h->attn_trace}, {"latcache", &h->latcache}};
        for (auto& t : taps)
            if (!strcmp(name, t.n)) {
                if (!t.b->p) return fail(h, TEXOCR_ERR_STATE, "buffer '%s' not allocated", name);
                const int64_t n = std::min<int64_t>((int64_t)(t.b->bytes / 4), cap_elems);
                CK(cudaMemcpy(out, t.b->p, (size_t)n * 4, is_device_ptr(out) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost));
                return n;
            }
    }
    return fail(h, TEXOCR_ERR_ARG, "unknown debug tap '%s'", name);
}

int texocr_debug_gemm(texocr_handle* h, const void* A, const void* W, void* C, int32_t M, int32_t N, int32_t K, int32_t lda,
                      int32_t ldw, int32_t ldc, int32_t epi, int32_t dt_a, int32_t dt_c, const float* bias, const float* res,
                      int32_t ldres, int32_t use_tc, const void* A2, const void* W2, void* stream) {
    if (!h) return TEXOCR_ERR_ARG;
    CK(cudaSetDevice(h->device));
    StreamHop hop__(h, stream);
    cudaStream_t st = hop__.work;
    GemmArgs g = mk_gemm(A, lda, W, ldw, C, ldc, M, N, K, epi, dt_a, dt_c, bias, res, ldres);
    g.A2 = A2; g.W2 = W2;
    if (use_tc) {
        if (!tc_gemm_supported(g)) return fail(h, TEXOCR_ERR_ARG, "shape not supported by the tcgen05 GEMM");
        LAUNCH(KC_MISC, 1, gemm_bytes(g, 2), gemm_flops(g), launch_gemm_tc(g, st));
    } else {
        LAUNCH(KC_MISC, 1, gemm_bytes(g, dt_a == DT_BF16 ? 2 : 4), gemm_flops(g), launch_gemm_simt(g, st));
    }
    return 0;
}

int texocr_debug_attn_decode(texocr_handle* h, int32_t self, const void* q, int32_t ldq, const void* knew, const void* vnew,
                             int32_t ldnew, void* kv, int64_t kv_rows, int32_t ldkv, int32_t col0, int32_t tcap,
                             const int32_t* k_off_dev, const int32_t* step_dev, void* out, int32_t batch, int32_t max_keys,
                             int32_t use_tma, void* stream) {
    if (!h) return TEXOCR_ERR_ARG;
    (void)ldkv; (void)col0;
    CK(cudaSetDevice(h->device));
    StreamHop hop__(h, stream);
    cudaStream_t st = hop__.work;
    AttnDecodeArgs a{};
    char* base = (char*)kv;
    a.q = q; a.ldq = ldq; a.o = out; a.ldo = 512; a.batch = batch; a.dt = DT_BF16; a.ldkv = 128;
    a.kcache = base; a.vcache = base + 64 * 2;
    KvLayout lay{kv, (long)kv_rows, 128, 128, 0, 0, 64, 0, 0};
    if (self) {     // kv: [batch][8][tcap][128]
        a.knew = knew; a.vnew = vnew; a.ldnew = ldnew; a.batch_stride = (int64_t)tcap * 1024; a.head_stride = (int64_t)tcap * 128;
        a.step = step_dev; lay.row_h = tcap; lay.row_b = 8 * tcap;
    } else {        // kv: [8][ntok][128], kv_rows = 8 * ntok
        a.k_off = k_off_dev; a.head_stride = (kv_rows / 8) * 128; lay.row_h = (int)(kv_rows / 8);
    }
    if (use_tma) {
        if (!attn_decode_tma_supported(a)) return fail(h, TEXOCR_ERR_ARG, "not supported by the TMA attention kernel");
        LAUNCH(KC_MISC, 1, 0.0, 0.0, launch_attn_decode_tma(a, lay, h->num_sms * h->attn_ctas_per_sm, st));
    } else {
        LAUNCH(KC_MISC, 1, 0.0, 0.0, launch_attn_decode(a, max_keys, st));
    }
    return 0;
}


int texocr_debug_attn_abs(texocr_handle* h, const void* q, void* latent, int64_t latent_rows, const int32_t* k_off_dev, const void* znew,
                          int32_t tcap, const int32_t* step_dev, void* out, int32_t batch, void* stream) {
    if (!h || !q || !latent || !out || batch <= 0 || latent_rows <= 0 || (!k_off_dev && !znew)) return TEXOCR_ERR_ARG;
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    AttnAbsArgs ab{};
    ab.q = q; ab.ldq = 2048; ab.latent = latent; ab.latent_rows = (long)latent_rows; ab.k_off = k_off_dev; ab.o = out; ab.ldo = 2048; ab.batch = batch;
    if (znew) { ab.znew = znew; ab.ldz = 256; ab.tcap = tcap; ab.step = step_dev; }
    CK(g_attn_seq ? launch_attn_seq(ab, h->num_sms * h->attn_ctas_per_sm, st) : launch_attn_abs(ab, h->num_sms * h->attn_ctas_per_sm, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

int texocr_debug_fold_absorbed(const float* wq, const float* wk, const float* wv, const float* wo, float* wqk_out, float* wvo_out) {
    if (!wq || !wk || !wv || !wo || !wqk_out || !wvo_out) return TEXOCR_ERR_ARG;
    HostTensor q, k, v, o;
    q.shape = k.shape = v.shape = {512, 256}; o.shape = {512, 512};
    q.data.assign(wq, wq + 512 * 256); k.data.assign(wk, wk + 512 * 256); v.data.assign(wv, wv + 512 * 256); o.data.assign(wo, wo + 512 * 512);
    std::vector<float> wqk, wvoi;
    fold_absorbed(q, k, v, o, wqk, wvoi);
    memcpy(wqk_out, wqk.data(), wqk.size() * sizeof(float));
    memcpy(wvo_out, wvoi.data(), wvoi.size() * sizeof(float));
    return 0;
}
}  // extern "C"
