// Weight ingestion: the reference's state_dict by name -> packed device weights (standardised NHWC convolution
// filters, concatenated q/k/v, interleaved GLU / GeGLU rows, absorbed-attention folds, bf16 / split-bf16 copies).
#include "engine_internal.h"

// ------------------------------------------------------------------------------------------------ weights
template <typename T> static int dev_upload(texocr_handle* h, const std::vector<T>& v, void** out) {
    void* p = nullptr;
    CK(cudaMalloc(&p, std::max(v.size() * sizeof(T), (size_t)16)));
    CK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    h->weight_allocs.push_back(p);
    *out = p;
    return 0;
}
static uint16_t f2bf(float f) {     // round-to-nearest-even, like __float2bfloat16_rn
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static int upload_f32(texocr_handle* h, const std::vector<float>& v, float** out) { return dev_upload<float>(h, v, (void**)out); }
static int upload_act(texocr_handle* h, const std::vector<float>& v, void** out) {      // in the GEMM operand type
    if (h->dt == DT_F32) return dev_upload<float>(h, v, out);
    std::vector<uint16_t> b(v.size());
    for (size_t i = 0; i < v.size(); ++i) b[i] = f2bf(v[i]);
    return dev_upload<uint16_t>(h, b, out);
}

// v -> (bf16(v), bf16(v - bf16(v))): the operand pair of the tcgen05 bf16x3 mode
static float bf2f(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }
static int upload_split(texocr_handle* h, const std::vector<float>& v, void** hi, void** lo) {
    std::vector<uint16_t> a(v.size()), b(v.size());
    for (size_t i = 0; i < v.size(); ++i) { a[i] = f2bf(v[i]); b[i] = f2bf(v[i] - bf2f(a[i])); }
    int r = dev_upload<uint16_t>(h, a, hi);
    if (r) return r;
    return dev_upload<uint16_t>(h, b, lo);
}

static const HostTensor* find_w(texocr_handle* h, const std::string& key, std::initializer_list<int64_t> shape) {
    auto it = h->sd.find(key);
    if (it == h->sd.end()) { fail(h, TEXOCR_ERR_WEIGHT, "missing state_dict entry '%s'", key.c_str()); return nullptr; }
    if (it->second.shape != std::vector<int64_t>(shape)) {
        fail(h, TEXOCR_ERR_WEIGHT, "state_dict entry '%s' has the wrong shape", key.c_str());
        return nullptr;
    }
    return &it->second;
}
#define GETW(var, key, ...) const HostTensor* var = find_w(h, (key), {__VA_ARGS__}); if (!var) return TEXOCR_ERR_WEIGHT

// rows interleaved for the GLU / GeGLU epilogues: packed row 2j = W[j] (value), 2j+1 = W[half + j] (gate)
static void interleave_rows(const std::vector<float>& w, int rows, int cols, std::vector<float>& out) {
    const int half = rows / 2;
    out.resize(w.size());
    for (int j = 0; j < half; ++j) {
        memcpy(&out[(size_t)(2 * j) * cols], &w[(size_t)j * cols], cols * sizeof(float));
        memcpy(&out[(size_t)(2 * j + 1) * cols], &w[(size_t)(half + j) * cols], cols * sizeof(float));
    }
}

// Absorbed K / V projections (attn_abs_kernel), per head h:
//   Wqk[h*256 + c][i] = sum_d Wk[h*64+d][c] * Wq[h*64+d][i]      Q'_h = xn . Wqk_h^T = (xn Wq_h^T) Wk_h
//   Wvo[o][h*256 + c] = sum_d Wo[o][h*64+d] * Wv[h*64+d][c]      y = sum_h C_h . Wvo_h^T = sum_h (C_h Wv_h^T) Wo_h^T
// (fp64 accumulation; Wvo rows interleaved like Wo for the GLU epilogue)
// Replica handles of one model (texocr_b200/pipeline.py) fold the same matrices: the result is cached per process, keyed by a
// 64-bit FNV-1a hash of the four source matrices.
static std::map<uint64_t, std::pair<std::vector<float>, std::vector<float>>> g_fold_cache;
static std::mutex g_fold_mu;
static uint64_t fnv1a(const std::vector<float>& v, uint64_t hsh) {
    const unsigned char* p = reinterpret_cast<const unsigned char*>(v.data());
    const size_t n = v.size() * sizeof(float);
    for (size_t i = 0; i < n; ++i) { hsh ^= p[i]; hsh *= 1099511628211ull; }
    return hsh;
}
static void fold_absorbed_compute(const HostTensor& q, const HostTensor& k, const HostTensor& v, const HostTensor& wo, std::vector<float>& wqk,
                                  std::vector<float>& wvoi);
void fold_absorbed(const HostTensor& q, const HostTensor& k, const HostTensor& v, const HostTensor& wo, std::vector<float>& wqk,
                   std::vector<float>& wvoi) {
    const uint64_t key = fnv1a(wo.data, fnv1a(v.data, fnv1a(k.data, fnv1a(q.data, 14695981039346656037ull))));
    {
        std::lock_guard<std::mutex> lk(g_fold_mu);
        auto it = g_fold_cache.find(key);
        if (it != g_fold_cache.end()) { wqk = it->second.first; wvoi = it->second.second; return; }
    }
    fold_absorbed_compute(q, k, v, wo, wqk, wvoi);
    std::lock_guard<std::mutex> lk(g_fold_mu);
    if (g_fold_cache.size() >= 64) g_fold_cache.clear();
    g_fold_cache[key] = std::make_pair(wqk, wvoi);
}
static void fold_absorbed_compute(const HostTensor& q0, const HostTensor& k0, const HostTensor& v0, const HostTensor& wo0, std::vector<float>& wqk,
                                  std::vector<float>& wvoi) {
    // The bf16 tier's weights ARE their bf16 roundings (what the projected path multiplies with, and what a 'mixed' weight blob
    // stores): fold those, so that a model built from fp32 weights and one built from the blob stay bit-identical.
    HostTensor q = q0, k = k0, v = v0, wo = wo0;
    for (HostTensor* t : {&q, &k, &v, &wo})
        for (float& x : t->data) x = bf2f(f2bf(x));
    wqk.assign((size_t)2048 * 256, 0.f);
    std::vector<float> wvo((size_t)512 * 2048);
    std::vector<double> acc(256);
    for (int hh = 0; hh < 8; ++hh)
        for (int cc = 0; cc < 256; ++cc) {
            std::fill(acc.begin(), acc.end(), 0.0);
            for (int d = 0; d < 64; ++d) {
                const double kv = k.data[(size_t)(hh * 64 + d) * 256 + cc];
                const float* qr = &q.data[(size_t)(hh * 64 + d) * 256];
                for (int i = 0; i < 256; ++i) acc[i] += kv * qr[i];
            }
            for (int i = 0; i < 256; ++i) wqk[(size_t)(hh * 256 + cc) * 256 + i] = (float)acc[i];
        }
    for (int o2 = 0; o2 < 512; ++o2)
        for (int hh = 0; hh < 8; ++hh) {
            std::fill(acc.begin(), acc.end(), 0.0);
            for (int d = 0; d < 64; ++d) {
                const double ov = wo.data[(size_t)o2 * 512 + hh * 64 + d];
                const float* vr = &v.data[(size_t)(hh * 64 + d) * 256];
                for (int cc = 0; cc < 256; ++cc) acc[cc] += ov * vr[cc];
            }
            for (int cc = 0; cc < 256; ++cc) wvo[(size_t)o2 * 2048 + hh * 256 + cc] = (float)acc[cc];
        }
    interleave_rows(wvo, 512, 2048, wvoi);
}

static int pack_attn(texocr_handle* h, const std::string& p, bool cross, AttnW& out, bool decoder = false) {
    GETW(q, p + ".q.weight", 512, 256);
    GETW(k, p + ".k.weight", 512, 256);
    GETW(v, p + ".v.weight", 512, 256);
    GETW(wo, p + ".fc_out.0.weight", 512, 512);
    GETW(bo, p + ".fc_out.0.bias", 512);
    int r;
    if (!cross) {
        std::vector<float> qkv;
        qkv.insert(qkv.end(), q->data.begin(), q->data.end());
        qkv.insert(qkv.end(), k->data.begin(), k->data.end());
        qkv.insert(qkv.end(), v->data.begin(), v->data.end());
        if ((r = upload_act(h, qkv, &out.wqkv))) return r;
    } else {
        if ((r = upload_act(h, q->data, &out.wq))) return r;
    }
    if (decoder && h->dt != DT_F32) {
        std::vector<float> wqk, wvoi;
        fold_absorbed(*q, *k, *v, *wo, wqk, wvoi);
        if ((r = upload_act(h, wqk, &out.wqk))) return r;
        if ((r = upload_act(h, wvoi, &out.wvo))) return r;
        if ((r = upload_act(h, v->data, &out.wv))) return r;      // [512 = head*64 + d, 256]: block-diagonal value projection of C
    }
    std::vector<float> woi, boi;
    interleave_rows(wo->data, 512, 512, woi);
    interleave_rows(bo->data, 512, 1, boi);
    if ((r = upload_act(h, woi, &out.wo))) return r;
    return upload_f32(h, boi, &out.bo);
}
static int pack_mlp(texocr_handle* h, const std::string& p, MlpW& out) {
    GETW(w1, p + ".fc_in.fc.weight", 2048, 256);
    GETW(b1, p + ".fc_in.fc.bias", 2048);
    GETW(w2, p + ".fc_out.weight", 256, 1024);
    GETW(b2, p + ".fc_out.bias", 256);
    std::vector<float> w1i, b1i;
    interleave_rows(w1->data, 2048, 256, w1i);
    interleave_rows(b1->data, 2048, 1, b1i);
    int r;
    if ((r = upload_act(h, w1i, &out.w1))) return r;
    if ((r = upload_f32(h, b1i, &out.b1))) return r;
    if ((r = upload_act(h, w2->data, &out.w2))) return r;
    return upload_f32(h, b2->data, &out.b2);
}

// model/resnet.py:61-64: w_hat = (w - mean) / sqrt(biased var + 1e-6) per output channel; folded once here
// (double accumulation), reordered [cout][cin][ky][kx] -> [cout][ky][kx][cin] for the NHWC implicit GEMM.
static void standardise_reorder(const HostTensor& w, int cout, int cin, int k, std::vector<float>& out) {
    const int n = cin * k * k;
    out.resize((size_t)cout * n);
    for (int o = 0; o < cout; ++o) {
        const float* src = &w.data[(size_t)o * n];
        double s = 0.0, q = 0.0;
        for (int i = 0; i < n; ++i) s += src[i];
        const double mean = s / n;
        for (int i = 0; i < n; ++i) { const double d = src[i] - mean; q += d * d; }
        const double rstd = 1.0 / sqrt(q / n + 1e-6);
        for (int c = 0; c < cin; ++c)
            for (int t = 0; t < k * k; ++t)
                out[(size_t)o * n + (size_t)t * cin + c] = (float)((src[(size_t)c * k * k + t] - mean) * rstd);
    }
}

int finalize_weights(texocr_handle* h) {
    const texocr_config& c = h->cfg;
    int r;
    const std::string E = "encoder.";
    if (c.encoder_kind == TEXOCR_ENC_HYBRID) {
        const std::string bb = E + "patch_embed.backbone_net.";
        {   // stem: [64][1][7][7] -> standardised [tap][oc]
            GETW(w, bb + "stem.0.weight", 64, 1, 7, 7);
            std::vector<float> ws, wt(49 * 64);
            standardise_reorder(*w, 64, 1, 7, ws);
            for (int o = 0; o < 64; ++o) for (int t = 0; t < 49; ++t) wt[t * 64 + o] = ws[o * 49 + t];
            if ((r = upload_f32(h, wt, &h->stem_w))) return r;
            if (h->dt == DT_BF16) {      // tensor-core stem: [64 out][64 k] with k = tap for k < 49, zero otherwise
                std::vector<float> wk(64 * 64, 0.f);
                for (int o = 0; o < 64; ++o) for (int t = 0; t < 49; ++t) wk[o * 64 + t] = ws[o * 49 + t];
                if ((r = upload_split(h, wk, &h->stem_w_hi, &h->stem_w_lo))) return r;
            }
            GETW(g, bb + "stem.1.weight", 64);
            GETW(b, bb + "stem.1.bias", 64);
            if ((r = upload_f32(h, g->data, &h->stem_g))) return r;
            if ((r = upload_f32(h, b->data, &h->stem_b))) return r;
        }
        const int depths[3] = {2, 4, 6}, chans[3] = {256, 512, 1024};
        int prev = 64;
        for (int s = 0; s < 3; ++s) {
            const int cout = chans[s], mid = cout / 4;
            for (int b = 0; b < depths[s]; ++b) {
                const int stride = (b == 0) ? (s == 0 ? 1 : 2) : 1;
                const std::string p = bb + "stages." + std::to_string(s) + ".stage_blocks." + std::to_string(b);
                struct Spec { std::string name, gn; int cin, cout, k, stride, act; };
                std::vector<Spec> specs;
                if (b == 0) specs.push_back({p + ".downsample.conv", p + ".downsample.norm", prev, cout, 1, stride, 0});
                specs.push_back({p + ".block_list.0", p + ".block_list.1", prev, mid, 1, 1, 1});
                specs.push_back({p + ".block_list.2", p + ".block_list.3", mid, mid, 3, stride, 1});
                specs.push_back({p + ".block_list.4", p + ".block_list.5", mid, cout, 1, 1, 0});
                for (auto& sp : specs) {
                    GETW(w, sp.name + ".weight", sp.cout, sp.cin, sp.k, sp.k);
                    GETW(g, sp.gn + ".weight", sp.cout);
                    GETW(be, sp.gn + ".bias", sp.cout);
                    ConvW cw;
                    cw.name = sp.name; cw.gn = sp.gn; cw.cin = sp.cin; cw.cout = sp.cout; cw.k = sp.k; cw.stride = sp.stride; cw.act = sp.act;
                    std::vector<float> ws;
                    standardise_reorder(*w, sp.cout, sp.cin, sp.k, ws);
                    if ((r = upload_f32(h, ws, &cw.w))) return r;
                    if (h->dt == DT_BF16 && (r = upload_split(h, ws, &cw.w_hi, &cw.w_lo))) return r;
                    if ((r = upload_f32(h, g->data, &cw.gamma))) return r;
                    if ((r = upload_f32(h, be->data, &cw.beta))) return r;
                    h->convs.push_back(cw);
                }
                prev = cout;
            }
        }
        GETW(pw, E + "patch_embed.proj.weight", 256, 1024, 1, 1);
        GETW(pb, E + "patch_embed.proj.bias", 256);
        if (h->dt == DT_BF16) { if ((r = upload_split(h, pw->data, &h->proj_w, &h->proj_w_lo))) return r; }
        else if ((r = upload_act(h, pw->data, &h->proj_w))) return r;
        if ((r = upload_f32(h, pb->data, &h->proj_b))) return r;
        h->proj_k = 1024;
    } else {
        GETW(pw, E + "patch_embed.proj.weight", 256, 1, 16, 16);
        GETW(pb, E + "patch_embed.proj.bias", 256);
        if ((r = upload_act(h, pw->data, &h->proj_w))) return r;      // [256][ky*16+kx] already K-major
        if ((r = upload_f32(h, pb->data, &h->proj_b))) return r;
        h->proj_k = 256;
    }
    const int64_t npos = (c.encoder_kind == TEXOCR_ENC_HYBRID ? 10 : 63) * 63 + 1;
    GETW(cls, E + "cls_token", 1, 1, 256);
    GETW(pos, E + "pos_embed", 1, npos, 256);
    if ((r = upload_f32(h, cls->data, &h->cls))) return r;
    if ((r = upload_f32(h, pos->data, &h->pos))) return r;
    {
        GETW(g, E + "attn_layers.layers.0.0.weight", 256);
        GETW(b, E + "attn_layers.layers.0.0.bias", 256);
        GETW(ng, E + "norm.weight", 256);
        GETW(nb, E + "norm.bias", 256);
        if ((r = upload_f32(h, g->data, &h->enc_ln_g))) return r;
        if ((r = upload_f32(h, b->data, &h->enc_ln_b))) return r;
        if ((r = upload_f32(h, ng->data, &h->enc_norm_g))) return r;
        if ((r = upload_f32(h, nb->data, &h->enc_norm_b))) return r;
    }
    h->enc_attn.resize(c.enc_layers); h->enc_mlp.resize(c.enc_layers);
    for (int l = 0; l < c.enc_layers; ++l) {
        if ((r = pack_attn(h, E + "attn_layers.layers." + std::to_string(2 * l) + ".1", false, h->enc_attn[l]))) return r;
        if ((r = pack_mlp(h, E + "attn_layers.layers." + std::to_string(2 * l + 1) + ".1", h->enc_mlp[l]))) return r;
    }
    const std::string Dn = "decoder.net.";
    GETW(te, Dn + "token_embedding.weight", c.vocab_size, 256);
    GETW(pe, Dn + "pos_embedding.embedding.weight", c.max_length, 256);
    if ((r = upload_f32(h, te->data, &h->tok_emb))) return r;
    if ((r = upload_f32(h, pe->data, &h->pos_emb))) return r;
    {
        GETW(g, Dn + "attn_layers.layers.0.0.weight", 256);
        GETW(b, Dn + "attn_layers.layers.0.0.bias", 256);
        GETW(ng, Dn + "norm.weight", 256);
        GETW(nb, Dn + "norm.bias", 256);
        if ((r = upload_f32(h, g->data, &h->dec_ln_g))) return r;
        if ((r = upload_f32(h, b->data, &h->dec_ln_b))) return r;
        if ((r = upload_f32(h, ng->data, &h->dec_norm_g))) return r;
        if ((r = upload_f32(h, nb->data, &h->dec_norm_b))) return r;
    }
    h->dec_self.resize(c.dec_layers); h->dec_cross.resize(c.dec_layers); h->dec_mlp.resize(c.dec_layers);
    std::vector<float> ckv;
    for (int l = 0; l < c.dec_layers; ++l) {
        const std::string base = Dn + "attn_layers.layers.";
        if ((r = pack_attn(h, base + std::to_string(3 * l) + ".1", false, h->dec_self[l], true))) return r;
        if ((r = pack_attn(h, base + std::to_string(3 * l + 1) + ".1", true, h->dec_cross[l], true))) return r;
        if ((r = pack_mlp(h, base + std::to_string(3 * l + 2) + ".1", h->dec_mlp[l]))) return r;
        GETW(k, base + std::to_string(3 * l + 1) + ".1.k.weight", 512, 256);
        GETW(v, base + std::to_string(3 * l + 1) + ".1.v.weight", 512, 256);
        ckv.insert(ckv.end(), k->data.begin(), k->data.end());
        ckv.insert(ckv.end(), v->data.begin(), v->data.end());
    }
    if ((r = upload_act(h, ckv, &h->w_crosskv))) return r;
    GETW(lw, Dn + "to_logits.weight", c.vocab_size, 256);
    GETW(lb, Dn + "to_logits.bias", c.vocab_size);
    if ((r = upload_act(h, lw->data, &h->w_logits))) return r;
    if ((r = upload_f32(h, lb->data, &h->b_logits))) return r;
    h->sd.clear();
    h->finalized = true;
    return 0;
}
