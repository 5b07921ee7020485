// Host-side engine + C-ABI of libtexocr_b200.so (include/texocr.h).
//
// What runs where (reference file:line in brackets):
//   texocr_encode            ResNetV2 stem + bottlenecks [model/resnet.py:141-149,251-254] as NHWC implicit GEMMs with
//                            GroupNorm statistics/apply kernels, 1x1 projection, cls/pos assembly
//                            [model/encoder.py:128-143], ViT blocks with the shared double LayerNorm
//                            [model/attention.py:237-259], final norm [model/encoder.py:148].
//   texocr_decoder_logits    teacher-forced Transformer.forward [model/decoder.py:41-67].
//   texocr_decoder_generate  KV-cached greedy loop, one CUDA graph per decode step, on-device EOS bookkeeping
//                            [model/decoder.py:77-122]; cross-attention K/V of the memory projected once.
// There is no CPU fallback anywhere: every entry point needs the sm_100 device the handle was created on.
#include "engine.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <mutex>

#include "decode_mega.h"
extern int g_tc_split_k;
extern int g_tc_persistent;
extern int g_tc_persistent_stages;
extern int g_tc_persist_min_tiles;
extern int g_tc_tiles_per_cta;
extern int g_tc_min_ctas;
extern int g_tc_deep_ring;
extern int g_tc_shallow_ring;
#include "tc_gemm.h"

static std::string g_create_error;
// Several handles may be driven from different host threads (texocr_b200/pipeline.py).  Stream capture and device-wide
// operations do not mix across threads (a cudaDeviceSynchronize / cudaFree in one thread invalidates a capture in
// another), so graph capture and (re)allocation take this process-wide lock.  Steady-state calls never hold it.
static std::recursive_mutex g_dev_mu;
// bits 0..5: programmatic dependent launch per kernel family (kernels.h); bit 8 / 9: LayerNorm / GEMM kernels release their
// dependents only after their stores (experiment switches; the default is an early trigger everywhere).
int g_texocr_pdl = 0x3f;
extern int g_attn_full_tail;
extern int g_attn_abs_minb;
extern int g_attn_l2_policy;

#define CK(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess) return fail_cuda(h, e__, #expr, __LINE__);                         \
    } while (0)

static int fail(texocr_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}
static int fail_cuda(texocr_handle* h, cudaError_t e, const char* what, int line) {
    return fail(h, TEXOCR_ERR_CUDA, "CUDA error %s (%s) at engine.cu:%d: %s", cudaGetErrorName(e), cudaGetErrorString(e), line, what);
}

// ------------------------------------------------------------------------------------------------ profiling / launch accounting
static const char* kclass_name[KC_COUNT] = {
    "stem_conv", "gn_stats", "gn_apply", "conv_gemm", "enc_gemm", "enc_attn", "enc_rowwise", "crosskv_gemm",
    "dec_gemm", "dec_attn_self", "dec_attn_cross", "dec_rowwise", "dec_argmax", "tf_gemm", "tf_attn", "tf_rowwise", "misc", "dec_mega",
    "dec_gemm_q", "dec_gemm_vproj", "dec_gemm_wo", "dec_gemm_w1", "dec_gemm_w2", "dec_gemm_logits"};

static cudaEvent_t get_event(texocr_handle* h) {
    if (!h->ev_pool.empty()) { cudaEvent_t e = h->ev_pool.back(); h->ev_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

#define LAUNCH(kc_, nkern, bytes_, flops_, expr)                                                   \
    do {                                                                                           \
        ProfRec pr__;                                                                              \
        if (h->prof_on) { pr__.cls = (kc_); pr__.bytes = (bytes_); pr__.flops = (flops_);          \
            pr__.e0 = get_event(h); pr__.e1 = get_event(h); cudaEventRecord(pr__.e0, st); }        \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess) return fail_cuda(h, e__, #expr, __LINE__);                         \
        h->launches += (nkern);                                                                    \
        if (h->prof_on) { cudaEventRecord(pr__.e1, st); h->prof.push_back(pr__); }                 \
    } while (0)

// ------------------------------------------------------------------------------------------------ memory helpers
static void drop_graphs(texocr_handle* h) {
    for (int i = 0; i < 16; ++i) {
        if (h->bgraph_exec[i]) { cudaGraphExecDestroy(h->bgraph_exec[i]); h->bgraph_exec[i] = nullptr; }
        if (h->bgraph[i]) { cudaGraphDestroy(h->bgraph[i]); h->bgraph[i] = nullptr; }
    }
    h->graph_exec = nullptr; h->graph = nullptr;
    for (int i = 0; i < 2; ++i) {
        if (h->cgraph_exec[i]) { cudaGraphExecDestroy(h->cgraph_exec[i]); h->cgraph_exec[i] = nullptr; }
        if (h->cgraph[i]) { cudaGraphDestroy(h->cgraph[i]); h->cgraph[i] = nullptr; }
    }
}

static int ensure(texocr_handle* h, DevBuf& b, size_t bytes) {
    if (b.bytes >= bytes && b.p) return 0;
    std::lock_guard<std::recursive_mutex> lk(g_dev_mu);
    if (b.p) { CK(cudaDeviceSynchronize()); CK(cudaFree(b.p)); b.p = nullptr; b.bytes = 0; }
    size_t want = std::max(bytes, (size_t)256);
    want = (want + 255) & ~(size_t)255;
    CK(cudaMalloc(&b.p, want));
    CK(cudaMemset(b.p, 0, want));        // fresh workspaces start zeroed (no NaN bit patterns in never-written KV rows)
    CK(cudaDeviceSynchronize());         // the memset runs on the legacy stream; our streams are non-blocking
    b.bytes = want;
    drop_graphs(h);      // pointers may have moved
    return 0;
}
#define ENSURE(buf, bytes) do { int r__ = ensure(h, (buf), (bytes)); if (r__) return r__; } while (0)

static bool is_device_ptr(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// Return a device pointer for `p` (bytes long): p itself if it is device memory, else a staged copy.
static int to_device(texocr_handle* h, const void* p, size_t bytes, DevBuf& stage, const void** out, cudaStream_t st) {
    if (is_device_ptr(p)) { *out = p; return 0; }
    ENSURE(stage, bytes);
    CK(cudaMemcpyAsync(stage.p, p, bytes, cudaMemcpyHostToDevice, st));
    *out = stage.p;
    return 0;
}
static int from_device(texocr_handle* h, void* dst, const void* src, size_t bytes, cudaStream_t st) {
    if (dst == src) return 0;
    CK(cudaMemcpyAsync(dst, src, bytes, is_device_ptr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    return 0;
}

static int upload_ints(texocr_handle* h, const std::vector<int>& v, cudaStream_t st) {
    const size_t bytes = v.size() * sizeof(int);
    if (h->h_geom_cap < bytes) {
        if (h->h_geom) { CK(cudaEventSynchronize(h->geom_ev)); CK(cudaFreeHost(h->h_geom)); }
        h->h_geom_cap = std::max(bytes * 2, (size_t)4096);
        CK(cudaMallocHost(&h->h_geom, h->h_geom_cap));
    }
    CK(cudaEventSynchronize(h->geom_ev));          // the previous upload has left the staging buffer
    memcpy(h->h_geom, v.data(), bytes);
    ENSURE(h->geom, bytes);
    CK(cudaMemcpyAsync(h->geom.p, h->h_geom, bytes, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(h->geom_ev, st));
    return 0;
}

// Debug aid: fill every workspace with 0xFF bytes (NaN patterns) so that any read of memory the current call did not write
// shows up in the results.  Enabled with texocr_set_option(h, "poison", 1).
static int poison_workspaces(texocr_handle* h, cudaStream_t st) {
    DevBuf* bufs[] = {&h->raw1, &h->act2, &h->actA, &h->actB, &h->rawMid, &h->actMid, &h->rawMid2, &h->actMid2, &h->raw3, &h->rawDs,
                      &h->gn_partial, &h->gn_stats[0], &h->gn_stats[1], &h->gn_stats[2], &h->gn_stats[3], &h->proj_out, &h->patch_cols,
                      &h->backbone_a, &h->col, &h->x, &h->s, &h->xn, &h->qkv, &h->o, &h->hid, &h->logits, &h->enc_out, &h->enc_a,
                      &h->crosskv, &h->crosskv_hm, &h->kvcache, &h->out_ids};
    for (DevBuf* b : bufs) if (b->p) CK(cudaMemsetAsync(b->p, 0xFF, b->bytes, st));
    return 0;
}

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
static void fold_absorbed(const HostTensor& q, const HostTensor& k, const HostTensor& v, const HostTensor& wo, std::vector<float>& wqk,
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

static int finalize_weights(texocr_handle* h) {
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

// ------------------------------------------------------------------------------------------------ GEMM dispatch
// tcgen05 path for bf16 operands when the shape fits its tiles, FFMA path otherwise (and always in the fp32 tier).
static cudaError_t run_gemm(texocr_handle* h, const GemmArgs& g, cudaStream_t st) {
    if ((h->dbg_skip & 8) && g.M <= 512) return cudaSuccess;
    if (h->use_tcgen05 && g.dt_a == DT_BF16 && !g.conv && tc_gemm_supported(g)) {
        if (h->attn_trace_on && h->attn_trace.p && g.M <= 4096) {      // decode-sized GEMM: per-CTA residency sums next to the attention timers
            GemmArgs gd = g;
            gd.dbg = h->attn_trace.as<unsigned long long>() + (size_t)16 * 3 * 2048 + 16;
            return launch_gemm_tc(gd, st);
        }
        return launch_gemm_tc(g, st);
    }
    if (g.a_block_k) return cudaErrorInvalidValue;      // block-diagonal mode exists in the tcgen05 kernel only
    return launch_gemm_simt(g, st);
}
static GemmArgs mk_gemm(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K, int epi,
                        int dt_a, int dt_c, const float* bias, const float* res, int ldres) {
    GemmArgs g{};
    g.A = A; g.W = W; g.C = C; g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldw = ldw; g.ldc = ldc;
    g.bias = bias; g.res = res; g.ldres = ldres; g.dt_a = dt_a; g.dt_c = dt_c; g.epi = epi; g.conv = nullptr;
    return g;
}
// algorithmic bytes of a GEMM launch: both operands once + what the epilogue reads / writes
static double gemm_bytes(const GemmArgs& g, size_t esz) {
    double out = 0.0;
    switch (g.epi) {
        case EPI_STORE: out = (double)g.M * g.N * (g.dt_c == DT_BF16 ? 2.0 : 4.0); break;
        case EPI_GLU_RES: out = (double)g.M * (g.N / 2) * 8.0; break;        // fp32 residual in, fp32 out
        case EPI_GEGLU: out = (double)g.M * (g.N / 2) * (double)esz; break;
        case EPI_BIAS_RES: out = (double)g.M * g.N * 8.0; break;
        default: break;
    }
    return (double)g.M * g.K * esz + (double)g.N * g.K * esz + out;
}
static double gemm_flops(const GemmArgs& g) { return 2.0 * g.M * (double)g.N * g.K; }

// ------------------------------------------------------------------------------------------------ encoder
struct EncGeom {
    int B = 0;
    std::vector<int> img_off, tok_off;
    long P[5] = {0, 0, 0, 0, 0};
    int ntok = 0, max_tok = 0;
    int uni_h = 0, uni_w = 0;          // > 0: every image has this size (enables the TMA im2col convolutions)
    const int* d_img_off = nullptr; const int* d_img_hw = nullptr; const int* d_tok_off = nullptr;
};

static int plan_geometry(texocr_handle* h, const int32_t* hw, int B, EncGeom& g, cudaStream_t st) {
    if (B <= 0) return fail(h, TEXOCR_ERR_ARG, "batch must be positive");
    g.B = B;
    g.img_off.assign(B + 1, 0); g.tok_off.assign(B + 1, 0);
    g.uni_h = hw[0]; g.uni_w = hw[1];
    for (int b = 0; b < B; ++b) {
        const int H = hw[2 * b], W = hw[2 * b + 1];
        if (H != g.uni_h || W != g.uni_w) g.uni_h = g.uni_w = 0;
        if (H <= 0 || W <= 0 || H % 16 || W % 16 || H > 160 || W > 1008)
            return fail(h, TEXOCR_ERR_ARG, "image %d is %dx%d: height and width must be multiples of 16 with H <= 160 and "
                        "W <= 1008 (10x63 position grid, model/encoder.py:137-143)", b, H, W);
        const long next = (long)g.img_off[b] + (long)H * W;
        if (next > 0x7fffffffL / 64) return fail(h, TEXOCR_ERR_ARG, "batch too large: more than 2^31 stem activations");
        g.img_off[b + 1] = (int)next;
        const int n = (H / 16) * (W / 16) + 1;
        g.tok_off[b + 1] = g.tok_off[b] + n;
        g.max_tok = std::max(g.max_tok, n);
    }
    for (int l = 0; l <= 4; ++l) g.P[l] = g.img_off[B] >> (2 * l);
    g.ntok = g.tok_off[B];
    std::vector<int> v;
    v.insert(v.end(), g.img_off.begin(), g.img_off.end());
    v.insert(v.end(), hw, hw + 2 * B);
    v.insert(v.end(), g.tok_off.begin(), g.tok_off.end());
    int r = upload_ints(h, v, st);
    if (r) return r;
    g.d_img_off = h->geom.as<int>();
    g.d_img_hw = g.d_img_off + (B + 1);
    g.d_tok_off = g.d_img_hw + 2 * B;
    return 0;
}

// GroupNorm work is split into per-image pixel chunks.  The grid always offers GN_MAX_CHUNKS chunks per image and each
// image uses min(GN_MAX_CHUNKS, ceil(pixels/64)) of them, a function of its OWN size only: the summation order of
// an image's statistics -- hence every bit of its result -- does not depend on what else is in the batch.
static int nchunk_for(long, int) { return 32; }

static int run_backbone(texocr_handle* h, const float* d_img, const EncGeom& g, cudaStream_t st, const float** feat_out) {
    const int B = g.B;
    ENSURE(h->raw1, (size_t)g.P[1] * 64 * 4);
    ENSURE(h->act2, (size_t)g.P[2] * 64 * 4);
    ENSURE(h->actA, (size_t)g.P[2] * 256 * 4);
    ENSURE(h->actB, (size_t)g.P[2] * 256 * 4);
    ENSURE(h->rawMid, (size_t)g.P[2] * 128 * 4);
    ENSURE(h->actMid, (size_t)g.P[2] * 128 * 4);
    ENSURE(h->rawMid2, (size_t)g.P[2] * 64 * 4);
    ENSURE(h->actMid2, (size_t)g.P[2] * 64 * 4);
    ENSURE(h->raw3, (size_t)g.P[2] * 256 * 4);
    ENSURE(h->rawDs, (size_t)g.P[2] * 256 * 4);
    ENSURE(h->gn_partial, (size_t)B * 32 * 32 * 2 * 8);
    for (int i = 0; i < 4; ++i) ENSURE(h->gn_stats[i], (size_t)B * 32 * 2 * 4);
    float* stats[4] = {h->gn_stats[0].as<float>(), h->gn_stats[1].as<float>(), h->gn_stats[2].as<float>(), h->gn_stats[3].as<float>()};
    double* partial = h->gn_partial.as<double>();

    // stem: conv 7x7/s2 -> GN+ReLU -> maxpool 3x3/s2  [model/resnet.py:218-222]
    LAUNCH(KC_STEM, 1, (double)g.P[0] * 4 + (double)g.P[1] * 64 * 4, 2.0 * 49 * 64 * g.P[1],
           launch_stem_conv(d_img, h->stem_w, h->raw1.as<float>(), g.d_img_off, g.d_img_hw, B, (int)g.P[1], st));
    int nc = nchunk_for(g.P[1], B);
    LAUNCH(KC_GN_STATS, 2, (double)g.P[1] * 64 * 4, 0.0,
           launch_gn_stats(h->raw1.as<float>(), 64, 1, g.d_img_off, B, nc, partial, stats[0], st));
    LAUNCH(KC_GN_APPLY, 1, (double)g.P[1] * 64 * 4 + (double)g.P[2] * 64 * 4, 0.0,
           launch_gn_apply_maxpool(h->raw1.as<float>(), stats[0], h->stem_g, h->stem_b, h->act2.as<float>(), nullptr, nullptr,
                                   g.d_img_off, g.d_img_hw, B, (int)g.P[2], st));

    const float* x = h->act2.as<float>();
    int Cx = 64, Lx = 2;
    float* pingpong[2] = {h->actA.as<float>(), h->actB.as<float>()};
    int pp = 0;
    size_t ci = 0;
    const int depths[3] = {2, 4, 6};
    auto conv = [&](const ConvW& cw, const float* in, int lin, int lout, float* out) -> int {
        const long M = g.P[lout];
        GemmArgs ga = mk_gemm(in, cw.cin, cw.w, cw.k * cw.k * cw.cin, out, cw.cout, (int)M, cw.cout, cw.k * cw.k * cw.cin,
                              EPI_STORE, DT_F32, DT_F32, nullptr, nullptr, 0);
        ConvGather cg{g.d_img_off, g.d_img_hw, B, lin, lout, cw.k, cw.stride, (cw.k == 3 && cw.stride == 1) ? 1 : 0, cw.cin};
        if (!(cw.k == 1 && cw.stride == 1)) ga.conv = &cg;       // 1x1/s1 is a plain GEMM over the pixel rows
        LAUNCH(KC_CONV, 1, gemm_bytes(ga, 4), gemm_flops(ga), launch_gemm_simt(ga, st));
        return 0;
    };
    auto gstats = [&](const float* raw, int C, int level, float* stt) -> int {
        LAUNCH(KC_GN_STATS, 2, (double)g.P[level] * C * 4, 0.0,
               launch_gn_stats(raw, C, level, g.d_img_off, B, nchunk_for(g.P[level], B), partial, stt, st));
        return 0;
    };
    int r;
    for (int s = 0; s < 3; ++s)
        for (int b = 0; b < depths[s]; ++b) {
            const ConvW* ds = nullptr;
            if (b == 0) ds = &h->convs[ci++];
            const ConvW& c1 = h->convs[ci++];
            const ConvW& c2 = h->convs[ci++];
            const ConvW& c3 = h->convs[ci++];
            const int Lout = Lx + (c2.stride == 2 ? 1 : 0);
            if (ds) {
                if ((r = conv(*ds, x, Lx, Lout, h->rawDs.as<float>()))) return r;
                if ((r = gstats(h->rawDs.as<float>(), ds->cout, Lout, stats[3]))) return r;
            }
            if ((r = conv(c1, x, Lx, Lx, h->rawMid.as<float>()))) return r;
            if ((r = gstats(h->rawMid.as<float>(), c1.cout, Lx, stats[0]))) return r;
            {
                GnApplyArgs a{};
                a.raw = h->rawMid.as<float>(); a.stats = stats[0]; a.gamma = c1.gamma; a.beta = c1.beta;
                a.out = h->actMid.as<float>(); a.C = c1.cout; a.level = Lx; a.relu = 1;
                LAUNCH(KC_GN_APPLY, 1, (double)g.P[Lx] * a.C * 8, 0.0, launch_gn_apply(a, g.d_img_off, B, nchunk_for(g.P[Lx], B), st));
            }
            if ((r = conv(c2, h->actMid.as<float>(), Lx, Lout, h->rawMid2.as<float>()))) return r;
            if ((r = gstats(h->rawMid2.as<float>(), c2.cout, Lout, stats[1]))) return r;
            {
                GnApplyArgs a{};
                a.raw = h->rawMid2.as<float>(); a.stats = stats[1]; a.gamma = c2.gamma; a.beta = c2.beta;
                a.out = h->actMid2.as<float>(); a.C = c2.cout; a.level = Lout; a.relu = 1;
                LAUNCH(KC_GN_APPLY, 1, (double)g.P[Lout] * a.C * 8, 0.0, launch_gn_apply(a, g.d_img_off, B, nchunk_for(g.P[Lout], B), st));
            }
            if ((r = conv(c3, h->actMid2.as<float>(), Lout, Lout, h->raw3.as<float>()))) return r;
            if ((r = gstats(h->raw3.as<float>(), c3.cout, Lout, stats[2]))) return r;
            {   // out = ReLU(GN3(y) + res)   [model/resnet.py:147-148]
                GnApplyArgs a{};
                a.raw = h->raw3.as<float>(); a.stats = stats[2]; a.gamma = c3.gamma; a.beta = c3.beta;
                if (ds) { a.raw2 = h->rawDs.as<float>(); a.stats2 = stats[3]; a.gamma2 = ds->gamma; a.beta2 = ds->beta; }
                else a.res = x;
                a.out = pingpong[pp]; a.C = c3.cout; a.level = Lout; a.relu = 1;
                LAUNCH(KC_GN_APPLY, 1, (double)g.P[Lout] * a.C * 12, 0.0, launch_gn_apply(a, g.d_img_off, B, nchunk_for(g.P[Lout], B), st));
            }
            x = pingpong[pp]; pp ^= 1; Cx = c3.cout; Lx = Lout;
        }
    (void)Cx;
    h->last_backbone_pixels = (int)g.P[4];
    *feat_out = x;
    return 0;
}

// bf16 tier: the same backbone with every convolution on the tensor cores in bf16x3 mode (tc_gemm.cu, SPLIT=3).
// Activations live as split-bf16 pairs (hi | lo halves of one buffer); 1x1/s1 convs are plain GEMMs over pixel rows,
// 3x3 and stride-2 convs go through an explicit im2col of the pair.  Raw conv outputs and all statistics stay fp32.
static int run_backbone_tc(texocr_handle* h, const float* d_img, const EncGeom& g, cudaStream_t st, const void** feat_hi,
                           const void** feat_lo) {
    const int B = g.B;
    ENSURE(h->raw1, (size_t)g.P[1] * 64 * 4);
    ENSURE(h->act2, (size_t)g.P[2] * 64 * 4);
    ENSURE(h->actA, (size_t)g.P[2] * 256 * 4);
    ENSURE(h->actB, (size_t)g.P[2] * 256 * 4);
    ENSURE(h->rawMid, (size_t)g.P[2] * 128 * 4);
    ENSURE(h->actMid, (size_t)g.P[2] * 128 * 4);
    ENSURE(h->rawMid2, (size_t)g.P[2] * 64 * 4);
    ENSURE(h->actMid2, (size_t)g.P[2] * 64 * 4);
    ENSURE(h->raw3, (size_t)g.P[2] * 256 * 4);
    ENSURE(h->rawDs, (size_t)g.P[2] * 256 * 4);
    ENSURE(h->col, (size_t)g.P[2] * 576 * 4);
    ENSURE(h->gn_partial, (size_t)B * 32 * 32 * 2 * 8);
    for (int i = 0; i < 4; ++i) ENSURE(h->gn_stats[i], (size_t)B * 32 * 2 * 4);
    float* stats[4] = {h->gn_stats[0].as<float>(), h->gn_stats[1].as<float>(), h->gn_stats[2].as<float>(), h->gn_stats[3].as<float>()};
    double* partial = h->gn_partial.as<double>();
    struct Pair { char* hi; char* lo; };
    auto pair_of = [](DevBuf& b, size_t elems) { Pair p; p.hi = (char*)b.p; p.lo = (char*)b.p + elems * 2; return p; };

    LAUNCH(KC_STEM, 1, (double)g.P[0] * 4 + (double)g.P[1] * 64 * 4, 2.0 * 49 * 64 * g.P[1],
           launch_stem_conv(d_img, h->stem_w, h->raw1.as<float>(), g.d_img_off, g.d_img_hw, B, (int)g.P[1], st));
    LAUNCH(KC_GN_STATS, 2, (double)g.P[1] * 64 * 4, 0.0,
           launch_gn_stats(h->raw1.as<float>(), 64, 1, g.d_img_off, B, nchunk_for(g.P[1], B), partial, stats[0], st));
    Pair x = pair_of(h->act2, (size_t)g.P[2] * 64);
    LAUNCH(KC_GN_APPLY, 1, (double)g.P[1] * 64 * 4 + (double)g.P[2] * 64 * 4, 0.0,
           launch_gn_apply_maxpool(h->raw1.as<float>(), stats[0], h->stem_g, h->stem_b, nullptr, x.hi, x.lo, g.d_img_off,
                                   g.d_img_hw, B, (int)g.P[2], st));
    int Lx = 2;
    DevBuf* pingpong[2] = {&h->actA, &h->actB};
    int pp = 0;
    size_t ci = 0;
    const int depths[3] = {2, 4, 6};
    // conv: split GEMM, through im2col unless 1x1/s1
    auto conv = [&](const ConvW& cw, Pair in, int lin, int lout, float* out) -> int {
        const long M = g.P[lout];
        const int K = cw.k * cw.k * cw.cin;
        const void *a_hi = in.hi, *a_lo = in.lo;
        const int pad_lo = (cw.k == 3 && cw.stride == 1) ? 1 : 0;
        // same-size batch: the GEMM fetches its A tiles straight from the NHWC activation with TMA im2col loads
        const bool implicit = h->use_im2col_tma && g.uni_h > 0 && cw.cin % 64 == 0 && !(cw.k == 1 && cw.stride == 1);
        if (!(cw.k == 1 && cw.stride == 1) && !implicit) {
            ConvGather cg{g.d_img_off, g.d_img_hw, B, lin, lout, cw.k, cw.stride, pad_lo, cw.cin};
            Pair c = pair_of(h->col, (size_t)M * K);
            LAUNCH(KC_GN_APPLY, 1, (double)M * K * 8, 0.0, launch_im2col_split(in.hi, in.lo, c.hi, c.lo, cg, M, st));
            a_hi = c.hi; a_lo = c.lo;
        }
        GemmArgs ga = mk_gemm(a_hi, K, cw.w_hi, K, out, cw.cout, (int)M, cw.cout, K, EPI_STORE, DT_BF16, DT_F32, nullptr, nullptr, 0);
        ga.A2 = a_lo; ga.W2 = cw.w_lo;
        if (implicit) {
            ga.lda = cw.cin;
            ga.im2col = {cw.k, cw.stride, pad_lo, cw.cin, g.uni_w >> lin, g.uni_h >> lin, B, g.uni_w >> lout, g.uni_h >> lout};
        }
        if (!tc_gemm_supported(ga)) return fail(h, TEXOCR_ERR_ARG, "backbone conv %s not supported by the tcgen05 GEMM", cw.name.c_str());
        LAUNCH(KC_CONV, 1, gemm_bytes(ga, 4), gemm_flops(ga), launch_gemm_tc(ga, st));
        return 0;
    };
    auto gstats = [&](const float* raw, int C, int level, float* stt) -> int {
        LAUNCH(KC_GN_STATS, 2, (double)g.P[level] * C * 4, 0.0,
               launch_gn_stats(raw, C, level, g.d_img_off, B, nchunk_for(g.P[level], B), partial, stt, st));
        return 0;
    };
    auto gapply = [&](GnApplyArgs& a, Pair out, int level, double bytes_per) -> int {
        a.out = nullptr; a.out_hi = out.hi; a.out_lo = out.lo; a.level = level;
        LAUNCH(KC_GN_APPLY, 1, (double)g.P[level] * a.C * bytes_per, 0.0, launch_gn_apply(a, g.d_img_off, B, nchunk_for(g.P[level], B), st));
        return 0;
    };
    int r;
    for (int s = 0; s < 3; ++s)
        for (int b = 0; b < depths[s]; ++b) {
            const ConvW* ds = nullptr;
            if (b == 0) ds = &h->convs[ci++];
            const ConvW& c1 = h->convs[ci++];
            const ConvW& c2 = h->convs[ci++];
            const ConvW& c3 = h->convs[ci++];
            const int Lout = Lx + (c2.stride == 2 ? 1 : 0);
            if (ds) {
                if ((r = conv(*ds, x, Lx, Lout, h->rawDs.as<float>()))) return r;
                if ((r = gstats(h->rawDs.as<float>(), ds->cout, Lout, stats[3]))) return r;
            }
            if ((r = conv(c1, x, Lx, Lx, h->rawMid.as<float>()))) return r;
            if ((r = gstats(h->rawMid.as<float>(), c1.cout, Lx, stats[0]))) return r;
            Pair m1 = pair_of(h->actMid, (size_t)g.P[Lx] * c1.cout);
            {
                GnApplyArgs a{};
                a.raw = h->rawMid.as<float>(); a.stats = stats[0]; a.gamma = c1.gamma; a.beta = c1.beta; a.C = c1.cout; a.relu = 1;
                if ((r = gapply(a, m1, Lx, 8))) return r;
            }
            if ((r = conv(c2, m1, Lx, Lout, h->rawMid2.as<float>()))) return r;
            if ((r = gstats(h->rawMid2.as<float>(), c2.cout, Lout, stats[1]))) return r;
            Pair m2 = pair_of(h->actMid2, (size_t)g.P[Lout] * c2.cout);
            {
                GnApplyArgs a{};
                a.raw = h->rawMid2.as<float>(); a.stats = stats[1]; a.gamma = c2.gamma; a.beta = c2.beta; a.C = c2.cout; a.relu = 1;
                if ((r = gapply(a, m2, Lout, 8))) return r;
            }
            if ((r = conv(c3, m2, Lout, Lout, h->raw3.as<float>()))) return r;
            if ((r = gstats(h->raw3.as<float>(), c3.cout, Lout, stats[2]))) return r;
            Pair out = pair_of(*pingpong[pp], (size_t)g.P[Lout] * c3.cout);
            {
                GnApplyArgs a{};
                a.raw = h->raw3.as<float>(); a.stats = stats[2]; a.gamma = c3.gamma; a.beta = c3.beta; a.C = c3.cout; a.relu = 1;
                if (ds) { a.raw2 = h->rawDs.as<float>(); a.stats2 = stats[3]; a.gamma2 = ds->gamma; a.beta2 = ds->beta; }
                else { a.res_hi = x.hi; a.res_lo = x.lo; }
                if ((r = gapply(a, out, Lout, 12))) return r;
            }
            x = out; pp ^= 1; Lx = Lout;
        }
    h->last_backbone_pixels = 0;     // the fp32 tap is only kept by the fp32 tier
    *feat_hi = x.hi; *feat_lo = x.lo;
    return 0;
}

// One (self-attention | cross-attention | MLP) sub-layer tail shared by encoder / decoder / decode step.
struct RowCtx {
    int rows; int kc_gemm, kc_row;
    const float* ln_g; const float* ln_b;
    int row0 = 0;        // first row of this sub-batch inside the row workspaces
};
static inline float* rowf(const DevBuf& b, const RowCtx& rc, int width) { return b.as<float>() + (size_t)rc.row0 * width; }
static inline void* rowa(texocr_handle* h, const DevBuf& b, const RowCtx& rc, int width) { return (char*)b.p + (size_t)rc.row0 * width * h->esz; }

static int sub_attn_out(texocr_handle* h, const RowCtx& rc, const AttnW& w, cudaStream_t st) {
    // y = o.Wo^T + bo -> GLU -> + residual   [model/attention.py:96-99,180 ; 254]
    GemmArgs ga = mk_gemm(rowa(h, h->o, rc, 512), 512, w.wo, 512, rowf(h->s, rc, 256), 256, rc.rows, 512, 512, EPI_GLU_RES, h->dt, DT_F32, w.bo, rowf(h->x, rc, 256), 256);
    LAUNCH(rc.kc_gemm == KC_DEC_GEMM ? KC_DEC_GEMM_WO : rc.kc_gemm, 1, gemm_bytes(ga, h->esz), gemm_flops(ga), run_gemm(h, ga, st));
    return 0;
}
static int sub_mlp(texocr_handle* h, const RowCtx& rc, const MlpW& w, cudaStream_t st) {
    GemmArgs g1 = mk_gemm(rowa(h, h->xn, rc, 256), 256, w.w1, 256, rowa(h, h->hid, rc, 1024), 1024, rc.rows, 2048, 256, EPI_GEGLU, h->dt, h->dt, w.b1, nullptr, 0);
    LAUNCH(rc.kc_gemm == KC_DEC_GEMM ? KC_DEC_GEMM_W1 : rc.kc_gemm, 1, gemm_bytes(g1, h->esz), gemm_flops(g1), run_gemm(h, g1, st));
    GemmArgs g2 = mk_gemm(rowa(h, h->hid, rc, 1024), 1024, w.w2, 1024, rowf(h->s, rc, 256), 256, rc.rows, 256, 1024, EPI_BIAS_RES, h->dt, DT_F32, w.b2, rowf(h->x, rc, 256), 256);
    LAUNCH(rc.kc_gemm == KC_DEC_GEMM ? KC_DEC_GEMM_W2 : rc.kc_gemm, 1, gemm_bytes(g2, h->esz), gemm_flops(g2), run_gemm(h, g2, st));
    return 0;
}
// x = LN(s); xn = LN(x)  (shared LayerNorm twice, model/attention.py:242-259), or the stack's final norm.
static int sub_norm(texocr_handle* h, const RowCtx& rc, bool last, const float* fin_g, const float* fin_b, float* fin_out_f,
                    void* fin_out_a, cudaStream_t st) {
    if ((h->dbg_skip & 4) && rc.kc_row == KC_DEC_ROW) return 0;
    Ln2Args a{};
    a.in = rowf(h->s, rc, 256); a.rows = rc.rows; a.dt_a = h->dt;
    if (!last) { a.g1 = rc.ln_g; a.b1 = rc.ln_b; a.g2 = rc.ln_g; a.b2 = rc.ln_b; a.o1f = rowf(h->x, rc, 256); a.o2a = rowa(h, h->xn, rc, 256); }
    else { a.g1 = fin_g; a.b1 = fin_b; a.o1f = fin_out_f; a.o1a = fin_out_a; }
    LAUNCH(rc.kc_row, 1, (double)rc.rows * 256 * (4 + 4 + h->esz), 0.0, launch_ln2(a, st));
    return 0;
}

static int ensure_rows(texocr_handle* h, long rows) {
    ENSURE(h->x, (size_t)rows * 256 * 4);
    ENSURE(h->s, (size_t)rows * 256 * 4);
    ENSURE(h->xn, (size_t)rows * 256 * h->esz);
    ENSURE(h->qkv, (size_t)rows * 1536 * h->esz);
    ENSURE(h->o, (size_t)rows * 512 * h->esz);
    ENSURE(h->hid, (size_t)rows * 1024 * h->esz);
    return 0;
}

// images (device) -> h->enc_out (fp32) and h->enc_a (GEMM operand type)
static int run_encoder(texocr_handle* h, const float* d_img, const EncGeom& g, cudaStream_t st) {
    const texocr_config& c = h->cfg;
    int r;
    ENSURE(h->proj_out, (size_t)g.P[4] * 256 * 4);
    if (c.encoder_kind == TEXOCR_ENC_HYBRID && h->dt == DT_BF16 && h->use_tcgen05) {
        const void *fh = nullptr, *fl = nullptr;
        if ((r = run_backbone_tc(h, d_img, g, st, &fh, &fl))) return r;
        GemmArgs ga = mk_gemm(fh, 1024, h->proj_w, 1024, h->proj_out.p, 256, (int)g.P[4], 256, 1024, EPI_STORE, DT_BF16, DT_F32, h->proj_b, nullptr, 0);
        ga.A2 = fl; ga.W2 = h->proj_w_lo;
        LAUNCH(KC_ENC_GEMM, 1, gemm_bytes(ga, 4), gemm_flops(ga), launch_gemm_tc(ga, st));
    } else if (c.encoder_kind == TEXOCR_ENC_HYBRID) {
        const float* feat = nullptr;
        if ((r = run_backbone(h, d_img, g, st, &feat))) return r;
        const void* a_ptr = feat;
        if (h->dt != DT_F32) {
            ENSURE(h->backbone_a, (size_t)g.P[4] * 1024 * h->esz);
            LAUNCH(KC_MISC, 1, (double)g.P[4] * 1024 * 6, 0.0, launch_cast_f32_to(feat, h->backbone_a.p, g.P[4] * 1024, h->dt, st));
            a_ptr = h->backbone_a.p;
        }
        GemmArgs ga = mk_gemm(a_ptr, 1024, h->proj_w, 1024, h->proj_out.p, 256, (int)g.P[4], 256, 1024, EPI_STORE, h->dt, DT_F32, h->proj_b, nullptr, 0);
        LAUNCH(KC_ENC_GEMM, 1, gemm_bytes(ga, h->esz), gemm_flops(ga), run_gemm(h, ga, st));
    } else {
        ENSURE(h->patch_cols, (size_t)g.P[4] * 256 * 4);
        LAUNCH(KC_MISC, 1, (double)g.P[0] * 8, 0.0, launch_im2col_patch(d_img, h->patch_cols.as<float>(), g.d_img_off, g.d_img_hw, g.B, (int)g.P[4], st));
        const void* a_ptr = h->patch_cols.p;
        if (h->dt != DT_F32) {
            ENSURE(h->backbone_a, (size_t)g.P[4] * 256 * h->esz);
            LAUNCH(KC_MISC, 1, (double)g.P[4] * 256 * 6, 0.0, launch_cast_f32_to(h->patch_cols.as<float>(), h->backbone_a.p, g.P[4] * 256, h->dt, st));
            a_ptr = h->backbone_a.p;
        }
        GemmArgs ga = mk_gemm(a_ptr, 256, h->proj_w, 256, h->proj_out.p, 256, (int)g.P[4], 256, 256, EPI_STORE, h->dt, DT_F32, h->proj_b, nullptr, 0);
        LAUNCH(KC_ENC_GEMM, 1, gemm_bytes(ga, h->esz), gemm_flops(ga), run_gemm(h, ga, st));
    }
    const int R = g.ntok;
    if ((r = ensure_rows(h, R))) return r;
    ENSURE(h->enc_out, (size_t)R * 256 * 4);
    ENSURE(h->enc_a, (size_t)R * 256 * h->esz);
    LAUNCH(KC_ENC_ROW, 1, (double)R * 256 * 12, 0.0,
           launch_assemble_tokens(h->proj_out.as<float>(), h->cls, h->pos, h->x.as<float>(), g.d_img_off, g.d_img_hw, g.d_tok_off, g.B, R, st));
    RowCtx rc{R, KC_ENC_GEMM, KC_ENC_ROW, h->enc_ln_g, h->enc_ln_b};
    {
        Ln2Args a{};
        a.in = h->x.as<float>(); a.rows = R; a.g2 = h->enc_ln_g; a.b2 = h->enc_ln_b; a.o2a = h->xn.p; a.dt_a = h->dt;
        LAUNCH(KC_ENC_ROW, 1, (double)R * 256 * (4 + h->esz), 0.0, launch_ln2(a, st));
    }
    for (int l = 0; l < c.enc_layers; ++l) {
        GemmArgs gq = mk_gemm(h->xn.p, 256, h->enc_attn[l].wqkv, 256, h->qkv.p, 1536, R, 1536, 256, EPI_STORE, h->dt, h->dt, nullptr, nullptr, 0);
        LAUNCH(KC_ENC_GEMM, 1, gemm_bytes(gq, h->esz), gemm_flops(gq), run_gemm(h, gq, st));
        AttnVarlenArgs av{};
        const char* base = (const char*)h->qkv.p;
        av.q = base; av.k = base + 512 * h->esz; av.v = base + 1024 * h->esz; av.ldq = av.ldk = av.ldv = 1536;
        av.o = h->o.p; av.ldo = 512; av.q_off = g.d_tok_off; av.k_off = g.d_tok_off; av.batch = g.B; av.max_q = g.max_tok;
        av.causal = 0; av.dt = h->dt;
        double aflops = 0.0;
        for (int b = 0; b < g.B; ++b) { const double n = g.tok_off[b + 1] - g.tok_off[b]; aflops += 4.0 * n * n * 512; }
        if (h->use_tcgen05 && attn_enc_mma_supported(av)) LAUNCH(KC_ENC_ATTN, 1, (double)R * 2048 * h->esz, aflops, launch_attn_enc_mma(av, st));
        else LAUNCH(KC_ENC_ATTN, 1, (double)R * 2048 * h->esz, aflops, launch_attn_varlen(av, st));
        if ((r = sub_attn_out(h, rc, h->enc_attn[l], st))) return r;
        if ((r = sub_norm(h, rc, false, nullptr, nullptr, nullptr, nullptr, st))) return r;
        if ((r = sub_mlp(h, rc, h->enc_mlp[l], st))) return r;
        const bool last = (l == c.enc_layers - 1);
        if ((r = sub_norm(h, rc, last, h->enc_norm_g, h->enc_norm_b, h->enc_out.as<float>(), h->dt == DT_F32 ? nullptr : h->enc_a.p, st))) return r;
    }
    return 0;
}

// memory (device fp32 or already-typed copy) -> h->crosskv [ntok, L*1024]   (K/V of every cross-attention layer, once)
// y = out-projection(C) + bo -> GLU -> + residual for the absorbed attention, C = [rows, 8 x 256] softmax-weighted latent averages:
// either one folded GEMM (K = 2048), or the per-head value projection (block-diagonal, K = 256) into `o` followed by the ordinary
// Wo GEMM -- 4x fewer FLOPs / weight bytes and a shorter dependent chain (12 instead of 32 k-blocks)
static int sub_abs_out(texocr_handle* h, const RowCtx& rc, const AttnW& w, const void* ca, cudaStream_t st) {
    if (!h->absorb_two_stage) {
        GemmArgs go = mk_gemm(ca, 2048, w.wvo, 2048, rowf(h->s, rc, 256), 256, rc.rows, 512, 2048, EPI_GLU_RES, h->dt, DT_F32, w.bo, rowf(h->x, rc, 256), 256);
        LAUNCH(rc.kc_gemm == KC_DEC_GEMM ? KC_DEC_GEMM_WO : rc.kc_gemm, 1, gemm_bytes(go, h->esz), gemm_flops(go), run_gemm(h, go, st));
        return 0;
    }
    GemmArgs gv = mk_gemm(ca, 2048, w.wv, 256, rowa(h, h->o, rc, 512), 512, rc.rows, 512, 256, EPI_STORE, h->dt, h->dt, nullptr, nullptr, 0);
    gv.a_block_k = 256;
    // algorithmic work of the block-diagonal GEMM: every row contracts 8 heads x (256 -> 64)
    LAUNCH(rc.kc_gemm == KC_DEC_GEMM ? KC_DEC_GEMM_VPROJ : rc.kc_gemm, 1, (double)rc.rows * 2048 * h->esz + 512.0 * 256 * h->esz + (double)rc.rows * 512 * h->esz,
           2.0 * rc.rows * 512.0 * 256, run_gemm(h, gv, st));
    return sub_attn_out(h, rc, w, st);
}

// generate loop, bf16 tier: absorbed cross-attention (needs the TMA attention path and the per-branch kernel graphs)
static bool use_absorb(const texocr_handle* h) {
    return h->dt == DT_BF16 && h->use_tcgen05 && (h->use_tma_attn == 1 || h->use_tma_attn == 3) && !h->decode_mega &&
           !h->fuse_ln;
}

static int run_crosskv(texocr_handle* h, const float* enc_f32, const void* enc_typed, int ntok, cudaStream_t st, bool for_generate = false) {
    const int L = h->cfg.dec_layers;
    h->self_abs_active = for_generate && h->self_absorb && use_absorb(h);
    if (for_generate && h->cross_absorb && use_absorb(h)) {      // no K/V projection at all: the decode loop streams the bf16 encoder memory itself
        if (!enc_typed) {
            ENSURE(h->enc_a, (size_t)ntok * 256 * h->esz);
            LAUNCH(KC_MISC, 1, (double)ntok * 256 * 6, 0.0, launch_cast_f32_to(enc_f32, h->enc_a.p, (int64_t)ntok * 256, h->dt, st));
            enc_typed = h->enc_a.p;
        }
        h->dec_enc = enc_typed;
        h->crosskv_rows = ntok;
        return 0;
    }
    h->dec_enc = nullptr;
    ENSURE(h->crosskv, (size_t)ntok * L * 1024 * h->esz);
    h->crosskv_rows = ntok;
    const void* a_ptr = enc_f32;
    if (h->dt != DT_F32) {
        if (!enc_typed) {
            ENSURE(h->enc_a, (size_t)ntok * 256 * h->esz);
            LAUNCH(KC_MISC, 1, (double)ntok * 256 * 6, 0.0, launch_cast_f32_to(enc_f32, h->enc_a.p, (int64_t)ntok * 256, h->dt, st));
            enc_typed = h->enc_a.p;
        }
        a_ptr = enc_typed;
    }
    GemmArgs ga = mk_gemm(a_ptr, 256, h->w_crosskv, 256, h->crosskv.p, L * 1024, ntok, L * 1024, 256, EPI_STORE, h->dt, h->dt, nullptr, nullptr, 0);
    LAUNCH(KC_CROSSKV_GEMM, 1, gemm_bytes(ga, h->esz), gemm_flops(ga), run_gemm(h, ga, st));
    return 0;
}

// ------------------------------------------------------------------------------------------------ decode step
constexpr int MAX_BRANCH = 16;
constexpr int FIFO_STRIDE = 32;      // events per (step, branch) of the coupled graph: 2 attention launches per decoder layer
static int sampling_k(const texocr_handle* h);
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
    const bool fuse = h->fuse_ln && h->dt == DT_BF16 && h->use_tcgen05 && !(h->dbg_skip & 12);
    float* sbuf = rowf(h->s, rc, 256);
    float* xbuf = rowf(h->x, rc, 256);
    void* xnbuf = rowa(h, h->xn, rc, 256);
    // GEMM whose A operand is the (double) LayerNorm of the residual stream in `s`: fused kernel, or LN kernel + plain GEMM
    auto gemm_ln = [&](const void* W, int N, void* C, int ldc, int epi, int dt_c, const float* bias, bool first_ln, bool write_x,
                       const float* fin_g, const float* fin_b) -> int {
        GemmArgs ga = mk_gemm(xnbuf, 256, W, 256, C, ldc, rows, N, 256, epi, h->dt, dt_c, bias, nullptr, 0);
        const float *g1 = nullptr, *b1 = nullptr, *g2 = h->dec_ln_g, *b2 = h->dec_ln_b;
        if (fin_g) { g1 = fin_g; b1 = fin_b; g2 = nullptr; b2 = nullptr; }
        else if (first_ln) { g1 = h->dec_ln_g; b1 = h->dec_ln_b; }
        LAUNCH(KC_DEC_GEMM, 1, gemm_bytes(ga, e) + (double)rows * 256 * 8, gemm_flops(ga),
               launch_gemm_tc_ln(ga, sbuf, g1, b1, g2, b2, write_x ? xbuf : nullptr, st));
        return 0;
    };
    // Coupled capture (run_generate, attn_fifo > 0): attention launch k of this branch waits for the attention launch the
    // FIFO names and records its own completion for the branch behind it.
    struct PdlGuard { int saved; PdlGuard() : saved(g_texocr_pdl) {} ~PdlGuard() { g_texocr_pdl = saved; } };
    auto fifo_before = [&](int k) -> int {
        if (h->fifo_wait && h->fifo_wait[k]) {
            CK(cudaStreamWaitEvent(st, h->fifo_wait[k], 0));
            if (!h->fifo_pdl) g_texocr_pdl &= ~((1 << PDL_ATTN_TMA) | (1 << PDL_ATTN_SIMPLE));
        }
        return 0;
    };
    auto fifo_after = [&](int k) -> int {
        if (h->fifo_rec) CK(cudaEventRecord(h->fifo_rec[k], st));
        return 0;
    };
    // the step's input (x = embedding, xn = LN(x)) was written by enqueue_first_embed (step 0) or by the previous step's token kernel
    const double tkeys = t_host >= 0 ? (double)(t_host + 1) : 0.0;
    char* qb = (char*)rowa(h, h->qkv, rc, 1536);
    for (int l = 0; l < L; ++l) {
        // ---- causal self-attention, absorbed form: the cache holds this layer's LayerNorm'd inputs (256 per position), Q' = xn . Wqk^T,
        // C_h = softmax(Q'_h . Z^T / 8) . Z over positions 0..t (t = this step's own row), y = C . Wvo^T + bo -> GLU -> + residual
        if (h->self_abs_active && !fuse) {
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
            {
                PdlGuard guard;
                if ((r = fifo_before(2 * l))) return r;
                if (!(h->dbg_skip & 1))
                    LAUNCH(KC_DEC_ATTN_SELF, 1, (double)rows * tkeys * 256 * e, 4.0 * rows * tkeys * 2048,
                           launch_attn_abs(ab, h->num_sms * h->attn_ctas_per_sm, st));
            }
            if ((r = fifo_after(2 * l))) return r;
            if ((r = sub_abs_out(h, rc, h->dec_self[l], ca, st))) return r;
        } else {
        // ---- causal self-attention over the KV cache
        if (fuse) {
            // layer 0: x = embedding (no LayerNorm before the first block input's residual); later layers: x = LN(s)
            if ((r = gemm_ln(h->dec_self[l].wqkv, 1536, qb, 1536, EPI_STORE, h->dt, nullptr, l > 0, true, nullptr, nullptr))) return r;
        } else {
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
        {
            PdlGuard guard;
            if ((r = fifo_before(2 * l))) return r;
            if (h->dbg_skip & 1) {}
            else if ((h->use_tma_attn == 1 || h->use_tma_attn == 2) && attn_decode_tma_supported(ad))
                LAUNCH(KC_DEC_ATTN_SELF, 1, (double)rows * tkeys * 1024 * e, 4.0 * rows * tkeys * 512,
                       launch_attn_decode_tma(ad, lay_self, h->num_sms * h->attn_ctas_per_sm, st));
            else
                LAUNCH(KC_DEC_ATTN_SELF, 1, (double)rows * tkeys * 1024 * e, 4.0 * rows * tkeys * 512, launch_attn_decode(ad, tcap, st));
        }
        if ((r = fifo_after(2 * l))) return r;
        if ((r = sub_attn_out(h, rc, h->dec_self[l], st))) return r;
        }
        // ---- cross-attention, absorbed form: Q' = LN.LN(s) . Wqk^T (8 x 256 per row), C_h = softmax(Q'_h . enc^T / 8) . enc over the
        // bf16 encoder memory itself, y = C . Wvo^T + bo -> GLU -> + residual
        if (h->dec_enc && !fuse) {
            if ((r = sub_norm(h, rc, false, nullptr, nullptr, nullptr, nullptr, st))) return r;
            void* qa = rowa(h, h->qabs, rc, 2048);
            void* ca = rowa(h, h->cabs, rc, 2048);
            GemmArgs gc = mk_gemm(xnbuf, 256, h->dec_cross[l].wqk, 256, qa, 2048, rows, 2048, 256, EPI_STORE, h->dt, h->dt, nullptr, nullptr, 0);
            LAUNCH(KC_DEC_GEMM_Q, 1, (double)rows * 256 * e + 2048.0 * 256 * e + (double)rows * 2048 * e, gemm_flops(gc), run_gemm(h, gc, st));
            AttnAbsArgs ab{};
            ab.q = qa; ab.ldq = 2048; ab.latent = h->dec_enc; ab.latent_rows = h->crosskv_rows; ab.k_off = d_enc_off + row0; ab.o = ca; ab.ldo = 2048; ab.batch = rows;
            if (h->attn_trace_on && h->attn_trace.p && 2 * l + 1 < 8) {
                ab.trace = h->attn_trace.as<unsigned long long>() + (size_t)branch * 3 * 2048; ab.trace_step = step; ab.trace_k = 2 * l + 1;
                ab.dbg = h->attn_trace.as<unsigned long long>() + (size_t)MAX_BRANCH * 3 * 2048 + 8;
            }
            {
                PdlGuard guard;
                if ((r = fifo_before(2 * l + 1))) return r;
                if (!(h->dbg_skip & 2))
                    LAUNCH(KC_DEC_ATTN_CROSS, 1, sum_s * rows / B * 256 * e, 4.0 * sum_s * rows / B * 2048,
                           launch_attn_abs(ab, h->num_sms * h->attn_ctas_per_sm, st));
            }
            if ((r = fifo_after(2 * l + 1))) return r;
            if ((r = sub_abs_out(h, rc, h->dec_cross[l], ca, st))) return r;
        } else {
        // ---- cross-attention over the (pre-projected) encoder memory; q goes to the first 512 columns of this branch's qkv rows
        if (fuse) {
            if ((r = gemm_ln(h->dec_cross[l].wq, 512, qb, 512, EPI_STORE, h->dt, nullptr, true, true, nullptr, nullptr))) return r;
        } else {
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
        {
            PdlGuard guard;
            if ((r = fifo_before(2 * l + 1))) return r;
            if (h->dbg_skip & 2) {}
            else if ((h->use_tma_attn == 1 || h->use_tma_attn == 3) && attn_decode_tma_supported(ac))
                LAUNCH(KC_DEC_ATTN_CROSS, 1, sum_s * rows / B * 1024 * e, 4.0 * sum_s * rows / B * 512,
                       launch_attn_decode_tma(ac, lay_cross, h->num_sms * h->attn_ctas_per_sm, st));
            else
                LAUNCH(KC_DEC_ATTN_CROSS, 1, sum_s * rows / B * 1024 * e, 4.0 * sum_s * rows / B * 512, launch_attn_decode(ac, max_s, st));
        }
        if ((r = fifo_after(2 * l + 1))) return r;
        if ((r = sub_attn_out(h, rc, h->dec_cross[l], st))) return r;
        }
        // ---- GeGLU MLP
        if (fuse) {
            if ((r = gemm_ln(h->dec_mlp[l].w1, 2048, rowa(h, h->hid, rc, 1024), 1024, EPI_GEGLU, h->dt, h->dec_mlp[l].b1, true, true, nullptr, nullptr))) return r;
            GemmArgs g2 = mk_gemm(rowa(h, h->hid, rc, 1024), 1024, h->dec_mlp[l].w2, 1024, sbuf, 256, rows, 256, 1024, EPI_BIAS_RES, h->dt, DT_F32,
                                  h->dec_mlp[l].b2, xbuf, 256);
            LAUNCH(KC_DEC_GEMM, 1, gemm_bytes(g2, e), gemm_flops(g2), run_gemm(h, g2, st));
        } else {
            if ((r = sub_norm(h, rc, false, nullptr, nullptr, nullptr, nullptr, st))) return r;
            if ((r = sub_mlp(h, rc, h->dec_mlp[l], st))) return r;
            if ((r = sub_norm(h, rc, l == L - 1, h->dec_norm_g, h->dec_norm_b, nullptr, xnbuf, st))) return r;
        }
    }
    float* lg = h->logits.as<float>() + (size_t)row0 * c.vocab_size;
    // greedy bf16 tier: the vocabulary GEMM reduces every 32-column tile to (max, first index) in its epilogue and the token
    // kernel finishes the argmax over those partials -- the logits never reach HBM (model/decoder.py:60,103 takes
    // logits[:, -1] of a full (B, T, V) tensor).  Sampling and the keep_logits debug option need the row of logits itself.
    const int nparts = (c.vocab_size + 31) / 32;
    const bool fused_amax = !fuse && h->dt == DT_BF16 && h->use_tcgen05 && h->samp_temp <= 0.0 && !h->keep_logits && h->amax_part.p;
    float2* parts = fused_amax ? h->amax_part.as<float2>() + (size_t)row0 * nparts : nullptr;
    if (fused_amax) {
        GemmArgs gl = mk_gemm(xnbuf, 256, h->w_logits, 256, parts, nparts, rows, c.vocab_size, 256, EPI_ARGMAX, h->dt, DT_F32, h->b_logits, nullptr, 0);
        LAUNCH(KC_DEC_GEMM_LOGITS, 1, (double)rows * 256 * e + (double)c.vocab_size * 256 * e + (double)rows * nparts * 8, gemm_flops(gl), run_gemm(h, gl, st));
    } else if (fuse) {
        if ((r = gemm_ln(h->w_logits, c.vocab_size, lg, c.vocab_size, EPI_STORE, DT_F32, h->b_logits, false, false, h->dec_norm_g, h->dec_norm_b))) return r;
    } else {
        GemmArgs gl = mk_gemm(xnbuf, 256, h->w_logits, 256, lg, c.vocab_size, rows, c.vocab_size, 256, EPI_STORE, h->dt, DT_F32, h->b_logits, nullptr, 0);
        LAUNCH(KC_DEC_GEMM_LOGITS, 1, gemm_bytes(gl, e), gemm_flops(gl), run_gemm(h, gl, st));
    }
    ArgmaxArgs aa{};
    aa.logits = lg; aa.partials = parts; aa.nparts = nparts; aa.B = rows; aa.V = c.vocab_size; aa.out_ids = h->out_ids.as<int64_t>() + (size_t)row0 * tcap; aa.out_ld = tcap;
    aa.cur_tok = ds.cur_tok + row0; aa.step = step; aa.seen_eos = ds.seen + row0; aa.done_step = ds.done_step + branch;
    aa.block_counter = ds.block_counter + branch; aa.eos = eos;
    aa.tok_emb = h->tok_emb; aa.pos_emb = h->pos_emb; aa.emb_dt = h->dt; aa.emb_max_pos = tcap;
    if (fuse) { aa.emb_x = sbuf; }
    else { aa.emb_x = xbuf; aa.emb_xn = xnbuf; aa.emb_g = h->dec_ln_g; aa.emb_b = h->dec_ln_b; }
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
    const bool fuse = h->fuse_ln && h->dt == DT_BF16 && h->use_tcgen05 && !(h->dbg_skip & 12);
    if (fuse) {
        LAUNCH(KC_DEC_ROW, 1, (double)rows * 256 * (8 + 4), 0.0,
               launch_embed_ln(ds.cur_tok + row0, ds.step + branch, 1, rows, h->tok_emb, h->pos_emb, c.vocab_size, nullptr, nullptr,
                               rowf(h->s, rc, 256), nullptr, h->dt, st));
    } else {
        LAUNCH(KC_DEC_ROW, 1, (double)rows * 256 * (8 + 4 + h->esz), 0.0,
               launch_embed_ln(ds.cur_tok + row0, ds.step + branch, 1, rows, h->tok_emb, h->pos_emb, c.vocab_size, h->dec_ln_g, h->dec_ln_b,
                               rowf(h->x, rc, 256), rowa(h, h->xn, rc, 256), h->dt, st));
    }
    return 0;
}

// k of the reference's top-k filter: int((1 - threshold) * vocab) in double arithmetic, as Python evaluates it (utils.py:87)
static int sampling_k(const texocr_handle* h) { return (int)((1.0 - h->samp_threshold) * (double)h->cfg.vocab_size); }

struct BranchPlan { int n; int row0[MAX_BRANCH]; int rows[MAX_BRANCH]; };
static BranchPlan plan_branches(texocr_handle* h, int B) {
    // ~86 rows per branch (6 branches at B = 512: measured 97.2 ms per generate vs 99.5 with 8 and 104.7 with 4), at most 8
    int n = h->decode_branches > 0 ? h->decode_branches : std::min(8, std::max(1, (B + 85) / 86));
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

// bf16 tier: the whole loop as launches of the cluster-persistent decode kernel (decode_mega.cu), `mega_steps` steps each.
static int run_generate_mega(texocr_handle* h, const DecState& ds, int eos, const int* d_enc_off, double sum_s, int B, int tcap,
                             int groups, int64_t* out_ids, int32_t* n_steps, cudaStream_t st) {
    const texocr_config& c = h->cfg;
    const int L = c.dec_layers;
    ENSURE(h->mega_part, (size_t)B * MEGA_CLUSTER * 8);
    for (int s2 = 0; s2 < 2; ++s2)
        if (!h->poll_ev[s2][0]) CK(cudaEventCreateWithFlags(&h->poll_ev[s2][0], cudaEventDisableTiming | cudaEventBlockingSync));
    MegaArgs a{};
    a.B = B; a.L = L; a.V = c.vocab_size; a.tcap = tcap; a.eos = eos; a.G = groups;
    for (int l = 0; l < L; ++l) {
        MegaLayerW& w = a.layer[l];
        w.wqkv = h->dec_self[l].wqkv; w.wo_s = h->dec_self[l].wo; w.bo_s = h->dec_self[l].bo;
        w.wq_c = h->dec_cross[l].wq; w.wo_c = h->dec_cross[l].wo; w.bo_c = h->dec_cross[l].bo;
        w.w1 = h->dec_mlp[l].w1; w.b1 = h->dec_mlp[l].b1; w.w2 = h->dec_mlp[l].w2; w.b2 = h->dec_mlp[l].b2;
    }
    a.w_logits = h->w_logits; a.b_logits = h->b_logits; a.tok_emb = h->tok_emb; a.pos_emb = h->pos_emb;
    a.ln_g = h->dec_ln_g; a.ln_b = h->dec_ln_b; a.fin_g = h->dec_norm_g; a.fin_b = h->dec_norm_b;
    a.cur_tok = ds.cur_tok; a.step = ds.step; a.done_step = ds.done_step; a.seen = ds.seen; a.out_ids = h->out_ids.as<int64_t>();
    a.x = h->x.as<float>(); a.s = h->s.as<float>(); a.qkv = h->qkv.p; a.o = h->o.p; a.hid = h->hid.p;
    a.part_val = h->mega_part.as<float>(); a.part_idx = h->mega_part.as<int>() + (size_t)B * MEGA_CLUSTER;
    a.kv_self = h->kvcache.p; a.kv_cross = h->crosskv_hm.p; a.ntok = h->crosskv_rows; a.enc_off = d_enc_off;
    static const bool timing = getenv("TEXOCR_MEGA_TIMING") != nullptr;      // debug: per-phase time breakdown on stderr
    if (timing) {
        ENSURE(h->mega_dbg, 16 * 8);
        CK(cudaMemsetAsync(h->mega_dbg.p, 0, 16 * 8, st));
        a.dbg_time = (unsigned long long*)h->mega_dbg.p;
    }
    int issued = 0, polls = 0;
    bool stop = false;
    while (issued < tcap && !stop) {
        a.nsteps = std::min(h->mega_steps, tcap - issued);
        // algorithmic traffic of the launch: every cached key / memory token is read once per step and layer (K + V, 8 heads)
        double keys = 0.0;
        for (int t = issued; t < issued + a.nsteps; ++t) keys += (double)B * (t + 1) + sum_s;
        const double flops_step = 2.0 * B * ((double)L * (256.0 * 1536 + 512.0 * 512 + 256.0 * 512 + 512.0 * 512 + 256.0 * 2048 + 1024.0 * 256) + 256.0 * c.vocab_size);
        LAUNCH(KC_DEC_MEGA, 1, keys * L * 1024 * h->esz, flops_step * a.nsteps + 4.0 * keys * L * 512, launch_decode_mega(a, st));
        issued += a.nsteps;
        if (eos >= 0 && issued < tcap) {      // host runs ahead of the device by at most two launches
            const int slot = polls & 1;
            if (polls >= 1) {
                CK(cudaEventSynchronize(h->poll_ev[slot ^ 1][0]));
                bool all = true;
                for (int i = 0; i < groups; ++i) all = all && h->h_poll[(slot ^ 1) * MAX_BRANCH + i] > 0;
                if (all) stop = true;
            }
            CK(cudaMemcpyAsync(&h->h_poll[slot * MAX_BRANCH], ds.done_step, (size_t)groups * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(h->poll_ev[slot][0], st));
            ++polls;
        }
    }
    CK(cudaMemcpyAsync(&h->h_poll[2 * MAX_BRANCH], ds.done_step, MAX_BRANCH * 4, cudaMemcpyDeviceToHost, st));
    int r;
    if ((r = from_device(h, out_ids, h->out_ids.p, (size_t)B * tcap * 8, st))) return r;
    // The host waits on blocking-sync events (the thread sleeps instead of spinning): with several batches in flight per GPU and
    // several ranks per box there are more waiting host threads than cores.
    if (!h->done_ev) CK(cudaEventCreateWithFlags(&h->done_ev, cudaEventDisableTiming | cudaEventBlockingSync));
    CK(cudaEventRecord(h->done_ev, st));
    CK(cudaEventSynchronize(h->done_ev));
    int done = 0;
    bool all_done = true;
    for (int i = 0; i < groups; ++i) {
        const int d = h->h_poll[2 * MAX_BRANCH + i];
        all_done = all_done && d > 0;
        done = std::max(done, d);
    }
    *n_steps = all_done ? done : tcap;
    if (timing) {
        unsigned long long tt[16];
        CK(cudaMemcpy(tt, h->mega_dbg.p, sizeof tt, cudaMemcpyDeviceToHost));
        static const char* nm[16] = {"qkv", "out_s", "q_c", "out_c", "ff1", "ff2", "logits", "self_attn", "cross_attn", "argmax", "cluster_sync", "kernel", "prologue", "operand_wait", "mma", "-"};
        const double ctas = (double)groups * MEGA_CLUSTER;
        fprintf(stderr, "[mega timing] B=%d steps=%d active_clusters=%d; per-CTA mean, us per step:", B, issued, decode_mega_active_clusters());
        for (int i = 0; i < 15; ++i) fprintf(stderr, " %s=%.1f", nm[i], tt[i] / ctas / issued * 1e-3);
        fprintf(stderr, "\n");
    }
    return 0;
}

// enc memory must already be projected into h->crosskv; d_enc_off = per-row token offsets (device, B+1)
static int run_generate(texocr_handle* h, const int64_t* d_start, int eos, const int* d_enc_off, int max_s, double sum_s, int B,
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
    {
        if (h->decode_mega && h->samp_temp <= 0.0 && h->dt == DT_BF16 && decode_mega_supported(B, c.dec_layers, c.vocab_size, nullptr)) {
            const int groups = decode_mega_groups(B);
            if (groups <= MAX_BRANCH) return run_generate_mega(h, ds, eos, d_enc_off, sum_s, B, tcap, groups, out_ids, n_steps, st);
        }
    }
    const BranchPlan bp = plan_branches(h, B);
    for (int i = 0; i < bp.n; ++i) {
        if (!h->branch_stream[i]) {
            // decode streams get the highest priority: their kernels are small and latency-bound, so with several batches in flight they
            // should be placed ahead of the wide encoder kernels of other handles, which then fill the gaps
            int pr_lo = 0, pr_hi = 0;
            CK(cudaDeviceGetStreamPriorityRange(&pr_lo, &pr_hi));
            CK(cudaStreamCreateWithPriority(&h->branch_stream[i], cudaStreamNonBlocking, h->decode_priority == 1 ? pr_hi : pr_lo));
            CK(cudaEventCreateWithFlags(&h->join_ev[i], cudaEventDisableTiming));
        }
    }
    h->own_stream2 = h->branch_stream[0];
    if (!h->fork_ev) CK(cudaEventCreateWithFlags(&h->fork_ev, cudaEventDisableTiming));

    const bool graph_ok = h->use_graph && !h->prof_on;
    // Stream of branch i: its own non-blocking stream (branch 0 included when there are several branches), so the
    // branches run as independent, phase-shifted pipelines that only meet again at the end of the call.
    cudaStream_t bst[MAX_BRANCH];
    // a single branch also moves to the engine's own (high-priority) decode stream, so that with several handles at work the small
    // decode kernels are placed ahead of other handles' wide encoder kernels
    const bool own = graph_ok && (bp.n > 1 || h->decode_priority);
    for (int i = 0; i < bp.n; ++i) bst[i] = !own ? st : (i == 0 ? h->own_stream2 : h->branch_stream[i]);
    const void* ckv_key = absorb ? h->dec_enc : h->crosskv_hm.p;
    const int samp_key = h->samp_temp > 0.0 ? sampling_k(h) : 0;      // seed and temperature are compared in full (gkey.samp_seed / samp_temp)
    const uint64_t seed_key = h->samp_temp > 0.0 ? h->samp_seed : 0;
    const double temp_key = h->samp_temp > 0.0 ? h->samp_temp : 0.0;
    // ---- coupled mode: one graph holds all branches of `spg` steps; attention launches are chained across branches (FIFO of
    // depth `fifo`), everything else of a branch only depends on the branch itself
    const int fifo = (graph_ok && bp.n > 1 && h->attn_fifo > 0 && 2 * c.dec_layers <= FIFO_STRIDE) ? std::min(h->attn_fifo, bp.n) : 0;
    if (fifo > 0) {
        cudaStream_t cs = h->own_stream2;
        const int spg = std::max(1, std::min(h->steps_per_graph, 16));
        const bool hit = h->cgraph_exec[0] && h->cgraph_exec[1] && h->gkey.B == B && h->gkey.tcap == tcap && h->gkey.eos == eos &&
                         h->gkey.max_s == max_s && h->gkey.samp == samp_key && h->gkey.samp_seed == seed_key && h->gkey.samp_temp == temp_key &&
                         h->gkey.enc_off == d_enc_off && h->gkey.ntok == h->crosskv_rows && h->gkey.kv == kv_key &&
                         h->gkey.ckv == ckv_key && h->gkey.x == h->x.p && h->gkey.nb == bp.n && h->gkey.fifo == fifo && h->gkey.spg == spg;
        if (!hit) {
            std::lock_guard<std::recursive_mutex> lk(g_dev_mu);
            drop_graphs(h);
            while (h->fifo_ev.size() < (size_t)spg * bp.n * FIFO_STRIDE) {
                cudaEvent_t ev = nullptr;
                CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                h->fifo_ev.push_back(ev);
            }
            for (int slot = 0; slot < 2; ++slot) {
                const int ns = slot == 0 ? 1 : spg;
                const int64_t before = h->launches;
                CK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed));
                cudaError_t ce = cudaEventRecord(h->fork_ev, cs);
                for (int i = 1; i < bp.n && ce == cudaSuccess; ++i) ce = cudaStreamWaitEvent(h->branch_stream[i], h->fork_ev, 0);
                r = 0;
                for (int s2 = 0; s2 < ns && !r && ce == cudaSuccess; ++s2)
                    for (int i = 0; i < bp.n && !r; ++i) {
                        const int j = s2 * bp.n + i;                 // position in the FIFO order (step-major, then branch)
                        h->fifo_rec = h->fifo_ev.data() + (size_t)j * FIFO_STRIDE;
                        h->fifo_wait = j >= fifo ? h->fifo_ev.data() + (size_t)(j - fifo) * FIFO_STRIDE : nullptr;
                        r = enqueue_decode_step(h, B, bp.row0[i], bp.rows[i], i, tcap, eos, d_enc_off, max_s, sum_s, -1,
                                                i == 0 ? cs : h->branch_stream[i]);
                    }
                h->fifo_rec = nullptr; h->fifo_wait = nullptr;
                for (int i = 1; i < bp.n && ce == cudaSuccess && !r; ++i) {
                    ce = cudaEventRecord(h->join_ev[i], h->branch_stream[i]);
                    if (ce == cudaSuccess) ce = cudaStreamWaitEvent(cs, h->join_ev[i], 0);
                }
                cudaError_t ee = cudaStreamEndCapture(cs, &h->cgraph[slot]);
                if (r) return r;
                CK(ce);
                CK(ee);
                CK(cudaGraphInstantiate(&h->cgraph_exec[slot], h->cgraph[slot], 0));
                h->cgraph_kernels[slot] = (int)(h->launches - before);
                h->launches = before;
            }
            h->gkey.samp = samp_key; h->gkey.samp_seed = seed_key; h->gkey.samp_temp = temp_key; h->gkey.enc_off = d_enc_off;
            h->gkey.B = B; h->gkey.tcap = tcap; h->gkey.eos = eos; h->gkey.max_s = max_s; h->gkey.nb = bp.n;
            h->gkey.kv = const_cast<void*>(kv_key); h->gkey.ckv = const_cast<void*>(ckv_key); h->gkey.x = h->x.p; h->gkey.ntok = h->crosskv_rows;
            h->gkey.fifo = fifo; h->gkey.spg = spg;
        }
        CK(cudaEventRecord(h->fork_ev, st));
        CK(cudaStreamWaitEvent(cs, h->fork_ev, 0));
        for (int s2 = 0; s2 < 2; ++s2)
            if (!h->poll_ev[s2][0]) CK(cudaEventCreateWithFlags(&h->poll_ev[s2][0], cudaEventDisableTiming | cudaEventBlockingSync));
        for (int i = 0; i < bp.n; ++i)
            if ((r = enqueue_first_embed(h, B, bp.row0[i], bp.rows[i], i, cs))) return r;
        const int POLL = 16;
        int polls = 0, next_poll = POLL;
        bool stop = false;
        for (int t = 0; t < max_len && !stop;) {
            const int slot = (max_len - t >= spg && spg > 1) ? 1 : 0;
            CK(cudaGraphLaunch(h->cgraph_exec[slot], cs));
            h->launches += h->cgraph_kernels[slot];
            t += slot ? spg : 1;
            if (eos >= 0 && t >= next_poll && t < max_len) {
                next_poll += POLL;
                const int ps = polls & 1;
                if (polls >= 1) {
                    CK(cudaEventSynchronize(h->poll_ev[ps ^ 1][0]));
                    bool all = true;
                    for (int i = 0; i < bp.n; ++i) all = all && h->h_poll[(ps ^ 1) * MAX_BRANCH + i] > 0;
                    if (all) stop = true;
                }
                CK(cudaMemcpyAsync(&h->h_poll[ps * MAX_BRANCH], ds.done_step, (size_t)bp.n * 4, cudaMemcpyDeviceToHost, cs));
                CK(cudaEventRecord(h->poll_ev[ps][0], cs));
                ++polls;
            }
        }
        CK(cudaEventRecord(h->join_ev[0], cs));
        CK(cudaStreamWaitEvent(st, h->join_ev[0], 0));
    }
    if (fifo > 0) {
    } else if (graph_ok) {
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
    for (int i = 0; i < bp.n && fifo == 0; ++i)
        if ((r = enqueue_first_embed(h, B, bp.row0[i], bp.rows[i], i, bst[i]))) return r;
    // Host runs ahead of the device by at most 2*POLL steps; an early exit costs at most that many extra steps.
    const int POLL = 16;
    int issued = 0, polls = 0;
    bool stop = fifo > 0;
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
    if (own && fifo == 0) {      // join
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
                      &h->row_loss, &h->scalars, &h->dec_state, &h->out_ids, &h->mega_part, &h->mega_dbg, &h->attn_trace, &h->amax_part, &h->qabs, &h->cabs, &h->latcache, &h->prep_meta, &h->prep_in, &h->prep_out};
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
    if (e != cudaSuccess) return fail_cuda(h, e, "debug_sample_step", __LINE__);
    return 0;
}

int texocr_set_option(texocr_handle* h, const char* name, int64_t value) {
    if (!h || !name) return TEXOCR_ERR_ARG;
    if (!strcmp(name, "cuda_graph")) { h->use_graph = value != 0; return 0; }
    if (!strcmp(name, "stagger_us")) { h->stagger_us = (int)value; return 0; }
    if (!strcmp(name, "attn_l2_policy")) { g_attn_l2_policy = (int)value; drop_graphs(h); return 0; }
    if (!strcmp(name, "attn_abs_minb")) { g_attn_abs_minb = (int)value; drop_graphs(h); return 0; }
    if (!strcmp(name, "decode_priority")) { h->decode_priority = (int)value; return 0; }      // before the first generate call
    if (!strcmp(name, "absorb_two_stage")) { h->absorb_two_stage = (int)value; drop_graphs(h); return 0; }
    if (!strcmp(name, "self_absorb")) { h->self_absorb = (int)value; drop_graphs(h); return 0; }
    if (!strcmp(name, "cross_absorb")) { h->cross_absorb = (int)value; drop_graphs(h); return 0; }
    if (!strcmp(name, "attn_trace")) { h->attn_trace_on = value != 0; drop_graphs(h); return 0; }
    if (!strcmp(name, "attn_fifo")) { h->attn_fifo = (int)std::max<int64_t>(0, std::min<int64_t>(MAX_BRANCH, value)); return 0; }
    if (!strcmp(name, "steps_per_graph")) { h->steps_per_graph = (int)std::max<int64_t>(1, std::min<int64_t>(16, value)); return 0; }
    if (!strcmp(name, "fifo_pdl")) { h->fifo_pdl = value != 0; drop_graphs(h); return 0; }
    if (!strcmp(name, "attn_full_tail")) { g_attn_full_tail = (int)value; drop_graphs(h); return 0; }
    if (!strcmp(name, "fuse_ln")) { h->fuse_ln = value != 0; drop_graphs(h); return 0; }
    if (!strcmp(name, "no_early_exit")) { h->no_early_exit = value != 0; return 0; }
    if (!strcmp(name, "keep_logits")) { h->keep_logits = value != 0; drop_graphs(h); return 0; }
    if (!strcmp(name, "poison")) { h->poison = value != 0; return 0; }
    if (!strcmp(name, "attn_ctas_per_sm")) { h->attn_ctas_per_sm = (int)std::max<int64_t>(1, std::min<int64_t>(8, value)); drop_graphs(h); return 0; }
    if (!strcmp(name, "dbg_skip")) { h->dbg_skip = (int)value; drop_graphs(h); return 0; }
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
    if (!strcmp(name, "gemm_shallow_ring")) { g_tc_shallow_ring = (int)value; drop_graphs(h); return 0; }
    if (!strcmp(name, "gemm_deep_ring")) { g_tc_deep_ring = (int)value; drop_graphs(h); return 0; }
    if (!strcmp(name, "gemm_min_ctas")) { g_tc_min_ctas = (int)value; drop_graphs(h); return 0; }
    if (!strcmp(name, "gemm_persist_min_tiles")) { g_tc_persist_min_tiles = (int)value; drop_graphs(h); return 0; }
    if (!strcmp(name, "gemm_tiles_per_cta")) { g_tc_tiles_per_cta = (int)value; drop_graphs(h); return 0; }
    if (!strcmp(name, "gemm_persistent_stages")) { g_tc_persistent_stages = (int)value; return 0; }
    if (!strcmp(name, "gemm_split_k")) { g_tc_split_k = (int)value; drop_graphs(h); return 0; }
    if (!strcmp(name, "im2col_tma")) { h->use_im2col_tma = value != 0; return 0; }
    if (!strcmp(name, "decode_mega")) { h->decode_mega = (int)value; return 0; }
    if (!strcmp(name, "mega_steps")) { h->mega_steps = (int)std::max<int64_t>(1, std::min<int64_t>(4096, value)); return 0; }
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
    CK(launch_attn_abs(ab, h->num_sms * h->attn_ctas_per_sm, st));
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
