// Internals shared by the engine's translation units (engine_core / _weights / _encoder / _decode / _api).  Host side only.
#pragma once
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <mutex>

#include "engine.h"
#include "tc_gemm.h"

extern int g_tc_persistent;
extern int g_tc_persistent_stages;
extern int g_tc_min_ctas;
extern int g_tc_epi_warps;
extern int g_tc_bn256;
extern int g_attn_full_tail;
extern int g_attn_abs_minb;

// Several handles may be driven from different host threads (texocr_b200/pipeline.py).  Stream capture and device-wide
// operations do not mix across threads (a cudaDeviceSynchronize / cudaFree in one thread invalidates a capture in
// another), so graph capture and (re)allocation take this process-wide lock.  Steady-state calls never hold it.
extern std::recursive_mutex g_dev_mu;
extern std::string g_create_error;
extern const char* kclass_name[KC_COUNT];

int fail(texocr_handle* h, int code, const char* fmt, ...);
int fail_cuda(texocr_handle* h, cudaError_t e, const char* what, int line, const char* file);
cudaEvent_t get_event(texocr_handle* h);

#define CK(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess) return fail_cuda(h, e__, #expr, __LINE__, __FILE__);               \
    } while (0)

// launch + accounting: kernel count of the handle, and (profiling on) CUDA events around the launch on its stream `st`
#define LAUNCH(kc_, nkern, bytes_, flops_, expr)                                                   \
    do {                                                                                           \
        ProfRec pr__;                                                                              \
        if (h->prof_on) { pr__.cls = (kc_); pr__.bytes = (bytes_); pr__.flops = (flops_);          \
            pr__.e0 = get_event(h); pr__.e1 = get_event(h); cudaEventRecord(pr__.e0, st); }        \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess) return fail_cuda(h, e__, #expr, __LINE__, __FILE__);               \
        h->launches += (nkern);                                                                    \
        if (h->prof_on) { cudaEventRecord(pr__.e1, st); h->prof.push_back(pr__); }                 \
    } while (0)

// ---- memory (engine_core.cu)
void drop_graphs(texocr_handle* h);
int ensure(texocr_handle* h, DevBuf& b, size_t bytes);
#define ENSURE(buf, bytes) do { int r__ = ensure(h, (buf), (bytes)); if (r__) return r__; } while (0)
bool is_device_ptr(const void* p);
int to_device(texocr_handle* h, const void* p, size_t bytes, DevBuf& stage, const void** out, cudaStream_t st);
int from_device(texocr_handle* h, void* dst, const void* src, size_t bytes, cudaStream_t st);
int upload_ints(texocr_handle* h, const std::vector<int>& v, cudaStream_t st);
int poison_workspaces(texocr_handle* h, cudaStream_t st);

// ---- GEMM dispatch (engine_core.cu)
cudaError_t run_gemm(texocr_handle* h, const GemmArgs& g, cudaStream_t st);
GemmArgs mk_gemm(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K, int epi,
                 int dt_a, int dt_c, const float* bias, const float* res, int ldres);
double gemm_bytes(const GemmArgs& g, size_t esz);
double gemm_flops(const GemmArgs& g);

// ---- weights (engine_weights.cu)
int finalize_weights(texocr_handle* h);
void fold_absorbed(const HostTensor& q, const HostTensor& k, const HostTensor& v, const HostTensor& wo, std::vector<float>& wqk,
                   std::vector<float>& wvoi);

// ---- encoder (engine_encoder.cu)
struct EncGeom {
    int B = 0;
    std::vector<int> img_off, tok_off;
    long P[5] = {0, 0, 0, 0, 0};
    int ntok = 0, max_tok = 0;
    int uni_h = 0, uni_w = 0;          // > 0: every image has this size (enables the TMA im2col convolutions)
    bool rows32[5] = {true, true, true, true, true};      // level L: every image has a multiple of 32 pixel rows (GroupNorm partials in the GEMM epilogue)
    const int* d_img_off = nullptr; const int* d_img_hw = nullptr; const int* d_tok_off = nullptr;
};
int plan_geometry(texocr_handle* h, const int32_t* hw, int B, EncGeom& g, cudaStream_t st);
int run_encoder(texocr_handle* h, const float* d_img, const EncGeom& g, cudaStream_t st);
int run_crosskv(texocr_handle* h, const float* enc_f32, const void* enc_typed, int ntok, cudaStream_t st, bool for_generate = false);
// One (self-attention | cross-attention | MLP) sub-layer tail shared by encoder / decoder / decode step.
struct RowCtx {
    int rows; int kc_gemm, kc_row;
    const float* ln_g; const float* ln_b;
    int row0 = 0;        // first row of this sub-batch inside the row workspaces
};
static inline float* rowf(const DevBuf& b, const RowCtx& rc, int width) { return b.as<float>() + (size_t)rc.row0 * width; }
static inline void* rowa(texocr_handle* h, const DevBuf& b, const RowCtx& rc, int width) { return (char*)b.p + (size_t)rc.row0 * width * h->esz; }
int sub_attn_out(texocr_handle* h, const RowCtx& rc, const AttnW& w, cudaStream_t st);
int sub_mlp(texocr_handle* h, const RowCtx& rc, const MlpW& w, cudaStream_t st);
int sub_norm(texocr_handle* h, const RowCtx& rc, bool last, const float* fin_g, const float* fin_b, float* fin_out_f,
             void* fin_out_a, cudaStream_t st);
int sub_abs_out(texocr_handle* h, const RowCtx& rc, const AttnW& w, const void* ca, cudaStream_t st);
int ensure_rows(texocr_handle* h, long rows);

// ---- decode (engine_decode.cu)
constexpr int MAX_BRANCH = 16;
int sampling_k(const texocr_handle* h);
int run_generate(texocr_handle* h, const int64_t* d_start, int eos, const int* d_enc_off, int max_s, double sum_s, int B,
                 int max_len, int64_t* out_ids, int32_t* n_steps, cudaStream_t st);
