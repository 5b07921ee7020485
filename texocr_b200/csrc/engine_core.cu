// Engine core: error reporting, launch accounting, grow-only device buffers / staging, GEMM dispatch (tcgen05 or FFMA).
#include "engine_internal.h"

std::string g_create_error;
std::recursive_mutex g_dev_mu;
// bits 0..5: programmatic dependent launch per kernel family (kernels.h); bit 8 / 9: LayerNorm / GEMM kernels release their
// dependents only after their stores (experiment switches; the default is an early trigger everywhere).
int g_texocr_pdl = 0x3f;
// where a kernel releases its programmatic dependent: bit set = late in the kernel (1.5-2 us before its end: the dependent's CTAs are
// scheduled when their launch + prologue just fits, instead of sitting in SM slots for the kernel's whole run); 0 = at entry.
// 1 = tcgen05 GEMM (accumulator complete), 2 = decode attention (warp's last chunk issued), 4 = LayerNorm (row loaded), 8 = token kernel
int g_texocr_pdl_mid = 7;

int fail(texocr_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}
int fail_cuda(texocr_handle* h, cudaError_t e, const char* what, int line, const char* file) {
    const char* base = strrchr(file, '/');
    return fail(h, TEXOCR_ERR_CUDA, "CUDA error %s (%s) at %s:%d: %s", cudaGetErrorName(e), cudaGetErrorString(e), base ? base + 1 : file, line, what);
}

// ------------------------------------------------------------------------------------------------ profiling / launch accounting
const char* kclass_name[KC_COUNT] = {
    "stem_conv", "gn_stats", "gn_apply", "conv_gemm", "enc_gemm", "enc_attn", "enc_rowwise", "crosskv_gemm",
    "dec_gemm", "dec_attn_self", "dec_attn_cross", "dec_rowwise", "dec_argmax", "tf_gemm", "tf_attn", "tf_rowwise", "misc", "unused",
    "dec_gemm_q", "dec_gemm_vproj", "dec_gemm_wo", "dec_gemm_w1", "dec_gemm_w2", "dec_gemm_logits"};

cudaEvent_t get_event(texocr_handle* h) {
    if (!h->ev_pool.empty()) { cudaEvent_t e = h->ev_pool.back(); h->ev_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}



// ------------------------------------------------------------------------------------------------ memory helpers
void drop_graphs(texocr_handle* h) {
    for (int i = 0; i < 16; ++i) {
        if (h->bgraph_exec[i]) { cudaGraphExecDestroy(h->bgraph_exec[i]); h->bgraph_exec[i] = nullptr; }
        if (h->bgraph[i]) { cudaGraphDestroy(h->bgraph[i]); h->bgraph[i] = nullptr; }
    }
    h->graph_exec = nullptr; h->graph = nullptr;
}

int ensure(texocr_handle* h, DevBuf& b, size_t bytes) {
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

bool is_device_ptr(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// Return a device pointer for `p` (bytes long): p itself if it is device memory, else a staged copy.
int to_device(texocr_handle* h, const void* p, size_t bytes, DevBuf& stage, const void** out, cudaStream_t st) {
    if (is_device_ptr(p)) { *out = p; return 0; }
    ENSURE(stage, bytes);
    CK(cudaMemcpyAsync(stage.p, p, bytes, cudaMemcpyHostToDevice, st));
    *out = stage.p;
    return 0;
}
int from_device(texocr_handle* h, void* dst, const void* src, size_t bytes, cudaStream_t st) {
    if (dst == src) return 0;
    CK(cudaMemcpyAsync(dst, src, bytes, is_device_ptr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    return 0;
}

int upload_ints(texocr_handle* h, const std::vector<int>& v, cudaStream_t st) {
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
int poison_workspaces(texocr_handle* h, cudaStream_t st) {
    DevBuf* bufs[] = {&h->raw1, &h->act2, &h->actA, &h->actB, &h->rawMid, &h->actMid, &h->rawMid2, &h->actMid2, &h->raw3, &h->rawDs,
                      &h->gn_partial, &h->gn_stats[0], &h->gn_stats[1], &h->gn_stats[2], &h->gn_stats[3], &h->proj_out, &h->patch_cols,
                      &h->backbone_a, &h->col, &h->x, &h->s, &h->xn, &h->qkv, &h->o, &h->hid, &h->logits, &h->enc_out, &h->enc_a,
                      &h->crosskv, &h->crosskv_hm, &h->kvcache, &h->out_ids};
    for (DevBuf* b : bufs) if (b->p) CK(cudaMemsetAsync(b->p, 0xFF, b->bytes, st));
    return 0;
}

// ------------------------------------------------------------------------------------------------ GEMM dispatch
// tcgen05 path for bf16 operands when the shape fits its tiles, FFMA path otherwise (and always in the fp32 tier).
cudaError_t run_gemm(texocr_handle* h, const GemmArgs& g, cudaStream_t st) {
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
GemmArgs mk_gemm(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K, int epi,
                        int dt_a, int dt_c, const float* bias, const float* res, int ldres) {
    GemmArgs g{};
    g.A = A; g.W = W; g.C = C; g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldw = ldw; g.ldc = ldc;
    g.bias = bias; g.res = res; g.ldres = ldres; g.dt_a = dt_a; g.dt_c = dt_c; g.epi = epi; g.conv = nullptr;
    return g;
}
// algorithmic bytes of a GEMM launch: both operands once + what the epilogue reads / writes
double gemm_bytes(const GemmArgs& g, size_t esz) {
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
double gemm_flops(const GemmArgs& g) { return 2.0 * g.M * (double)g.N * g.K; }
