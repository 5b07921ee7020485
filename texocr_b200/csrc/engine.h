// Engine state behind the C-ABI handle (include/texocr.h).  Host-side only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/texocr.h"
#include "kernels.h"

struct HostTensor {
    std::vector<float> data;
    std::vector<int64_t> shape;
    int64_t numel() const { int64_t n = 1; for (auto s : shape) n *= s; return n; }
};

struct DevBuf {          // grow-only device buffer
    void* p = nullptr;
    size_t bytes = 0;
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct ConvW {           // one weight-standardised convolution + its GroupNorm
    std::string name, gn;
    int cin, cout, k, stride, act;
    float* w = nullptr;          // [cout][k*k*cin] fp32, (ky,kx,c) order, standardised
    void* w_hi = nullptr; void* w_lo = nullptr;   // bf16 tier: split-bf16 pair of the same matrix (tcgen05 bf16x3 path)
    float* gamma = nullptr; float* beta = nullptr;
};

struct AttnW { void* wqkv = nullptr; void* wq = nullptr; void* wo = nullptr; float* bo = nullptr;
               void* wqk = nullptr; void* wvo = nullptr; void* wv = nullptr; };    // cross, bf16 tier: Wq^T.Wk per head [2048,256] and Wo.Wv per head [512,2048] (absorbed K / V projections)
struct MlpW { void* w1 = nullptr; float* b1 = nullptr; void* w2 = nullptr; float* b2 = nullptr; };

enum KClass {
    KC_STEM = 0, KC_GN_STATS, KC_GN_APPLY, KC_CONV, KC_ENC_GEMM, KC_ENC_ATTN, KC_ENC_ROW, KC_CROSSKV_GEMM,
    KC_DEC_GEMM, KC_DEC_ATTN_SELF, KC_DEC_ATTN_CROSS, KC_DEC_ROW, KC_DEC_ARGMAX, KC_TF_GEMM, KC_TF_ATTN, KC_TF_ROW,
    KC_MISC, KC_UNUSED,
    // the decode-step GEMMs by role (bench.py sums them into "dec_gemm"; KC_DEC_GEMM itself = the projected-K/V formulation's QKV / q GEMMs)
    KC_DEC_GEMM_Q, KC_DEC_GEMM_VPROJ, KC_DEC_GEMM_WO, KC_DEC_GEMM_W1, KC_DEC_GEMM_W2, KC_DEC_GEMM_LOGITS,
    KC_COUNT
};

struct ProfRec { int cls; cudaEvent_t e0, e1; double bytes, flops; };

struct texocr_handle {
    texocr_config cfg{};
    int device = 0;
    int dt = DT_F32;                 // element type of transformer GEMM operands / KV cache
    size_t esz = 4;
    std::string err;
    std::map<std::string, HostTensor> sd;
    bool finalized = false;
    std::vector<void*> weight_allocs;

    // ---- packed weights
    float* stem_w = nullptr; float* stem_g = nullptr; float* stem_b = nullptr;
    void* stem_w_hi = nullptr; void* stem_w_lo = nullptr;      // bf16 tier: [64][64] split-bf16 filter bank of the tensor-core stem (49 taps + zero padding)
    std::vector<ConvW> convs;                 // the 39 non-stem convolutions in execution order
    void* proj_w = nullptr; void* proj_w_lo = nullptr; float* proj_b = nullptr; int proj_k = 0;
    float* cls = nullptr; float* pos = nullptr;
    float* enc_ln_g = nullptr; float* enc_ln_b = nullptr; float* enc_norm_g = nullptr; float* enc_norm_b = nullptr;
    std::vector<AttnW> enc_attn; std::vector<MlpW> enc_mlp;
    float* tok_emb = nullptr; float* pos_emb = nullptr;
    float* dec_ln_g = nullptr; float* dec_ln_b = nullptr; float* dec_norm_g = nullptr; float* dec_norm_b = nullptr;
    std::vector<AttnW> dec_self, dec_cross; std::vector<MlpW> dec_mlp;
    void* w_crosskv = nullptr;                // [L*1024][256]: per layer K rows then V rows
    void* w_logits = nullptr; float* b_logits = nullptr;

    // ---- workspaces (grow-only)
    DevBuf geom;                               // int32: img_off[B+1] | img_hw[2B] | tok_off[B+1] | row_off[B+1]
    int* h_geom = nullptr; size_t h_geom_cap = 0;   // pinned staging for geom
    cudaEvent_t done_ev = nullptr;             // blocking-sync event the host waits on at the end of a generate call
    cudaEvent_t geom_ev = nullptr, hop_in = nullptr, hop_out = nullptr;
    cudaStream_t own_stream = nullptr, own_stream2 = nullptr;
    cudaStream_t branch_stream[16] = {nullptr}; cudaEvent_t join_ev[16] = {nullptr}; cudaEvent_t fork_ev = nullptr;
    int decode_branches = 0;                   // 0 = automatic (one branch per 128 rows = one GEMM row tile, at most 8)
    DevBuf img_stage;                          // device copy of host images
    DevBuf raw1, act2, actA, actB, rawMid, actMid, rawMid2, actMid2, raw3, rawDs;
    DevBuf gn_partial, gn_stats[4];
    DevBuf gn_part;                            // bf16 tier: per-32-row-block GroupNorm partial sums (gn_block.cuh)
    DevBuf proj_out, patch_cols, backbone_a, col;
    DevBuf x, s, xn, qkv, o, hid, logits;
    DevBuf amax_part;                          // decode step, bf16 tier greedy: [B][ceil(V/32)] {max, index} partials of the vocabulary GEMM (no logits in HBM)
    DevBuf qabs, cabs;                         // absorbed cross-attention: queries / memory averages [rows, 8 x 256]
    DevBuf enc_out, enc_a, crosskv, crosskv_hm, kvcache;      // crosskv: GEMM output [tok][L*1024]; crosskv_hm: head-major copy for decoding
    DevBuf ids_stage, mask_stage, enc_stage, tgt_stage, row_loss, scalars;
    DevBuf dec_state;                          // int64 cur_tok[B] | int32 step, done_step, block_counter, pad | int32 seen[B]
    DevBuf out_ids;                            // int64 [B, max_len]
    DevBuf attn_trace; bool attn_trace_on = false;   // debug timeline of the decode attention launches: [branch][3][256 steps][8] u64 ns
    DevBuf prep_meta, prep_in, prep_out;       // texocr_preprocess_u8 staging
    int* h_poll = nullptr;                     // pinned: done_step polls
    int64_t* h_bos = nullptr; size_t h_bos_cap = 0;   // pinned: the BOS start column of texocr_generate (constant content)
    int last_backbone_pixels = 0;
    int crosskv_rows = 0;

    // ---- decode-step CUDA graph
    bool use_graph = true;
    cudaGraph_t graph = nullptr; cudaGraphExec_t graph_exec = nullptr;        // branch 0 (kept under these names)
    cudaGraph_t bgraph[16] = {nullptr}; cudaGraphExec_t bgraph_exec[16] = {nullptr};   // one single-step graph per branch
    cudaEvent_t poll_ev[2][16] = {{nullptr}};
    int stagger_us = 30;                        // start offset between consecutive branches
    struct { int B = 0, tcap = 0, eos = 0, max_s = 0; void* kv = nullptr; void* ckv = nullptr; void* x = nullptr; int kernels = 0; int nb = 0; int ntok = 0; int samp = 0;
             const int* enc_off = nullptr; uint64_t samp_seed = 0; double samp_temp = 0.0; } gkey;   // enc_off: the cross-attention launches bake this device pointer in

    // ---- instrumentation
    int64_t launches = 0;
    bool prof_on = false;
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> ev_pool;
    double prof_ms[KC_COUNT] = {0}; double prof_bytes[KC_COUNT] = {0}; double prof_flops[KC_COUNT] = {0};
    int64_t prof_n[KC_COUNT] = {0};
    bool use_tcgen05 = true;
    int gn_fused = 1;             // bf16 tier: GroupNorm partial sums in the epilogue of the producing convolution GEMM (same-size batches whose
                                  // images have a multiple of 32 pixel rows at every level); 0 = always the stand-alone block kernel
    bool use_stem_tc = true;      // bf16 tier: stem convolution on the tensor cores (tc_stem_kernel) instead of the FFMA kernel
    bool use_conv_gather = true;  // bf16 tier, ragged batches: 3x3 / strided convolutions as implicit GEMMs with cp.async-gathered A tiles
    bool use_im2col_tma = true;   // bf16 tier, same-size batches: 3x3 / strided convolutions as implicit GEMMs (TMA im2col loads)
    // bf16 tier generate loop: cross-attention streams the [S,256] encoder memory once for all heads instead of per-head K/V
    // (K / V projections folded into the query / output projections).  0 = projected K/V cache.
    int cross_absorb = 1;
    int absorb_two_stage = 1;                  // out-projection of the absorbed attention as per-head value projection (block-diagonal GEMM, K = 256)
                                               // + the ordinary Wo GEMM (K = 512) instead of one folded K = 2048 GEMM
    int self_absorb = 1;                       // same for the self-attention: the cache holds the layer's 256-wide LayerNorm'd inputs
    bool self_abs_active = false;              // decided per generate call (run_crosskv)
    DevBuf latcache;                           // [layer][sequence][position][256] bf16 latent cache of the absorbed self-attention
    const void* dec_enc = nullptr;             // bf16 encoder memory of the current generate call [crosskv_rows, 256]
    int use_tma_attn = 1;     // 0 = simple kernel, 1 = TMA kernel for self + cross, 2 = self only, 3 = cross only
    bool no_early_exit = false; // texocr_generate runs all max_len steps even when every row has produced an EOS (sub-batches of one large call)
    bool keep_logits = false;   // debug / tests: the decode step also leaves its last-position logits in h->logits (texocr_debug_read "logits")
    bool poison = false;     // debug: NaN-fill all workspaces at the start of texocr_generate
    int dbg_skip = 0;        // timing experiments only: 1 self-attn, 2 cross-attn, 4 LayerNorms, 8 GEMMs (results are garbage)
    // sampling (texocr_set_sampling): temp <= 0 = greedy
    double samp_temp = 0.0, samp_threshold = 0.9; uint64_t samp_seed = 0; uint32_t samp_calls = 0;
    int num_sms = 148;
    int attn_ctas_per_sm = 4;     // persistent decode-attention CTAs per SM (64 KB ring each); leaves room for the GEMM CTAs of other branches
};
