// Launch wrappers of the TeXOCR kernels (host side).  Every wrapper enqueues on `st` and
// returns the cudaError_t of the launch.  DT_* select the element type of a buffer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

enum { DT_F32 = 0, DT_BF16 = 1 };

// Launch with programmatic dependent launch enabled (kernels call pdl_wait() before reading their inputs).
// Programmatic dependent launch per kernel family: bit k of g_texocr_pdl enables it for family k (engine_core.cu).
enum { PDL_GEMM = 0, PDL_LN = 1, PDL_EMBED = 2, PDL_ARGMAX = 3, PDL_ATTN_TMA = 4, PDL_ATTN_SIMPLE = 5 };
extern int g_texocr_pdl;
extern int g_texocr_pdl_mid;      // engine_core.cu
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int family, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = ((g_texocr_pdl >> family) & 1) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#endif

enum GemmEpi {
    EPI_STORE = 0,      // C = A.W^T (+bias)                       out type = dt_c
    EPI_GLU_RES = 1,    // pairs (2j,2j+1): (a+ba)*sigmoid(g+bg) + res    -> float [M, N/2]
    EPI_GEGLU = 2,      // pairs (2j,2j+1): (a+ba)*gelu_erf(g+bg)         -> dt_a  [M, N/2]
    EPI_BIAS_RES = 3,   // acc + bias + res                                -> float [M, N]
    // tcgen05 path only, 32-column tiles: per row the (max, lowest index of the max) of acc + bias over the tile's columns
    // -> {float max, int index} at C[row * ldc + tile] (8 bytes each; ldc = number of 32-column tiles).  The vocabulary
    // projection of the decode step: the logits themselves are never stored (model/decoder.py:60,103)
    EPI_ARGMAX = 4
};

struct ConvGather {      // implicit-GEMM view of a convolution over a ragged NHWC pixel batch
    const int* img_off;  // [B+1] full-resolution pixel offsets
    const int* img_hw;   // [B][2]
    int nimg;
    int lin, lout;       // input / output level (pixels are H>>L x W>>L)
    int ksz, stride, pad, cin;
};

struct GemmArgs {
    const void* A; const void* W; void* C;
    int M, N, K;
    int lda, ldw, ldc;
    const float* bias;
    const float* res; int ldres;
    int dt_a;            // type of A and W
    int dt_c;            // type of C for EPI_STORE
    int epi;
    const ConvGather* conv;   // non-null: A(m,k) gathered from NHWC input (fp32 only)
    const void* A2; const void* W2;   // tcgen05 bf16x3 mode: low-order bf16 parts of A and W (same shapes / strides)
    // tcgen05 path, implicit GEMM of a ksz x ksz convolution over a batch of N same-size NHWC images [N][H][W][C] (A = the
    // activation, K = ksz*ksz*C, k order (ky, kx, c)): the A tiles are fetched with TMA im2col loads, no im2col buffer.
    // pad_lo = padding before (TF-SAME, utils.py:93-123); M = N * Ho * Wo output pixels.  ksz == 0: plain GEMM.
    struct Im2col { int ksz, stride, pad_lo, C, W, H, N, Wo, Ho; } im2col;
    // tcgen05 path, bf16x3, ragged batch: implicit GEMM of the convolution `gather` describes over the split-bf16 NHWC activation
    // (A, A2) = (hi, lo) -- the A tiles are gathered row by row with cp.async (tc_conv_gather_kernel), no im2col buffer.  null = off
    const ConvGather* gather;
    // tcgen05 path, block-diagonal GEMM: output columns [64j, 64j+64) contract A[:, j*a_block_k .. j*a_block_k + K) with W rows
    // [64j, 64j+64) (W is [N, K]); A has N/64 * a_block_k columns.  0 = ordinary GEMM.  (per-head value projection of the absorbed
    // attention: C_h [256] -> 64 values with Wv_h)
    int a_block_k;
    // tcgen05 path, EPI_STORE / fp32 output of a convolution over same-size images of gn_rpi pixel rows each (gn_rpi % 32 == 0):
    // the epilogue also leaves the GroupNorm(32) partial sums of every 32-row block, gn_part[(row >> 5) + image][32][2]
    // (gn_block.cuh; summed per image by launch_gn_finalize_blocks).  null = off.  gn_rpi == 0: images of different sizes, every one a
    // multiple of 32 rows at this level (gn_img_off / gn_nimg / gn_level: the offset table the image of a row block is looked up in)
    float* gn_part; int gn_rpi; const int* gn_img_off; int gn_nimg, gn_level;
    // debug (engine option attn_trace): per-CTA residency sums of the launch: [0] ns waiting for the predecessor grid,
    // [1] ns from there to CTA exit, [2] CTAs.  null = off
    unsigned long long* dbg;
};
cudaError_t launch_gemm_simt(const GemmArgs& g, cudaStream_t st);

// ---- backbone pieces (conv_gn.cu)
cudaError_t launch_stem_conv(const float* img, const float* w /*[49][64] std*/, float* raw1, const int* img_off,
                             const int* img_hw, int nimg, int total_p1, cudaStream_t st);
cudaError_t launch_gn_stats(const float* raw, int C, int level, const int* img_off, int nimg, int nchunk,
                            double* partial, float* stats, cudaStream_t st);
// bf16 tier: block-partial statistics (gn_block.cuh).  launch_gn_stats_blocks is the stand-alone producer of the partials
// (ragged batches, or images whose row count is not a multiple of 32); launch_gn_finalize_blocks turns an image's partials
// into stats[b][32] = (mean, rstd).  part holds ((total rows >> 5) + nimg + 1) * 64 floats.
cudaError_t launch_gn_stats_blocks(const float* raw, int C, int level, const int* img_off, int nimg, long total_rows, float* part,
                                   cudaStream_t st);
cudaError_t launch_gn_finalize_blocks(const float* part, int C, int level, const int* img_off, int nimg, float* stats, cudaStream_t st);
struct GnApplyArgs {
    const float* raw; const float* stats; const float* gamma; const float* beta;       // main input
    const float* raw2; const float* stats2; const float* gamma2; const float* beta2;   // optional normalised residual
    const float* res;                                                                    // optional plain residual (fp32)
    const void* res_hi; const void* res_lo;                                              // ... or as a split-bf16 pair
    float* out;                                                                          // fp32 output, or
    void* out_hi; void* out_lo;                                                          // split-bf16 pair: v = hi + lo
    int C, level, relu;
};
cudaError_t launch_gn_apply(const GnApplyArgs& a, const int* img_off, int nimg, int nchunk, cudaStream_t st);
// out2 (fp32) or the split-bf16 pair (out_hi, out_lo)
cudaError_t launch_gn_apply_maxpool(const float* raw1, const float* stats, const float* gamma, const float* beta,
                                    float* out2, void* out_hi, void* out_lo, const int* img_off, const int* img_hw, int nimg,
                                    int total_p2, cudaStream_t st);
// Explicit im2col of a split-bf16 NHWC activation: out[m][(ky*k+kx)*C + c] for output pixel m (level lout) of a
// k x k / stride convolution with TF-SAME padding `pad` before (utils.py:93-123); zero where the tap is out of range.
cudaError_t launch_im2col_split(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, const ConvGather& cg,
                                long total_out_pixels, cudaStream_t st);
cudaError_t launch_im2col_patch(const float* img, float* cols, const int* img_off, const int* img_hw, int nimg,
                                int total_p4, cudaStream_t st);
cudaError_t launch_assemble_tokens(const float* proj /*[P4,256]*/, const float* cls, const float* pos, float* x0,
                                   const int* img_off, const int* img_hw, const int* tok_off, int nimg, int total_tok,
                                   cudaStream_t st);

// ---- input side (preprocess.cu): uint8 H x W x C (C = 1 or 3, interleaved) images at in + in_off[b] -> float32 (Hp, Wp) at
// out + out_off[b]: 1 - gray / 255, zero outside the source image.  hwc = [B][3] (H, W, C), out_hw = [B][2] (Hp, Wp).
cudaError_t launch_preprocess_u8(const uint8_t* in, const long* in_off, const int* hwc, const long* out_off, const int* out_hw,
                                 float* out, int nimg, long max_out_pixels, cudaStream_t st);

// ---- row-wise kernels (rowwise.cu)
struct Ln2Args {
    const float* in; int rows;
    const float* g1; const float* b1;      // nullable: first LN skipped (x1 = in)
    const float* g2; const float* b2;      // nullable: second LN skipped
    float* o1f; void* o1a;                 // x1 as fp32 and/or as activation type (nullable)
    void* o2a;                             // x2 as activation type (nullable)
    int dt_a;
    int late_trigger;                      // PDL: release the dependent kernel after the stores instead of at entry
};
cudaError_t launch_ln2(const Ln2Args& a, cudaStream_t st);
// x[b] = tok_emb[id[b]] + pos_emb[pos]; xn = LN(x).  Decode step: ids = cur_tok [B], pos = *step.
// Teacher-forced: ids = [B*T] row-major, pos = row % T (step == nullptr).
cudaError_t launch_embed_ln(const int64_t* ids, const int* step, int T, int rows, const float* tok_emb,
                            const float* pos_emb, int vocab, const float* g, const float* b, float* x, void* xn,
                            int dt_a, cudaStream_t st);
struct ArgmaxArgs {
    const float* logits; int B, V;
    const float2* partials; int nparts;  // greedy only, non-null: per row `nparts` {max, index bits} partials of the logits (EPI_ARGMAX GEMM) instead of `logits`
    int64_t* out_ids; int out_ld;        // out_ids[b*out_ld + step]
    int64_t* cur_tok;                    // [B] next input token
    int* step;                           // device step counter (incremented by the last block)
    int* seen_eos;                       // [B]
    int* done_step;                      // first step count at which all rows had an EOS (0 = not yet)
    int* block_counter;                  // scratch, zero-initialised
    int eos;
    // sampling (model/decoder.py:103-108): topk > 0 replaces the argmax by a draw from softmax(top-k logits / temp);
    // u = Philox4x32-10(key = seed, counter = (row_base + row, step, *call_ctr, 0)); requires V <= 1024
    int topk; float inv_temp; unsigned long long seed; int row_base; const unsigned* call_ctr;
    // next step's input, produced by the warp that picked the token (saves the embed launch on the step's critical path):
    // x = tok_emb[token] + pos_emb[step + 1] (fp32), xn = LN(x) in the activation type (skipped when emb_g == nullptr);
    // nothing is written when step + 1 == emb_max_pos.  emb_x == nullptr disables it.
    const float* tok_emb; const float* pos_emb; const float* emb_g; const float* emb_b; float* emb_x; void* emb_xn;
    int emb_dt, emb_max_pos;
};
cudaError_t launch_argmax_step(const ArgmaxArgs& a, cudaStream_t st);
cudaError_t launch_cross_entropy(const float* logits, const int64_t* tgt, int64_t rows, int V, float* row_loss,
                                 float* loss, cudaStream_t st);
// busy-wait for `ns` nanoseconds on the stream (used to phase-shift the concurrent decode branches)
cudaError_t launch_delay(long ns, cudaStream_t st);
cudaError_t launch_cast_f32_to(const float* in, void* out, int64_t n, int dt, cudaStream_t st);

// ---- attention (attention.cu)
struct AttnVarlenArgs {
    const void* q; int ldq;            // row-major, head h at columns [h*64, h*64+64)
    const void* k; int ldk;
    const void* v; int ldv;
    void* o; int ldo;
    const int* q_off; const int* q_len;   // per batch row ranges (q_len may be null -> q_off[b+1]-q_off[b])
    const int* k_off; const int* k_len;
    const uint8_t* q_mask; const uint8_t* k_mask;   // per ROW (packed like q / k), nullable
    int batch, max_q, causal, dt;
};
cudaError_t launch_attn_varlen(const AttnVarlenArgs& a, cudaStream_t st);
// bf16, unmasked, non-causal (encoder self-attention): mma.sync flash-attention kernel (attn_enc_mma.cu)
bool attn_enc_mma_supported(const AttnVarlenArgs& a);
cudaError_t launch_attn_enc_mma(const AttnVarlenArgs& a, cudaStream_t st);
struct AttnDecodeArgs {
    const void* q; int ldq;            // [B, ldq], head h at q + h*64
    const void* knew; const void* vnew; int ldnew;   // self: this step's k/v rows (appended at *step); null for cross
    void* kcache; void* vcache;        // K / V of head 0, key 0 (self: of sequence 0); V usually = K + 64 (head-major rows)
    int ldkv;                          // key (row) stride in elements
    int64_t batch_stride;              // self: elements between sequences; cross: unused
    int64_t head_stride;               // elements between heads (0 -> 64: heads side by side inside a row)
    const int* k_off; const int* k_len;   // cross: per-row range; null for self
    const int* step;                   // self: number of cached keys before this step = *step
    void* o; int ldo;
    int batch, dt;
    // debug timeline (TMA kernel only; null = off): entry / ready / end globaltimer stamps of launch `trace_k` of step *trace_step
    unsigned long long* trace; const int* trace_step; int trace_k;
};
// nk_cap >= the largest key count any row can have (sizes the per-head score buffer in shared memory)
cudaError_t launch_attn_decode(const AttnDecodeArgs& a, int nk_cap, cudaStream_t st);

// bf16 tier: persistent TMA-fed streaming version of the decode attention (attn_decode_tma.cu).
// map_base / map_rows / map_cols describe the 2-D K|V matrix the rows live in (row stride a.ldkv); col0 = column of
// head 0's K inside a row; tcap = cache rows per sequence (self-attention).
// Where the K|V rows of (sequence b, head h) live inside the 2-D matrix the tensor map covers:
//   column of K = col0 + h*col_h, column of V = that + v_col, first row = (self ? b*row_b : k_off[b]) + h*row_h.
// Head-major cache (the engine's layout): rows of 128 = [K 64 | V 64], col_h = 0, v_col = 64, row_h = keys per head.
struct KvLayout {
    const void* map_base; long map_rows; int map_cols; int ld;
    int col0, col_h, v_col, row_h, row_b;
};
bool attn_decode_tma_supported(const AttnDecodeArgs& a);
// bf16 tier: decode attention with the K / V projections absorbed into the query / output projections (attn_decode_tma.cu).
// q [batch, ldq] holds 8 absorbed queries of 256 per row; `latent` is the bf16 matrix [latent_rows, 256] the keys live in:
//   cross (znew == null): the encoder memory, k_off[batch + 1] = token range of every sequence;
//   self  (znew != null): the layer's latent cache [batch][tcap][256] (*step rows valid per sequence); znew [batch, ldz] is this
//   step's own latent row, which the kernel uses as key *step and appends to the cache.
// o [batch, ldo] receives 8 x 256 softmax-weighted latent averages per row.
struct AttnAbsArgs {
    const void* q; int ldq;
    const void* latent; long latent_rows;
    const int* k_off;
    int uni_nk;                   // cross: > 0 when every sequence has uni_nk memory tokens laid out back to back FROM ROW 0 OF `latent` (k_off[b] = b * uni_nk)
    const void* znew; int ldz; int tcap; const int* step;
    void* o; int ldo;
    int batch;
    unsigned long long* trace; const int* trace_step; int trace_k;
    unsigned long long* dbg;      // debug phase timers (5 words), null = off
};
cudaError_t launch_attn_abs(const AttnAbsArgs& a, int max_ctas, cudaStream_t st);
// the same attention with one warp per sequence (attn_decode_seq.cu); g_attn_seq selects it for the generate loop
cudaError_t launch_attn_seq(const AttnAbsArgs& a, int max_ctas, cudaStream_t st);
extern int g_attn_seq;
cudaError_t launch_attn_decode_tma(const AttnDecodeArgs& a, const KvLayout& lay, int max_ctas, cudaStream_t st);
// [tok][L*1024] (per layer: K 512 | V 512, heads side by side; the cross-K/V GEMM output) -> [L][8][tok][K 64 | V 64]
cudaError_t launch_crosskv_head_major(const void* in, void* out, int ntok, int layers, int dt, cudaStream_t st);
