/*
 * texocr.h -- C-ABI of the B200-native TeXOCR inference path (libtexocr_b200.so).
 *
 * The reference (olibridge01/TeXOCR) is pure Python/PyTorch and has no FFI of its own
 * (SURVEY.md 2.1); its boundary for this path is the nn.Module surface of
 * model/ocr_model.py.  Each entry point below is what a binding for that surface calls,
 * and names the reference interface it replaces.  Plain pointers and sizes only: no
 * torch types cross this line.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *  - every function returns 0 on success, a negative texocr_status otherwise;
 *    texocr_last_error(h) (or texocr_last_error(NULL) for a failed create) gives the message.
 *  - data pointers may be HOST or DEVICE memory; the library asks the CUDA runtime which
 *    (cudaPointerGetAttributes) and stages host buffers through its own pinned/device staging
 *    on `stream`.  Device pointers must belong to the handle's device.
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  All work is
 *    enqueued on it; calls that return host-visible results synchronise that stream.
 *  - a handle is not thread-safe; different handles (also on the same GPU) are independent and may be driven from different
 *    host threads concurrently -- that is how several batches are kept in flight (texocr_b200/pipeline.py).
 *  - hw, enc_len and n_steps are always HOST arrays (they drive the host-side launch plan).
 *  - images: float32, single channel, row-major, ink = 1 on a 0 background
 *    (data_wrangling/dataset.py:365-371).  A batch is RAGGED: image i is hw[2*i] x hw[2*i+1],
 *    stored back to back (image i starts at sum_{j<i} H_j*W_j floats).  H,W must be multiples of
 *    16 with H <= 160 and W <= 1008 (position grid 10x63, model/encoder.py:137-143).
 *  - encoder memory: float32, packed tokens [sum_i N_i, 256], N_i = (H_i/16)*(W_i/16)+1,
 *    cls token first, then row-major patches (model/encoder.py:128-152).
 */
#ifndef TEXOCR_H_
#define TEXOCR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TEXOCR_ABI_VERSION 1
#if defined(__GNUC__)
#define TEXOCR_API __attribute__((visibility("default")))
#else
#define TEXOCR_API
#endif

typedef struct texocr_handle texocr_handle;

typedef enum {
    TEXOCR_OK = 0,
    TEXOCR_ERR_ARG = -1,        /* bad argument / unsupported shape                       */
    TEXOCR_ERR_CUDA = -2,       /* a CUDA call failed; message holds cudaGetErrorString   */
    TEXOCR_ERR_STATE = -3,      /* call order (weights not finalised, ...)                 */
    TEXOCR_ERR_WEIGHT = -4,     /* missing / mis-shaped state_dict entry                  */
    TEXOCR_ERR_NODEVICE = -5    /* no sm_100 device: there is NO CPU fallback             */
} texocr_status;

typedef enum { TEXOCR_FP32 = 0, TEXOCR_BF16 = 1 } texocr_precision;
typedef enum { TEXOCR_ENC_HYBRID = 0, TEXOCR_ENC_PATCH = 1 } texocr_encoder_kind;

/* Hyper-parameters: the keys create_model reads from config.yml
 * (model/ocr_model.py:113-130, model/encoder.py:172-191, model/decoder.py:148-173). */
typedef struct {
    int32_t abi_version;     /* TEXOCR_ABI_VERSION                                            */
    int32_t vocab_size;      /* config['vocab_size'] (tokenizer_clean_1k.txt:1 -> 1000)       */
    int32_t max_length;      /* config['max_length']: decoder positional table length         */
    int32_t enc_layers;      /* encoder.num_layers                                            */
    int32_t dec_layers;      /* decoder.num_layers                                            */
    int32_t bos_token, eos_token, pad_token;
    int32_t encoder_kind;    /* texocr_encoder_kind                                           */
    int32_t precision;       /* texocr_precision: FP32 = parity tier (FFMA everywhere);
                                BF16 = bf16 operands / attention cache, fp32 accumulate + statistics; the
                                generate loop runs its attention in the absorbed (latent) form, DESIGN.md 5c    */
} texocr_config;

/* replaces: create_model(config) (model/ocr_model.py:113-130) */
TEXOCR_API int texocr_create(const texocr_config* cfg, int device, texocr_handle** out);
TEXOCR_API void texocr_destroy(texocr_handle* h);
TEXOCR_API const char* texocr_last_error(const texocr_handle* h);

/* replaces: nn.Module.load_state_dict(state_dict) (utils.py:63-71, model/ocr_model.py:82-90).
 * One call per state_dict entry, with the reference's key name; float32,
 * contiguous, host or device.  Aliased keys (layers.{i}.0.*, block.*) may be passed or omitted. */
TEXOCR_API int texocr_set_weight(texocr_handle* h, const char* name, const float* data, int32_t ndim, const int64_t* shape);
/* Folds weight standardisation (model/resnet.py:61-64), concatenates q/k/v, interleaves GLU / GeGLU
 * columns, converts to the compute precision and uploads.  Must be called after the last set_weight. */
TEXOCR_API int texocr_finalize_weights(texocr_handle* h);

/* replaces: VisionEncoder.forward (model/encoder.py:128-152) incl. the ResNetV2 stem
 * (model/resnet.py:251-254).  enc_out: float32 [sum N_i, 256]. */
TEXOCR_API int texocr_encode(texocr_handle* h, const float* images, const int32_t* hw, int32_t batch,
                  float* enc_out, void* stream);

/* replaces: Transformer.forward(ids, mask=, enc=) (model/decoder.py:41-67) -- teacher-forced logits.
 * ids int64 [batch, T]; mask uint8 [batch, T] (1 = real token) or NULL; enc packed float32 with
 * enc_len[batch] tokens per row; logits_out float32 [batch, T, vocab]. */
TEXOCR_API int texocr_decoder_logits(texocr_handle* h, const int64_t* ids, const uint8_t* mask, const float* enc,
                          const int32_t* enc_len, int32_t batch, int32_t T, float* logits_out, void* stream);

/* replaces: AutoRegressiveDecoder.generate(start_tokens (B,1), eos_tok, max_len, enc=) (model/decoder.py:77-122),
 * greedy (argmax; temp -> 0 limit of lines 104-108).  out_ids int64 [batch, max_len] (row stride max_len);
 * *n_steps = number of valid columns = first step at which every row has produced eos_tok, else max_len
 * (model/decoder.py:115-118).  eos_tok < 0 disables the early exit.  Requires max_len <= max_length. */
TEXOCR_API int texocr_decoder_generate(texocr_handle* h, const int64_t* start_tokens, int32_t eos_tok, const float* enc,
                            const int32_t* enc_len, int32_t batch, int32_t max_len, int64_t* out_ids,
                            int32_t* n_steps, void* stream);

/* replaces: OCRModel.generate(src, max_len) (model/ocr_model.py:46-66): encoder once, BOS column,
 * decode loop; bos/eos from the config.  Same outputs as texocr_decoder_generate. */
TEXOCR_API int texocr_generate(texocr_handle* h, const float* images, const int32_t* hw, int32_t batch, int32_t max_len,
                    int64_t* out_ids, int32_t* n_steps, void* stream);

/* replaces: the deterministic part of img_transform (data_wrangling/dataset.py:365-371: ToTensor -> Grayscale(1) ->
 * Invert; the RandomAffine in front of it is train-time noise and is not applied) for a ragged batch: uint8 images,
 * H x W x C interleaved with C = 1 or 3, concatenated in `pixels`; hwc int32 [batch][3] = (H, W, C).  Output: float32
 * images 1 - gray/255 packed like texocr_encode / texocr_generate take them, every image zero-padded (white background)
 * at the right / bottom to a multiple of pad_multiple (16 = the encoder's patch grid; 1 = none); out_hw int32 [batch][2].
 * pixels / out_images may be host or device pointers; bit-exact with torchvision on the CPU.  Needs no weights. */
TEXOCR_API int texocr_preprocess_u8(texocr_handle* h, const uint8_t* pixels, const int32_t* hwc, int32_t batch,
                         int32_t pad_multiple, float* out_images, int32_t* out_hw, void* stream);

/* replaces: the sampling branch of AutoRegressiveDecoder.generate (model/decoder.py:103-108) for every later
 * texocr_generate / texocr_decoder_generate call on this handle: keep the k = int((1 - threshold) * vocab) largest
 * logits (utils.py:85-91; threshold 0.9 -> k = 99 of 1000), p = softmax(kept / temp), one draw per row and step.
 * The draw is the inverse CDF in vocabulary order at u = (x >> 8) * 2^-24 with x = word 0 of
 * Philox4x32-10(key = seed, counter = (row, step, call, 0)), call = number of sampled generate calls since this
 * function: reproducible and independent of how the batch is split; parity with torch.multinomial is
 * distribution-level (tests/test_gpu_sampling.py).  temp <= 0 restores greedy decoding.  vocab <= 1024. */
TEXOCR_API int texocr_set_sampling(texocr_handle* h, double temp, double threshold, uint64_t seed);

/* replaces: AutoRegressiveDecoder.forward / OCRModel.forward loss (model/decoder.py:124-145):
 * mean cross-entropy (no ignore_index) of logits [rows, vocab] against targets int64 [rows]. */
TEXOCR_API int texocr_cross_entropy(texocr_handle* h, const float* logits, const int64_t* targets, int64_t rows,
                         float* loss_out, void* stream);

/* ---- instrumentation (bench.py / tests; not part of the reference surface) ---------------- */
/* Number of kernels this library launched since the handle was created (CUDA-graph replays count
 * the kernels they contain). */
TEXOCR_API int64_t texocr_kernel_launches(const texocr_handle* h);
/* Per-kernel-class device timing: when enabled, every launch is bracketed by CUDA events on the
 * launching stream (CUDA graphs are bypassed).  texocr_profile_read synchronises and reports, per
 * class, launches, total milliseconds and total algorithmic bytes / flops as accounted by the engine.
 * Returns the number of classes written (<= cap). */
typedef struct {
    char name[48];
    int64_t launches;
    double ms;
    double bytes;
    double flops;
} texocr_profile_row;
TEXOCR_API int texocr_profile_enable(texocr_handle* h, int32_t on);
TEXOCR_API int texocr_profile_read(texocr_handle* h, texocr_profile_row* rows, int32_t cap);
/* Tuning switches (name = "cuda_graph" | "tcgen05" | ...); returns TEXOCR_ERR_ARG for unknown names.  Round-2 additions (all default on,
 * results are bit-identical either way; the tests toggle them): "gemm_epi_warps" 4 | 8 epilogue warps per tcgen05 GEMM CTA, "gn_fused"
 * GroupNorm partial sums from the convolution GEMM's epilogue, "gemm_bn256" 128 x 256 tiles for the wide convolutions, "conv_gather"
 * cp.async-gathered implicit GEMMs for ragged batches (0 = explicit im2col buffer), "pdl_mid" bit mask of kernel families that release
 * their programmatic dependent late (1 GEMM, 2 decode attention, 4 LayerNorm). */
TEXOCR_API int texocr_set_option(texocr_handle* h, const char* name, int64_t value);
/* Debug tap: copy an internal activation of the last texocr_encode call to `out` (host or device).
 * name = "backbone" -> float32 [sum h_i*w_i, 1024] (NHWC pixels).  Returns element count or <0. */
TEXOCR_API int64_t texocr_debug_read(texocr_handle* h, const char* name, float* out, int64_t cap_elems);

/* Test hook: one token-selection step (greedy or, after texocr_set_sampling, sampled) on device logits
 * float32 [rows, vocab] with the given step / call counters; out_ids int64 [rows] (host or device). */
TEXOCR_API int texocr_debug_sample_step(texocr_handle* h, const float* logits, int32_t rows, int32_t step, uint32_t call,
                             int64_t* out_ids);

/* Test hook: run one GEMM  C[M,N] = epi(A[M,K] . W[N,K]^T)  on device buffers through the engine's own kernels.
 * dt_a: 0 = fp32 operands (FFMA kernel), 1 = bf16 operands; use_tc != 0 selects the tcgen05 kernel (bf16 only);
 * A2/W2 non-NULL select the bf16x3 split mode (low-order parts).  epi / dt_c as in csrc/kernels.h (GemmEpi, DT_*). */
TEXOCR_API int texocr_debug_gemm(texocr_handle* h, const void* A, const void* W, void* C, int32_t M, int32_t N, int32_t K,
                      int32_t lda, int32_t ldw, int32_t ldc, int32_t epi, int32_t dt_a, int32_t dt_c, const float* bias,
                      const float* res, int32_t ldres, int32_t use_tc, const void* A2, const void* W2, void* stream);

/* Test hook: one decode-attention launch (bf16) on the engine's head-major K/V layout (rows of 128 = K 64 | V 64).
 * self != 0: q/knew/vnew are [batch, ld] rows (head h at h*64), kv is the cache [batch][8 heads][tcap][128] (kv_rows =
 * batch*8*tcap), *step_dev keys already cached; the kernel appends key `step`.  self == 0: kv is [8 heads][ntok][128]
 * (kv_rows = 8*ntok), k_off_dev[batch+1] gives each sequence's token range.  ldkv / col0 are ignored.
 * use_tma selects the persistent TMA kernel, otherwise the simple per-sequence kernel.  out: bf16 [batch, 512]. */
TEXOCR_API int texocr_debug_attn_decode(texocr_handle* h, int32_t self, const void* q, int32_t ldq, const void* knew, const void* vnew,
                             int32_t ldnew, void* kv, int64_t kv_rows, int32_t ldkv, int32_t col0, int32_t tcap,
                             const int32_t* k_off_dev, const int32_t* step_dev, void* out, int32_t batch, int32_t max_keys,
                             int32_t use_tma, void* stream);

/* Test hook: one launch of the absorbed decode-attention kernel of the bf16 generate loop: q bf16
 * [batch, 8 x 256] absorbed queries; out bf16 [batch, 8 x 256] = softmax(q_h . Z^T / 8) . Z per head over the sequence's latent rows Z.
 * Cross (znew == NULL): latent bf16 [latent_rows, 256] = encoder memory, k_off_dev int32 [batch + 1] token ranges (device).
 * Self (znew != NULL): latent = cache bf16 [batch][tcap][256] with *step_dev valid rows per sequence; znew bf16 [batch, 256] is this
 * step's own row: used as key *step_dev and appended to the cache. */
TEXOCR_API int texocr_debug_attn_abs(texocr_handle* h, const void* q, void* latent, int64_t latent_rows, const int32_t* k_off_dev,
                          const void* znew, int32_t tcap, const int32_t* step_dev, void* out, int32_t batch, void* stream);

/* Test hook, host only (needs no device and no handle): the weight folding of the absorbed attention.
 * wq / wk / wv float32 [512, 256], wo float32 [512, 512] of one MultiHeadAttention (model/attention.py:87-99) ->
 * wqk_out float32 [2048, 256] (row h*256 + c = sum_d Wk[h*64+d][c] * Wq[h*64+d][:]) and wvo_out float32 [512, 2048] (column
 * h*256 + c = sum_d Wo[:, h*64+d] * Wv[h*64+d][c]; rows interleaved (value, gate) like the out-projection for its GLU epilogue),
 * both computed from the bf16 roundings of the inputs, as the bf16 tier does. */
TEXOCR_API int texocr_debug_fold_absorbed(const float* wq, const float* wk, const float* wv, const float* wo, float* wqk_out,
                               float* wvo_out);

#ifdef __cplusplus
}
#endif
#endif /* TEXOCR_H_ */
