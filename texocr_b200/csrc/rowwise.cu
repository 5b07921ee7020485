// Row-wise HBM-bound kernels: the shared double LayerNorm of model/attention.py:242-259, the decoder
// embedding (model/decoder.py:51-53), greedy argmax + EOS bookkeeping (model/decoder.py:103-116 in the
// temp -> 0 limit) and the teacher-forcing cross-entropy (model/decoder.py:140).
// One warp per 256-wide row, 8 elements per lane (two float4), fp32 statistics via warp shuffles.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int D = 256;
constexpr float LN_EPS = 1e-5f;

// gg / bb: the lane's 8 gamma / beta values (weights: the callers fetch them BEFORE griddepcontrol.wait)
TX_DEVINL void layer_norm8(float* v, const float* gg, const float* bb) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
    const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / D) + LN_EPS);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * rstd * gg[i] + bb[i];
}

template <typename T> TX_DEVINL void store8(T* p, const float* v) {
    st4(p, make_float4(v[0], v[1], v[2], v[3]));
    st4(p + 4, make_float4(v[4], v[5], v[6], v[7]));
}

template <typename TAct>
__global__ void __launch_bounds__(256) ln2_kernel(Ln2Args a) {
    if (!a.late_trigger) pdl_launch_dependents();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31, col = lane * 8;
    float g1[8], b1[8], g2[8], b2[8];
    if (a.g1) { ld8(a.g1 + col, g1); ld8(a.b1 + col, b1); }
    if (a.g2) { ld8(a.g2 + col, g2); ld8(a.b2 + col, b2); }
    pdl_wait();
    if (row >= a.rows) { if (a.late_trigger) pdl_launch_dependents(); return; }
    float v[8];
    ld8cg(a.in + (size_t)row * D + col, v);
    if (a.late_trigger == 2 && v[0] == v[0]) pdl_launch_dependents();      // once the row has arrived: the predicate makes the release wait for the load (a NaN row releases at exit)
    if (a.g1) layer_norm8(v, g1, b1);
    if (a.o1f) store8(a.o1f + (size_t)row * D + col, v);
    if (a.o1a) store8(reinterpret_cast<TAct*>(a.o1a) + (size_t)row * D + col, v);
    if (a.g2) {
        layer_norm8(v, g2, b2);
        if (a.o2a) store8(reinterpret_cast<TAct*>(a.o2a) + (size_t)row * D + col, v);
    }
    if (a.late_trigger == 1) pdl_launch_dependents();
}

template <typename TAct>
__global__ void __launch_bounds__(256) embed_ln_kernel(const int64_t* __restrict__ ids, const int* __restrict__ step, int T, int rows,
                                                       const float* __restrict__ tok_emb, const float* __restrict__ pos_emb, int vocab,
                                                       const float* __restrict__ g, const float* __restrict__ b,
                                                       float* __restrict__ x, TAct* __restrict__ xn) {
    pdl_launch_dependents();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31, col = lane * 8;
    float gg[8], bb[8];
    if (g != nullptr) { ld8(g + col, gg); ld8(b + col, bb); }
    pdl_wait();
    if (row >= rows) return;
    long id = (long)__ldcg(reinterpret_cast<const long long*>(ids) + row);
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const int pos = step ? ldcg_i32(step) : (row % T);
    float v[8], pe[8];
    ld8(tok_emb + (size_t)id * D + col, v);
    ld8(pos_emb + (size_t)pos * D + col, pe);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += pe[i];
    store8(x + (size_t)row * D + col, v);
    if (g == nullptr) return;          // embedding only: the consumer GEMM applies the LayerNorm itself (tc_gemm_ln_kernel)
    layer_norm8(v, gg, bb);
    store8(xn + (size_t)row * D + col, v);
}

// Philox4x32-10 (Salmon et al., SC'11): counter-based, so a draw depends only on (seed, row, step)
TX_DEVINL uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * ctr.x, p1 = (unsigned long long)0xCD9E8D57u * ctr.z;
        ctr = make_uint4((uint32_t)(p1 >> 32) ^ ctr.y ^ key.x, (uint32_t)p1, (uint32_t)(p0 >> 32) ^ ctr.w ^ key.y, (uint32_t)p0);
        key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
    }
    return ctr;
}
TX_DEVINL uint32_t order_key(float v) {          // monotonic float -> uint (larger float, larger key)
    const uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// One draw per row (one warp): top-k filter (k-th largest found by a bitwise radix descent over the warp's <= 1024 values),
// p = exp((v - max) / temp) over the kept values, inverse CDF in vocabulary order at u.  model/decoder.py:103-108, utils.py:85-91.
TX_DEVINL int sample_row(const float* __restrict__ l, int V, int k, float inv_temp, float u, int lane) {
    const int per = (V + 31) >> 5, i0 = lane * per;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (i < per && i0 + i < V) ? __ldcg(l + i0 + i) : -INFINITY;
    uint32_t prefix = 0;
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t cand = prefix | (1u << bit);
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) cnt += (i < per && i0 + i < V && order_key(v[i]) >= cand) ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (cnt >= k) prefix = cand;
    }
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, v[i]);
    mx = warp_max(mx);
    // exactly k logits survive (torch.topk + scatter, utils.py:85-91): everything above the k-th largest value, and of the values equal
    // to it the lowest indices until k are kept (lanes own ascending index ranges: an exclusive warp scan of the per-lane tie counts)
    int n_gt = 0, n_eq = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const bool in = i < per && i0 + i < V;
        n_gt += (in && order_key(v[i]) > prefix) ? 1 : 0;
        n_eq += (in && order_key(v[i]) == prefix) ? 1 : 0;
    }
    int tot_gt = n_gt, eq_before = n_eq;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot_gt += __shfl_xor_sync(0xffffffffu, tot_gt, o);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, eq_before, o);
        if (lane >= o) eq_before += t;
    }
    eq_before -= n_eq;                            // ties in lower lanes (= at lower indices)
    int eq_left = k - tot_gt - eq_before;         // ties this lane may still keep
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const bool in = i < per && i0 + i < V;
        const uint32_t key = order_key(v[i]);
        bool keep = in && key > prefix;
        if (in && key == prefix) { keep = eq_left > 0; --eq_left; }
        v[i] = keep ? expf((v[i] - mx) * inv_temp) : 0.f;
        part += v[i];
    }
    float incl = part;                            // inclusive scan of the lane sums
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const float total = __shfl_sync(0xffffffffu, incl, 31);
    const float target = u * total;
    const unsigned hit = __ballot_sync(0xffffffffu, incl > target && part > 0.f);
    int tok = -1;
    if (hit) {
        const int L = __ffs(hit) - 1;
        if (lane == L) {
            float c = incl - part;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                c += v[i];
                if (tok < 0 && v[i] > 0.f && c > target) tok = i0 + i;
            }
            if (tok < 0) {                        // rounding: the lane's last kept value
#pragma unroll
                for (int i = 0; i < 32; ++i) if (v[i] > 0.f) tok = i0 + i;
            }
        }
        tok = __shfl_sync(0xffffffffu, tok, L);
    } else {                                      // target rounded up to the total: last kept index of the row
        int last = -1;
#pragma unroll
        for (int i = 0; i < 32; ++i) if (v[i] > 0.f) last = i0 + i;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
        tok = last < 0 ? 0 : last;
    }
    return tok;
}

// x = tok_emb[tok] + pos_emb[pos]; xn = LN(x): the warp of a row prepares the row's input of the next decode step
TX_DEVINL void embed_next(const ArgmaxArgs& a, int row, int tok, int pos, int lane, const float* gg, const float* bb) {
    const int col = lane * 8;
    long id = tok < 0 ? 0 : (tok >= a.V ? a.V - 1 : tok);
    float v[8], pe[8];
    ld8(a.tok_emb + (size_t)id * D + col, v);
    ld8(a.pos_emb + (size_t)pos * D + col, pe);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += pe[i];
    store8(a.emb_x + (size_t)row * D + col, v);
    if (a.emb_g == nullptr) return;
    layer_norm8(v, gg, bb);
    if (a.emb_dt == DT_F32) store8(reinterpret_cast<float*>(a.emb_xn) + (size_t)row * D + col, v);
    else store8(reinterpret_cast<bf16*>(a.emb_xn) + (size_t)row * D + col, v);
}

__global__ void __launch_bounds__(256) argmax_step_kernel(ArgmaxArgs a) {
    __shared__ int s_last;
    pdl_launch_dependents();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    float egg[8], ebb[8];
    if (a.emb_x && a.emb_g) { ld8(a.emb_g + lane * 8, egg); ld8(a.emb_b + lane * 8, ebb); }
    pdl_wait();
    const int t = ldcg_i32(a.step);
    if (row < a.B && a.topk > 0) {
        const uint4 rnd = philox4x32_10(make_uint4((uint32_t)(a.row_base + row), (uint32_t)t, a.call_ctr ? __ldcg(a.call_ctr) : 0u, 0u),
                                        make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
        const float u = (float)(rnd.x >> 8) * (1.0f / 16777216.0f);
        const int tok = sample_row(a.logits + (size_t)row * a.V, a.V, a.topk, a.inv_temp, u, lane);
        if (lane == 0) {
            a.out_ids[(size_t)row * a.out_ld + t] = tok;
            a.cur_tok[row] = tok;
            if (a.eos >= 0 && tok == a.eos) a.seen_eos[row] = 1;
        }
        if (a.emb_x && t + 1 < a.emb_max_pos) embed_next(a, row, tok, t + 1, lane, egg, ebb);
    } else if (row < a.B) {
        float best = -INFINITY;
        int bi = 0x7fffffff;
        if (a.partials) {      // the vocabulary GEMM already reduced every 32-column tile to its (max, first index)
            const float2* pp = a.partials + (size_t)row * a.nparts;
            for (int i = lane; i < a.nparts; i += 32) {
                const float2 v = __ldcg(pp + i);
                const int vi = __float_as_int(v.y);
                if (v.x > best || (v.x == best && vi < bi)) { best = v.x; bi = vi; }
            }
        } else {
            const float* l = a.logits + (size_t)row * a.V;
            for (int i = lane; i < a.V; i += 32) {
                const float v = __ldcg(l + i);
                if (v > best) { best = v; bi = i; }       // ascending scan: first maximum wins (torch argmax)
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (bi == 0x7fffffff) bi = 0;             // every lane holds the row's result after the butterfly
        if (lane == 0) {
            a.out_ids[(size_t)row * a.out_ld + t] = bi;
            a.cur_tok[row] = bi;
            if (a.eos >= 0 && bi == a.eos) a.seen_eos[row] = 1;
        }
        if (a.emb_x && t + 1 < a.emb_max_pos) embed_next(a, row, bi, t + 1, lane, egg, ebb);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(a.block_counter, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (s_last) {
        __threadfence();
        int all = 1;
        for (int b = threadIdx.x; b < a.B; b += blockDim.x) all &= (*(volatile int*)(a.seen_eos + b) != 0);
        all = __syncthreads_and(all);
        if (threadIdx.x == 0) {
            if (all && *a.done_step == 0) *a.done_step = t + 1;
            *a.step = t + 1;
            *a.block_counter = 0;
        }
    }
}

__global__ void __launch_bounds__(256) ce_rows_kernel(const float* __restrict__ logits, const int64_t* __restrict__ tgt, long rows,
                                                      int V, float* __restrict__ row_loss) {
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* l = logits + (size_t)row * V;
    float mx = -INFINITY;
    for (int i = lane; i < V; i += 32) mx = fmaxf(mx, l[i]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int i = lane; i < V; i += 32) s += expf(l[i] - mx);
    s = warp_sum(s);
    if (lane == 0) {
        long t = (long)tgt[row];
        t = t < 0 ? 0 : (t >= V ? V - 1 : t);
        row_loss[row] = logf(s) + mx - l[t];
    }
}

__global__ void __launch_bounds__(1024) mean_kernel(const float* __restrict__ v, long n, float* __restrict__ out) {
    __shared__ double sh[1024];
    double s = 0.0;
    for (long i = threadIdx.x; i < n; i += 1024) s += (double)v[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = (float)(sh[0] / (double)n);
}

__global__ void delay_kernel(long ns) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
        __nanosleep(500);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    } while ((long)(t1 - t0) < ns);
}

template <typename T>
__global__ void __launch_bounds__(256) cast_kernel(const float* __restrict__ in, T* __restrict__ out, long n4) {
    const long i = (long)blockIdx.x * 256 + threadIdx.x;
    if (i < n4) st4(out + i * 4, ld4(in + i * 4));
}

}  // namespace

cudaError_t launch_ln2(const Ln2Args& a_in, cudaStream_t st) {
    Ln2Args a = a_in;
    a.late_trigger = (g_texocr_pdl_mid & 4) ? 2 : ((g_texocr_pdl >> 8) & 1);
    if (a.rows <= 0) return cudaSuccess;
    const int blocks = (a.rows + 7) / 8;
    if (a.dt_a == DT_F32) return launch_pdl(PDL_LN, ln2_kernel<float>, dim3(blocks), dim3(256), 0, st, a);
    return launch_pdl(PDL_LN, ln2_kernel<bf16>, dim3(blocks), dim3(256), 0, st, a);
}

cudaError_t launch_embed_ln(const int64_t* ids, const int* step, int T, int rows, const float* tok_emb,
                            const float* pos_emb, int vocab, const float* g, const float* b, float* x, void* xn,
                            int dt_a, cudaStream_t st) {
    if (rows <= 0) return cudaSuccess;
    const int blocks = (rows + 7) / 8;
    if (dt_a == DT_F32)
        return launch_pdl(PDL_EMBED, embed_ln_kernel<float>, dim3(blocks), dim3(256), 0, st, ids, step, T, rows, tok_emb, pos_emb, vocab, g, b, x, (float*)xn);
    return launch_pdl(PDL_EMBED, embed_ln_kernel<bf16>, dim3(blocks), dim3(256), 0, st, ids, step, T, rows, tok_emb, pos_emb, vocab, g, b, x, (bf16*)xn);
}

cudaError_t launch_argmax_step(const ArgmaxArgs& a, cudaStream_t st) {
    return launch_pdl(PDL_ARGMAX, argmax_step_kernel, dim3((a.B + 7) / 8), dim3(256), 0, st, a);
}

cudaError_t launch_cross_entropy(const float* logits, const int64_t* tgt, int64_t rows, int V, float* row_loss,
                                 float* loss, cudaStream_t st) {
    if (rows <= 0) return cudaErrorInvalidValue;
    ce_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(logits, tgt, (long)rows, V, row_loss);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    mean_kernel<<<1, 1024, 0, st>>>(row_loss, (long)rows, loss);
    return cudaGetLastError();
}

cudaError_t launch_delay(long ns, cudaStream_t st) {
    if (ns <= 0) return cudaSuccess;
    delay_kernel<<<1, 1, 0, st>>>(ns);
    return cudaGetLastError();
}

cudaError_t launch_cast_f32_to(const float* in, void* out, int64_t n, int dt, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    if (n % 4 != 0) return cudaErrorInvalidValue;
    const long n4 = n / 4;
    const unsigned blocks = (unsigned)((n4 + 255) / 256);
    if (dt == DT_F32) cast_kernel<float><<<blocks, 256, 0, st>>>(in, (float*)out, n4);
    else cast_kernel<bf16><<<blocks, 256, 0, st>>>(in, (bf16*)out, n4);
    return cudaGetLastError();
}
