// Attention cores of model/attention.py:148-173 (energy = q.k^T * 0.125, masks, softmax, .v), fp32 math.
//
//  attn_varlen : ragged batches of (q rows, k rows) -- encoder self-attention over images of different
//                widths (each image attends only to its own h*w+1 tokens), and the
//                teacher-forced decoder (causal self with pad masks, cross).  Mask semantics follow the
//                reference exactly: disallowed entries are filled with -FLT_MAX (utils.py:81-83), so a
//                fully masked query row softmaxes to the uniform average over ALL keys.
//  attn_decode : one query row per (sequence, head) against the KV cache -- the HBM-bound kernel of the
//                generate loop.  Self-attention appends this step's k/v to the cache first.
#include <float.h>

#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int HD = 64;       // dim_head (model/attention.py:76)
constexpr int NH = 8;
constexpr float SCALE = 0.125f;

// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) attn_varlen_kernel(AttnVarlenArgs a) {
    constexpr int QT = 32, KT = 32, R = 4;
    __shared__ float Qs[QT][HD];
    __shared__ float Ks[KT][HD + 1];
    __shared__ float Vs[KT][HD];
    __shared__ __align__(16) float Ps[8][KT][R];
    __shared__ int s_need_all;

    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int q0 = a.q_off[b], nq = a.q_len ? a.q_len[b] : a.q_off[b + 1] - q0;
    const int k0 = a.k_off[b], nk = a.k_len ? a.k_len[b] : a.k_off[b + 1] - k0;
    if (qt * QT >= nq || nk <= 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const T* __restrict__ Q = reinterpret_cast<const T*>(a.q);
    const T* __restrict__ K = reinterpret_cast<const T*>(a.k);
    const T* __restrict__ V = reinterpret_cast<const T*>(a.v);

    if (tid == 0) s_need_all = 0;
    __syncthreads();
    {   // Q tile -> smem (rows beyond nq are zero)
        const int r = tid >> 3, c = (tid & 7) * 8;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const int i = qt * QT + r;
        if (i < nq) {
            ld8(Q + (size_t)(q0 + i) * a.ldq + h * HD + c, v);
            if (a.q_mask && !a.q_mask[q0 + i]) s_need_all = 1;     // benign race: all writers store 1
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) Qs[r][c + e] = v[e];
    }
    __syncthreads();

    // rows of this warp
    int qi[R]; bool qok[R], qm[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        qi[r] = qt * QT + warp * R + r;
        qok[r] = qi[r] < nq;
        qm[r] = qok[r] && (!a.q_mask || a.q_mask[q0 + qi[r]]);
    }
    const int shift = nk - nq;                      // causal: key j allowed iff j <= i + (J - I)
    float m[R], l[R], acc0[R], acc1[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { m[r] = -INFINITY; l[r] = 0.f; acc0[r] = 0.f; acc1[r] = 0.f; }

    int k_end = nk;
    if (a.causal && !s_need_all) k_end = min(nk, qt * QT + QT + shift);   // later keys are masked for every row here
    if (k_end <= 0) k_end = nk;
    for (int kc = 0; kc < k_end; kc += KT) {
        __syncthreads();
        {   // K / V chunk -> smem
            const int r = tid >> 3, c = (tid & 7) * 8;
            float kv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, vv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const int j = kc + r;
            if (j < nk) {
                ld8(K + (size_t)(k0 + j) * a.ldk + h * HD + c, kv);
                ld8(V + (size_t)(k0 + j) * a.ldv + h * HD + c, vv);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) { Ks[r][c + e] = kv[e]; Vs[r][c + e] = vv[e]; }
        }
        __syncthreads();
        // scores: lane = key j, R rows at once
        float s[R];
#pragma unroll
        for (int r = 0; r < R; ++r) s[r] = 0.f;
#pragma unroll 4
        for (int d = 0; d < HD; d += 4) {
            const float k0v = Ks[lane][d], k1v = Ks[lane][d + 1], k2v = Ks[lane][d + 2], k3v = Ks[lane][d + 3];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float4 q = *reinterpret_cast<const float4*>(&Qs[warp * R + r][d]);
                s[r] = fmaf(q.x, k0v, s[r]); s[r] = fmaf(q.y, k1v, s[r]);
                s[r] = fmaf(q.z, k2v, s[r]); s[r] = fmaf(q.w, k3v, s[r]);
            }
        }
        const int j = kc + lane;
        const bool jin = j < nk;
        const bool km = jin && (!a.k_mask || a.k_mask[k0 + j]);
        float p[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float e = s[r] * SCALE;
            const bool allowed = qm[r] && km && (!a.causal || j <= qi[r] + shift);
            if (!allowed) e = -FLT_MAX;
            if (!jin) e = -INFINITY;                 // tile padding: contributes exactly nothing
            const float mn = fmaxf(m[r], warp_max(e));
            const float corr = expf(m[r] - mn);      // m = -inf on the first chunk -> 0
            p[r] = expf(e - mn);
            l[r] = l[r] * corr + warp_sum(p[r]);
            acc0[r] *= corr; acc1[r] *= corr;
            m[r] = mn;
        }
        *reinterpret_cast<float4*>(&Ps[warp][lane][0]) = make_float4(p[0], p[1], p[2], p[3]);
        __syncwarp();
        // P.V: lane owns output dims (2*lane, 2*lane+1)
#pragma unroll 8
        for (int jj = 0; jj < KT; ++jj) {
            const float4 pp = *reinterpret_cast<const float4*>(&Ps[warp][jj][0]);
            const float2 vv = *reinterpret_cast<const float2*>(&Vs[jj][2 * lane]);
            acc0[0] = fmaf(pp.x, vv.x, acc0[0]); acc1[0] = fmaf(pp.x, vv.y, acc1[0]);
            acc0[1] = fmaf(pp.y, vv.x, acc0[1]); acc1[1] = fmaf(pp.y, vv.y, acc1[1]);
            acc0[2] = fmaf(pp.z, vv.x, acc0[2]); acc1[2] = fmaf(pp.z, vv.y, acc1[2]);
            acc0[3] = fmaf(pp.w, vv.x, acc0[3]); acc1[3] = fmaf(pp.w, vv.y, acc1[3]);
        }
        __syncwarp();
    }
    T* __restrict__ O = reinterpret_cast<T*>(a.o);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (!qok[r]) continue;
        const float inv = 1.0f / l[r];
        st2(O + (size_t)(q0 + qi[r]) * a.ldo + h * HD + 2 * lane, acc0[r] * inv, acc1[r] * inv);
    }
}

// ------------------------------------------------------------------------------------------------
// grid = batch, block = 8 warps = 8 heads.  Dynamic smem: 8 * nk_cap floats of scores.
template <typename T>
__global__ void __launch_bounds__(256) attn_decode_kernel(AttnDecodeArgs a, int nk_cap) {
    extern __shared__ float s_scores[];
    __shared__ __align__(16) float s_q[NH][HD];
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* ps = s_scores + (size_t)h * nk_cap;
    const T* __restrict__ Q = reinterpret_cast<const T*>(a.q) + (size_t)b * a.ldq + h * HD;
    const T* kbase; const T* vbase;
    int nk;
    if (a.knew) {     // self-attention: append this step's key / value at position t, then attend to 0..t
        const int t = *a.step;
        const size_t hs = a.head_stride ? (size_t)a.head_stride : (size_t)HD;
        T* kc = reinterpret_cast<T*>(a.kcache) + (size_t)b * a.batch_stride + h * hs;
        T* vc = reinterpret_cast<T*>(a.vcache) + (size_t)b * a.batch_stride + h * hs;
        const T* kn = reinterpret_cast<const T*>(a.knew) + (size_t)b * a.ldnew + h * HD;
        const T* vn = reinterpret_cast<const T*>(a.vnew) + (size_t)b * a.ldnew + h * HD;
        kc[(size_t)t * a.ldkv + 2 * lane] = kn[2 * lane];
        kc[(size_t)t * a.ldkv + 2 * lane + 1] = kn[2 * lane + 1];
        vc[(size_t)t * a.ldkv + 2 * lane] = vn[2 * lane];
        vc[(size_t)t * a.ldkv + 2 * lane + 1] = vn[2 * lane + 1];
        kbase = kc; vbase = vc; nk = t + 1;
    } else {
        const int off = a.k_off[b];
        nk = a.k_len ? a.k_len[b] : a.k_off[b + 1] - off;
        const size_t hs = a.head_stride ? (size_t)a.head_stride : (size_t)HD;
        kbase = reinterpret_cast<const T*>(a.kcache) + (size_t)off * a.ldkv + h * hs;
        vbase = reinterpret_cast<const T*>(a.vcache) + (size_t)off * a.ldkv + h * hs;
    }
    s_q[h][2 * lane] = to_f(Q[2 * lane]);
    s_q[h][2 * lane + 1] = to_f(Q[2 * lane + 1]);
    __syncwarp();
    if (nk > nk_cap) nk = nk_cap;      // cannot happen: the host sizes nk_cap from max_length / max S

    // phase 1: lane = key
    float mx = -INFINITY;
    for (int j = lane; j < nk; j += 32) {
        const T* kr = kbase + (size_t)j * a.ldkv;
        float kv[HD];
#pragma unroll
        for (int c = 0; c < HD; c += 8) ld8(kr + c, kv + c);
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
            const float4 q = *reinterpret_cast<const float4*>(&s_q[h][d]);
            s = fmaf(q.x, kv[d], s); s = fmaf(q.y, kv[d + 1], s); s = fmaf(q.z, kv[d + 2], s); s = fmaf(q.w, kv[d + 3], s);
        }
        s *= SCALE;
        ps[j] = s;
        mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < nk; j += 32) {
        const float p = expf(ps[j] - mx);
        ps[j] = p;
        sum += p;
    }
    sum = warp_sum(sum);
    __syncwarp();
    // phase 2: lane = (key group kg of 4, dims dg*8 .. dg*8+7)
    const int dg = lane & 7, kg = lane >> 3;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int j = kg; j < nk; j += 4) {
        float vv[8];
        ld8(vbase + (size_t)j * a.ldkv + dg * 8, vv);
        const float p = ps[j];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(p, vv[e], acc[e]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
        acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
    }
    if (kg == 0) {
        const float inv = 1.0f / sum;
        T* o = reinterpret_cast<T*>(a.o) + (size_t)b * a.ldo + h * HD + dg * 8;
        st4(o, make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv));
        st4(o + 4, make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv));
    }
}

}  // namespace

template <typename T>
__global__ void __launch_bounds__(256) crosskv_head_major_kernel(const T* __restrict__ in, T* __restrict__ out, int ntok, int layers) {
    // one thread per 8 elements of the output: idx -> (l, h, tok, part K/V, chunk of 8)
    const long idx = (long)blockIdx.x * 256 + threadIdx.x;
    const long total = (long)layers * 8 * ntok * 16;
    if (idx >= total) return;
    const int chunk = (int)(idx & 7), part = (int)((idx >> 3) & 1);
    long rest = idx >> 4;
    const int tok = (int)(rest % ntok); rest /= ntok;
    const int h = (int)(rest & 7), l = (int)(rest >> 3);
    const T* src = in + ((size_t)tok * layers + l) * 1024 + part * 512 + h * 64 + chunk * 8;
    T* dst = out + idx * 8;
    float v[8];
    ld8(src, v);
    st4(dst, make_float4(v[0], v[1], v[2], v[3]));
    st4(dst + 4, make_float4(v[4], v[5], v[6], v[7]));
}

cudaError_t launch_crosskv_head_major(const void* in, void* out, int ntok, int layers, int dt, cudaStream_t st) {
    const long total = (long)layers * 8 * ntok * 16;
    if (total <= 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    if (dt == DT_F32) crosskv_head_major_kernel<float><<<blocks, 256, 0, st>>>((const float*)in, (float*)out, ntok, layers);
    else crosskv_head_major_kernel<bf16><<<blocks, 256, 0, st>>>((const bf16*)in, (bf16*)out, ntok, layers);
    return cudaGetLastError();
}

cudaError_t launch_attn_varlen(const AttnVarlenArgs& a, cudaStream_t st) {
    if (a.batch <= 0 || a.max_q <= 0) return cudaSuccess;
    dim3 grid((a.max_q + 31) / 32, NH, a.batch);
    if (a.dt == DT_F32) attn_varlen_kernel<float><<<grid, 256, 0, st>>>(a);
    else attn_varlen_kernel<bf16><<<grid, 256, 0, st>>>(a);
    return cudaGetLastError();
}

static int g_decode_smem_set[2] = {0, 0};

cudaError_t launch_attn_decode(const AttnDecodeArgs& a, int nk_cap, cudaStream_t st) {
    if (a.batch <= 0) return cudaSuccess;
    const size_t smem = (size_t)NH * nk_cap * sizeof(float);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    const int ti = a.dt == DT_F32 ? 0 : 1;
    if (smem > 40 * 1024 && g_decode_smem_set[ti] < (int)smem) {
        cudaError_t e = a.dt == DT_F32
            ? cudaFuncSetAttribute(attn_decode_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
            : cudaFuncSetAttribute(attn_decode_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        g_decode_smem_set[ti] = (int)smem;
    }
    if (a.dt == DT_F32) return launch_pdl(PDL_ATTN_SIMPLE, attn_decode_kernel<float>, dim3(a.batch), dim3(256), smem, st, a, nk_cap);
    return launch_pdl(PDL_ATTN_SIMPLE, attn_decode_kernel<bf16>, dim3(a.batch), dim3(256), smem, st, a, nk_cap);
}
