// Shared device helpers for the TeXOCR sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

typedef __nv_bfloat16 bf16;

#define TX_DEVINL __device__ __forceinline__

template <typename T> TX_DEVINL float to_f(T v);
template <> TX_DEVINL float to_f<float>(float v) { return v; }
template <> TX_DEVINL float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> TX_DEVINL T from_f(float v);
template <> TX_DEVINL float from_f<float>(float v) { return v; }
template <> TX_DEVINL bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 4 consecutive elements <-> float4 (pointer must be aligned to 4 elements)
TX_DEVINL float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
TX_DEVINL float4 ld4(const bf16* p) {
    uint2 r = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}
TX_DEVINL void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
TX_DEVINL void st4(bf16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = r;
}
TX_DEVINL void st2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
TX_DEVINL void st2(bf16* p, float a, float b) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}
// 8 consecutive elements -> 8 floats (aligned to 8 elements)
TX_DEVINL void ld8(const float* p, float* o) {
    float4 a = ld4(p), b = ld4(p + 4);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
TX_DEVINL void ld8(const bf16* p, float* o) {
    uint4 r = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(h[i]);
        o[2 * i] = f.x;
        o[2 * i + 1] = f.y;
    }
}

// Loads of data that earlier kernels (re)write at the same address every step / sub-layer go to L2 (ld.global.cg): with
// several streams' kernels resident on an SM, a line cached in its L1 by an earlier kernel can survive into a later one
// (observed on B200 as rare stale reads of the step counter / query rows once K/V stopped streaming through L1).
TX_DEVINL int ldcg_i32(const int* p) { return __ldcg(p); }
TX_DEVINL uint4 ldcg_u4(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
TX_DEVINL uint32_t ldcg_u32(const void* p) { return __ldcg(reinterpret_cast<const unsigned int*>(p)); }
TX_DEVINL float4 ld4cg(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
TX_DEVINL float4 ld4cg(const bf16* p) {
    uint2 r = __ldcg(reinterpret_cast<const uint2*>(p));
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}
TX_DEVINL void ld8cg(const float* p, float* o) {
    float4 a = ld4cg(p), b = ld4cg(p + 4);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
TX_DEVINL void ld8cg(const bf16* p, float* o) {
    uint4 r = __ldcg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(h[i]);
        o[2 * i] = f.x;
        o[2 * i + 1] = f.y;
    }
}

TX_DEVINL float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
TX_DEVINL float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

TX_DEVINL float gelu_erf(float g) { return 0.5f * g * (1.0f + erff(g * 0.70710678118654752440f)); }
TX_DEVINL float sigmoidf_(float g) { return 1.0f / (1.0f + expf(-g)); }
// bf16 tier epilogues (tcgen05 GEMM): the result is rounded to 8 mantissa bits (GeGLU -> bf16 hidden) or added to an O(1) residual, so
// the exact-erf GELU (model/attention.py:17) and the sigmoid of nn.GLU are evaluated with hardware ex2 / rcp and the
// Abramowitz-Stegun 7.1.26 rational erf (|error| <= 1.5e-7): ~12 instead of ~50 instructions per value in a one-thread-per-row
// epilogue that is on the critical path of every decode GEMM.  The fp32 parity tier (FFMA kernels) keeps erff / expf.
TX_DEVINL float sigmoid_fast(float g) { return __fdividef(1.0f, 1.0f + __expf(-g)); }
TX_DEVINL float gelu_erf_fast(float g) {
    const float x = fabsf(g) * 0.70710678118654752440f;
    const float t = __fdividef(1.0f, fmaf(0.3275911f, x, 1.0f));
    const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
    const float erf_abs = 1.0f - poly * __expf(-x * x);
    return 0.5f * g * (1.0f + copysignf(erf_abs, g));
}

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may start
// while its predecessor drains.  pdl_launch_dependents() lets the successor be scheduled early; pdl_wait() blocks until
// the predecessor grid has completed and its memory is visible.  Both are no-ops for ordinary launches.  Every kernel
// on the decode path calls them first thing, before touching global memory.
TX_DEVINL void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
TX_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Ragged image batch geometry.  img_off[b] = pixel offset of image b at full resolution
// (sum of H*W of the images before it); every level L (stride 2^L) has H>>L x W>>L pixels
// starting at img_off[b] >> (2L) because H and W are multiples of 16.
struct ImgGeom {
    const int* img_off;   // [B+1]
    const int* img_hw;    // [B][2]
    int nimg;
};

// Same answer, for callers that walk the pixels in increasing order (persistent tile loops): a few forward steps from the image of the
// previous call instead of a binary search of dependent loads (ncu: the searches were a third of the stem kernel's stall samples).
// hint = the previous result (0 for the first call); falls back to the binary search when the hint is ahead of pix or far behind.
TX_DEVINL int find_image(const int* off, int nimg, int level, int pix);
TX_DEVINL int find_image_from(const int* off, int nimg, int level, int pix, int hint) {
    if (hint < 0 || hint >= nimg || (off[hint] >> (2 * level)) > pix) return find_image(off, nimg, level, pix);
    int b = hint;
#pragma unroll 1
    for (int step = 0; step < 8; ++step) {
        if (b + 1 >= nimg || (off[b + 1] >> (2 * level)) > pix) return b;
        ++b;
    }
    return find_image(off, nimg, level, pix);
}

TX_DEVINL int find_image(const int* off, int nimg, int level, int pix) {
    // largest b with (off[b] >> 2L) <= pix
    int lo = 0, hi = nimg - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if ((off[mid] >> (2 * level)) <= pix) lo = mid; else hi = mid - 1;
    }
    return lo;
}
