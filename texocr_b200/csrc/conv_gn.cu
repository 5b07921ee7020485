// Backbone pieces that are not GEMMs (model/resnet.py): the 7x7/s2 single-channel stem convolution,
// GroupNorm(32) statistics + apply (+ residual + ReLU), the stem max-pool, patch im2col and the
// cls / positional-embedding token assembly of model/encoder.py:128-143.
// Activations are ragged NHWC pixel batches [sum_i h_i*w_i, C] (see ImgGeom in common.cuh).
// All of these are HBM-bound: coalesced float4 rows, fp32 math, deterministic reductions.
#include "common.cuh"
#include "gn_block.cuh"
#include "kernels.h"

namespace {

// ------------------------------------------------------------------ stem conv 7x7 stride 2, Cin = 1
// model/resnet.py:219 + utils.py:98-123: SAME padding of an even extent with k=7,s=2 is (2,3).
// CTA = one output row of one image; the 49x64 standardised filter bank sits in shared memory for the whole row, the
// input window (7 rows x 69 columns) of each 32-pixel segment is staged in shared memory.  Thread = (output channel,
// group of 8 pixels): filter taps are conflict-free reads, the input pixel is a warp-wide broadcast, stores are 256-byte
// rows of the NHWC output.
__global__ void __launch_bounds__(256) stem_conv_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                        float* __restrict__ raw1, ImgGeom g) {
    __shared__ float ws[49 * 64];
    __shared__ float win[7][72];
    const int b = blockIdx.y, oy = blockIdx.x;
    const int H = g.img_hw[2 * b], W = g.img_hw[2 * b + 1];
    const int H1 = H >> 1, W1 = W >> 1;
    if (oy >= H1) return;
    for (int i = threadIdx.x; i < 49 * 64; i += 256) ws[i] = w[i];
    const float* im = img + g.img_off[b];
    float* out = raw1 + ((size_t)(g.img_off[b] >> 2) + (size_t)oy * W1) * 64;
    const int c = threadIdx.x & 63, pg = threadIdx.x >> 6;
    for (int x0 = 0; x0 < W1; x0 += 32) {
        __syncthreads();
        for (int i = threadIdx.x; i < 7 * 69; i += 256) {          // input columns 2*x0-2 .. 2*x0+66, rows 2*oy-2 .. 2*oy+4
            const int ky = i / 69, cx = i - ky * 69;
            const int iy = 2 * oy + ky - 2, ix = 2 * x0 + cx - 2;
            win[ky][cx] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(im + (size_t)iy * W + ix) : 0.f;
        }
        __syncthreads();
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ky = 0; ky < 7; ++ky) {
#pragma unroll
            for (int kx = 0; kx < 7; ++kx) {
                const float wt = ws[(ky * 7 + kx) * 64 + c];
#pragma unroll
                for (int px = 0; px < 8; ++px) acc[px] = fmaf(win[ky][2 * (pg * 8 + px) + kx], wt, acc[px]);
            }
        }
#pragma unroll
        for (int px = 0; px < 8; ++px) {
            const int ox = x0 + pg * 8 + px;
            if (ox < W1) out[(size_t)ox * 64 + c] = acc[px];
        }
    }
}

// v = hi + lo with hi = bf16(v), lo = bf16(v - hi): ~16 mantissa bits through two bf16 tensor-core operands
TX_DEVINL void store_split4(bf16* hi, bf16* lo, const float* r) {
    float h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { h[i] = __bfloat162float(__float2bfloat16_rn(r[i])); l[i] = r[i] - h[i]; }
    st4(hi, make_float4(h[0], h[1], h[2], h[3]));
    st4(lo, make_float4(l[0], l[1], l[2], l[3]));
}

TX_DEVINL uint32_t pack2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

// chunks an image of `npix` pixels is split into: depends on the image alone (batch-composition independent)
TX_DEVINL int image_chunks(int npix, int nchunk) { return max(1, min(nchunk, (npix + 63) / 64)); }

// ------------------------------------------------------------------ GroupNorm statistics
// partial[b][chunk][32][2] (double) then stats[b][32] = (mean, rstd).  Fixed summation order.
template <int C>
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ raw, const int* __restrict__ img_off,
                                                       int level, int nchunk, double* __restrict__ partial) {
    constexpr int C4 = C / 4;             // float4 columns
    constexpr int RPP = 256 / C4;         // rows per pass
    constexpr int CPG = C / 32;
    __shared__ double s_sum[RPP * C];     // RPP*C == 1024
    __shared__ double s_sq[RPP * C];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int p0 = img_off[b] >> (2 * level), p1 = img_off[b + 1] >> (2 * level);
    const int npix = p1 - p0;
    const int used = image_chunks(npix, nchunk);
    if (chunk >= used) return;
    const int per = (npix + used - 1) / used;
    const int lo = p0 + chunk * per, hi = min(p1, lo + per);
    const int c4 = threadIdx.x % C4, rl = threadIdx.x / C4;
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
    // 4 rows in flight per thread (same summation order as one row at a time: the partial sums are added in row order)
    int p = lo + rl;
    for (; p + 3 * RPP < hi; p += 4 * RPP) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ld4(raw + (size_t)(p + u * RPP) * C + c4 * 4);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            s[0] += v[u].x; s[1] += v[u].y; s[2] += v[u].z; s[3] += v[u].w;
            q[0] = fmaf(v[u].x, v[u].x, q[0]); q[1] = fmaf(v[u].y, v[u].y, q[1]); q[2] = fmaf(v[u].z, v[u].z, q[2]); q[3] = fmaf(v[u].w, v[u].w, q[3]);
        }
    }
    for (; p < hi; p += RPP) {
        float4 v = ld4(raw + (size_t)p * C + c4 * 4);
        s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
        q[0] = fmaf(v.x, v.x, q[0]); q[1] = fmaf(v.y, v.y, q[1]); q[2] = fmaf(v.z, v.z, q[2]); q[3] = fmaf(v.w, v.w, q[3]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        s_sum[rl * C + c4 * 4 + i] = (double)s[i];
        s_sq[rl * C + c4 * 4 + i] = (double)q[i];
    }
    __syncthreads();
    // per-channel totals into row 0
    for (int c = threadIdx.x; c < C; c += 256) {
        double a = 0.0, bq = 0.0;
        for (int r = 0; r < RPP; ++r) { a += s_sum[r * C + c]; bq += s_sq[r * C + c]; }
        s_sum[c] = a; s_sq[c] = bq;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        double a = 0.0, bq = 0.0;
        for (int i = 0; i < CPG; ++i) { a += s_sum[threadIdx.x * CPG + i]; bq += s_sq[threadIdx.x * CPG + i]; }
        double* out = partial + (((size_t)b * nchunk + chunk) * 32 + threadIdx.x) * 2;
        out[0] = a; out[1] = bq;
    }
}

__global__ void gn_finalize_kernel(const double* __restrict__ partial, const int* __restrict__ img_off, int level,
                                   int nchunk, int cpg, float* __restrict__ stats) {
    const int b = blockIdx.x, g = threadIdx.x;
    double a = 0.0, q = 0.0;
    const int used = image_chunks((img_off[b + 1] >> (2 * level)) - (img_off[b] >> (2 * level)), nchunk);
    for (int c = 0; c < used; ++c) {
        const double* in = partial + (((size_t)b * nchunk + c) * 32 + g) * 2;
        a += in[0]; q += in[1];
    }
    const double n = (double)((img_off[b + 1] >> (2 * level)) - (img_off[b] >> (2 * level))) * cpg;
    const double mean = a / n;
    double var = q / n - mean * mean;       // biased variance (F.group_norm)
    if (var < 0.0) var = 0.0;
    stats[((size_t)b * 32 + g) * 2 + 0] = (float)mean;
    stats[((size_t)b * 32 + g) * 2 + 1] = (float)(1.0 / sqrt(var + 1e-5));
}

// ------------------------------------------------------------------ GroupNorm statistics, bf16 tier: block partial sums
// Stand-alone producer of the partials of gn_block.cuh (the convolution GEMM's epilogue is the other one): warp = one
// (32-row block of an image, 32-channel chunk), coalesced 16-byte loads in exactly the lane mapping of the GEMM epilogue.
// slot = (first row of the block >> 5) + image index; slots that belong to no image are skipped.
__global__ void __launch_bounds__(256) gn_stats_blocks_kernel(const float* __restrict__ raw, const int* __restrict__ img_off, int nimg,
                                                              int level, int C, int nslots, float* __restrict__ part) {
    const int chunks = C >> 5;
    const long w = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int slot = (int)(w / chunks), chunk = (int)(w - (long)slot * chunks);
    if (slot >= nslots) return;
    // largest b with (row0_b >> 5) + b <= slot
    int lo = 0, hi = nimg - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (((img_off[mid] >> (2 * level)) >> 5) + mid <= slot) lo = mid; else hi = mid - 1;
    }
    const int b = lo;
    const int p0 = img_off[b] >> (2 * level), p1 = img_off[b + 1] >> (2 * level);
    const int blk = slot - ((p0 >> 5) + b);
    const int r0 = p0 + blk * 32;
    if (blk < 0 || r0 >= p1) return;
    const int rows_ok = min(32, p1 - r0);
    const int rr = lane >> 3, ch = lane & 7;
    const float* src = raw + (size_t)r0 * C + chunk * 32 + ch * 4;
    float4 f[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int row = it * 4 + rr;
        f[it] = row < rows_ok ? ld4(src + (size_t)row * C) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int it = 0; it < 8; ++it)
        if (it * 4 + rr < rows_ok) gn_block_acc(s, q, f[it]);
    gn_block_finish(s, q, lane, C >> 5, chunk * 32, part + (size_t)slot * 64);
}

// stats[b][g] = (mean, rstd) from the image's block partials, added in double in a fixed order: eight interleaved slices (warp w adds
// blocks w, w + 8, ...), then the slices in order.  (A level-1 image of 64x384 has 192 blocks: one thread per group was 0.1 ms.)
__global__ void __launch_bounds__(256) gn_finalize_blocks_kernel(const float* __restrict__ part, const int* __restrict__ img_off, int level,
                                                                 int cpg, float* __restrict__ stats) {
    __shared__ double sa[8][32], sq[8][32];
    const int b = blockIdx.x, g = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int p0 = img_off[b] >> (2 * level), p1 = img_off[b + 1] >> (2 * level);
    const int nblk = (p1 - p0 + 31) >> 5;
    const float2* in = reinterpret_cast<const float2*>(part) + ((size_t)((p0 >> 5) + b)) * 32 + g;
    double a = 0.0, q = 0.0;
    int k = w;
    for (; k + 24 < nblk; k += 32) {
        float2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = in[(size_t)(k + 8 * u) * 32];
#pragma unroll
        for (int u = 0; u < 4; ++u) { a += (double)v[u].x; q += (double)v[u].y; }
    }
    for (; k < nblk; k += 8) { const float2 v = in[(size_t)k * 32]; a += (double)v.x; q += (double)v.y; }
    sa[w][g] = a; sq[w][g] = q;
    __syncthreads();
    if (w == 0) {
        a = 0.0; q = 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) { a += sa[u][g]; q += sq[u][g]; }
        const double n = (double)(p1 - p0) * cpg;
        const double mean = a / n;
        double var = q / n - mean * mean;       // biased variance (F.group_norm)
        if (var < 0.0) var = 0.0;
        stats[((size_t)b * 32 + g) * 2 + 0] = (float)mean;
        stats[((size_t)b * 32 + g) * 2 + 1] = (float)(1.0 / sqrt(var + 1e-5));
    }
}

// ------------------------------------------------------------------ GroupNorm apply (+residual, +ReLU)
__global__ void __launch_bounds__(256) gn_apply_kernel(GnApplyArgs a, const int* __restrict__ img_off, int nchunk) {
    const int C = a.C, C4 = C / 4, cpg = C / 32;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int p0 = img_off[b] >> (2 * a.level), p1 = img_off[b + 1] >> (2 * a.level);
    const int used = image_chunks(p1 - p0, nchunk);
    if (chunk >= used) return;
    const int per = (p1 - p0 + used - 1) / used;
    const int lo = p0 + chunk * per, hi = min(p1, lo + per);
    const int rpp = 256 / min(C4, 256);
    const int c4 = threadIdx.x % C4, rl = threadIdx.x / C4;     // C4 <= 256
    const int c = c4 * 4;
    float mean[4], sc[4], be[4], mean2[4], sc2[4], be2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gi = (c + i) / cpg;
        mean[i] = a.stats[((size_t)b * 32 + gi) * 2];
        sc[i] = a.stats[((size_t)b * 32 + gi) * 2 + 1] * a.gamma[c + i];
        be[i] = a.beta[c + i];
        if (a.raw2) {
            mean2[i] = a.stats2[((size_t)b * 32 + gi) * 2];
            sc2[i] = a.stats2[((size_t)b * 32 + gi) * 2 + 1] * a.gamma2[c + i];
            be2[i] = a.beta2[c + i];
        } else { mean2[i] = sc2[i] = be2[i] = 0.f; }
    }
    for (int p = lo + rl; p < hi; p += rpp) {
        const size_t o = (size_t)p * C + c;
        float4 v = ld4(a.raw + o);
        float r[4] = {(v.x - mean[0]) * sc[0] + be[0], (v.y - mean[1]) * sc[1] + be[1],
                      (v.z - mean[2]) * sc[2] + be[2], (v.w - mean[3]) * sc[3] + be[3]};
        if (a.raw2) {
            float4 u = ld4(a.raw2 + o);
            r[0] += (u.x - mean2[0]) * sc2[0] + be2[0]; r[1] += (u.y - mean2[1]) * sc2[1] + be2[1];
            r[2] += (u.z - mean2[2]) * sc2[2] + be2[2]; r[3] += (u.w - mean2[3]) * sc2[3] + be2[3];
        }
        if (a.res) {
            float4 u = ld4(a.res + o);
            r[0] += u.x; r[1] += u.y; r[2] += u.z; r[3] += u.w;
        }
        if (a.res_hi) {
            float4 u = ld4(reinterpret_cast<const bf16*>(a.res_hi) + o), w = ld4(reinterpret_cast<const bf16*>(a.res_lo) + o);
            r[0] += u.x + w.x; r[1] += u.y + w.y; r[2] += u.z + w.z; r[3] += u.w + w.w;
        }
        if (a.relu) {
#pragma unroll
            for (int i = 0; i < 4; ++i) r[i] = fmaxf(r[i], 0.f);
        }
        if (a.out) st4(a.out + o, make_float4(r[0], r[1], r[2], r[3]));
        else store_split4(reinterpret_cast<bf16*>(a.out_hi) + o, reinterpret_cast<bf16*>(a.out_lo) + o, r);
    }
}

// bf16 tier (split-bf16 output, one raw input, optional split-bf16 residual): eight channels per thread, so that every access
// of the bf16 planes is 16 bytes per lane (the four-channel kernel above reads / writes them 8 bytes at a time: the residual-pair
// variant ran at 62-80 % of the bytes-per-second of the two-raw-inputs variant on the same byte count).  Two rows in flight per
// thread, the residual kept packed until it is used.  Same arithmetic per element as gn_apply_kernel.
template <bool RES>
__global__ void __launch_bounds__(256, 3) gn_apply8_kernel(GnApplyArgs a, const int* __restrict__ img_off, int nchunk) {
    const int C = a.C, C8 = C / 8, cpg = C / 32;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int p0 = img_off[b] >> (2 * a.level), p1 = img_off[b + 1] >> (2 * a.level);
    const int used = image_chunks(p1 - p0, nchunk);
    if (chunk >= used) return;
    const int per = (p1 - p0 + used - 1) / used;
    const int lo = p0 + chunk * per, hi = min(p1, lo + per);
    const int rpp = 256 / C8;                                   // C8 <= 128
    const int c8 = threadIdx.x % C8, rl = threadIdx.x / C8;
    const int c = c8 * 8;
    float mean[8], sc[8], be[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int gi = (c + i) / cpg;
        mean[i] = a.stats[((size_t)b * 32 + gi) * 2];
        sc[i] = a.stats[((size_t)b * 32 + gi) * 2 + 1] * a.gamma[c + i];
        be[i] = a.beta[c + i];
    }
    const bf16* res_hi = reinterpret_cast<const bf16*>(a.res_hi);
    const bf16* res_lo = reinterpret_cast<const bf16*>(a.res_lo);
    bf16* out_hi = reinterpret_cast<bf16*>(a.out_hi);
    bf16* out_lo = reinterpret_cast<bf16*>(a.out_lo);
    const bool relu = a.relu;
    auto one = [&](const float4 va, const float4 vb, const uint4 rh, const uint4 rl2, size_t o) {
        const float v[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
        float r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = (v[i] - mean[i]) * sc[i] + be[i];
        if (RES) {
            const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&rh);
            const __nv_bfloat162* ll = reinterpret_cast<const __nv_bfloat162*>(&rl2);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 fh = __bfloat1622float2(hh[i]), fl = __bfloat1622float2(ll[i]);
                r[2 * i] += fh.x + fl.x; r[2 * i + 1] += fh.y + fl.y;
            }
        }
        if (relu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = fmaxf(r[i], 0.f);
        }
        float h[8], l[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { h[i] = __bfloat162float(__float2bfloat16_rn(r[i])); l[i] = r[i] - h[i]; }
        uint4 ph, pl;
        ph.x = pack2(h[0], h[1]); ph.y = pack2(h[2], h[3]); ph.z = pack2(h[4], h[5]); ph.w = pack2(h[6], h[7]);
        pl.x = pack2(l[0], l[1]); pl.y = pack2(l[2], l[3]); pl.z = pack2(l[4], l[5]); pl.w = pack2(l[6], l[7]);
        *reinterpret_cast<uint4*>(out_hi + o) = ph;
        *reinterpret_cast<uint4*>(out_lo + o) = pl;
    };
    const uint4 z = make_uint4(0, 0, 0, 0);
    int p = lo + rl;
    for (; p + rpp < hi; p += 2 * rpp) {                      // two rows in flight
        const size_t o0 = (size_t)p * C + c, o1 = (size_t)(p + rpp) * C + c;
        const float4 a0 = ld4(a.raw + o0), b0 = ld4(a.raw + o0 + 4), a1 = ld4(a.raw + o1), b1 = ld4(a.raw + o1 + 4);
        uint4 h0 = z, l0 = z, h1 = z, l1 = z;
        if (RES) {
            h0 = *reinterpret_cast<const uint4*>(res_hi + o0); l0 = *reinterpret_cast<const uint4*>(res_lo + o0);
            h1 = *reinterpret_cast<const uint4*>(res_hi + o1); l1 = *reinterpret_cast<const uint4*>(res_lo + o1);
        }
        one(a0, b0, h0, l0, o0);
        one(a1, b1, h1, l1, o1);
    }
    for (; p < hi; p += rpp) {
        const size_t o0 = (size_t)p * C + c;
        const float4 a0 = ld4(a.raw + o0), b0 = ld4(a.raw + o0 + 4);
        uint4 h0 = z, l0 = z;
        if (RES) { h0 = *reinterpret_cast<const uint4*>(res_hi + o0); l0 = *reinterpret_cast<const uint4*>(res_lo + o0); }
        one(a0, b0, h0, l0, o0);
    }
}

// ------------------------------------------------------------------ stem: GN + ReLU + maxpool 3x3 s2
// model/resnet.py:69-79: SAME pad (0,1) filled with -inf => out-of-range taps are skipped.
__global__ void __launch_bounds__(256) gn_apply_maxpool_kernel(const float* __restrict__ raw1, const float* __restrict__ stats,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               float* __restrict__ out2, bf16* __restrict__ out_hi,
                                                               bf16* __restrict__ out_lo, ImgGeom g, int total_p2) {
    const int c4 = threadIdx.x & 15, rl = threadIdx.x >> 4;       // 16 float4 columns (C = 64), 16 pixels per block pass
    const int p = blockIdx.x * 16 + rl;
    if (p >= total_p2) return;
    const int b = find_image(g.img_off, g.nimg, 2, p);
    const int H1 = g.img_hw[2 * b] >> 1, W1 = g.img_hw[2 * b + 1] >> 1, W2 = W1 >> 1;
    const int local = p - (g.img_off[b] >> 4);
    const int oy = local / W2, ox = local - oy * W2;
    const float* in = raw1 + (size_t)(g.img_off[b] >> 2) * 64;
    const int c = c4 * 4;
    float mean[4], sc[4], be[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gi = (c + i) >> 1;     // 64 channels / 32 groups
        mean[i] = stats[((size_t)b * 32 + gi) * 2];
        sc[i] = stats[((size_t)b * 32 + gi) * 2 + 1] * gamma[c + i];
        be[i] = beta[c + i];
    }
    float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = 2 * oy + ky;
        if (iy >= H1) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = 2 * ox + kx;
            if (ix >= W1) continue;
            float4 v = ld4(in + ((size_t)iy * W1 + ix) * 64 + c);
            m[0] = fmaxf(m[0], fmaxf((v.x - mean[0]) * sc[0] + be[0], 0.f));
            m[1] = fmaxf(m[1], fmaxf((v.y - mean[1]) * sc[1] + be[1], 0.f));
            m[2] = fmaxf(m[2], fmaxf((v.z - mean[2]) * sc[2] + be[2], 0.f));
            m[3] = fmaxf(m[3], fmaxf((v.w - mean[3]) * sc[3] + be[3], 0.f));
        }
    }
    if (out2) st4(out2 + (size_t)p * 64 + c, make_float4(m[0], m[1], m[2], m[3]));
    else store_split4(out_hi + (size_t)p * 64 + c, out_lo + (size_t)p * 64 + c, m);
}

// ------------------------------------------------------------------ explicit im2col for the tcgen05 (bf16x3) conv path
// thread = 8 channels (16 B) of one tap of one output pixel; both halves of the split pair are copied.
__global__ void __launch_bounds__(256) im2col_split_kernel(const bf16* __restrict__ in_hi, const bf16* __restrict__ in_lo,
                                                           bf16* __restrict__ out_hi, bf16* __restrict__ out_lo, ConvGather cg,
                                                           long total_items, int c8, int taps) {
    const long item = (long)blockIdx.x * 256 + threadIdx.x;
    if (item >= total_items) return;
    const int cc = (int)(item % c8);
    const long rest = item / c8;
    const int tap = (int)(rest % taps);
    const long m = rest / taps;
    const int b = find_image(cg.img_off, cg.nimg, cg.lout, (int)m);
    const int H = cg.img_hw[2 * b], W = cg.img_hw[2 * b + 1];
    const int wo = W >> cg.lout, hin = H >> cg.lin, win = W >> cg.lin;
    const int local = (int)m - (cg.img_off[b] >> (2 * cg.lout));
    const int oy = local / wo, ox = local - oy * wo;
    const int ky = tap / cg.ksz, kx = tap - ky * cg.ksz;
    const int iy = oy * cg.stride + ky - cg.pad, ix = ox * cg.stride + kx - cg.pad;
    uint4 vh = make_uint4(0, 0, 0, 0), vl = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < hin && ix >= 0 && ix < win) {
        const size_t src = ((size_t)(cg.img_off[b] >> (2 * cg.lin)) + (size_t)iy * win + ix) * cg.cin + cc * 8;
        vh = *reinterpret_cast<const uint4*>(in_hi + src);
        vl = *reinterpret_cast<const uint4*>(in_lo + src);
    }
    const size_t dst = ((size_t)m * taps + tap) * cg.cin + cc * 8;
    *reinterpret_cast<uint4*>(out_hi + dst) = vh;
    *reinterpret_cast<uint4*>(out_lo + dst) = vl;
}

// ------------------------------------------------------------------ patch variant: 16x16/s16 im2col (model/encoder.py:22-27)
__global__ void __launch_bounds__(256) im2col_patch_kernel(const float* __restrict__ img, float* __restrict__ cols, ImgGeom g,
                                                           int total_p4) {
    // one warp per patch pair... simple mapping: thread = (patch, float4 of the 256-element patch row)
    const int p = blockIdx.x * 4 + (threadIdx.x >> 6);
    const int q = threadIdx.x & 63;                // float4 index in [0,64): ky = q/4, kx4 = q%4
    if (p >= total_p4) return;
    const int b = find_image(g.img_off, g.nimg, 4, p);
    const int W = g.img_hw[2 * b + 1], w4 = W >> 4;
    const int local = p - (g.img_off[b] >> 8);
    const int r = local / w4, cc = local - r * w4;
    const int ky = q >> 2, kx = (q & 3) * 4;
    float4 v = ld4(img + g.img_off[b] + (size_t)(r * 16 + ky) * W + cc * 16 + kx);
    st4(cols + (size_t)p * 256 + q * 4, v);
}

// ------------------------------------------------------------------ cls + positional embedding (model/encoder.py:128-143)
// x0[tok_off[b]] = cls + pos[0];  x0[tok_off[b] + 1 + r*w + c] = proj[pixel] + pos[r*63 + c + 1]
__global__ void __launch_bounds__(256) assemble_tokens_kernel(const float* __restrict__ proj, const float* __restrict__ cls,
                                                              const float* __restrict__ pos, float* __restrict__ x0, ImgGeom g,
                                                              const int* __restrict__ tok_off, int total_tok) {
    const int t = blockIdx.x * 4 + (threadIdx.x >> 6);
    const int q = (threadIdx.x & 63) * 4;
    if (t >= total_tok) return;
    // tok_off[b] = (img_off[b] >> 8) + b  => find b by binary search on tok_off
    int lo = 0, hi = g.nimg - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (tok_off[mid] <= t) lo = mid; else hi = mid - 1;
    }
    const int b = lo, local = t - tok_off[b];
    float4 v, pe;
    if (local == 0) {
        v = ld4(cls + q);
        pe = ld4(pos + q);
    } else {
        const int w4 = g.img_hw[2 * b + 1] >> 4;
        const int r = (local - 1) / w4, c = (local - 1) - r * w4;
        v = ld4(proj + ((size_t)(g.img_off[b] >> 8) + (local - 1)) * 256 + q);
        pe = ld4(pos + (size_t)(r * 63 + c + 1) * 256 + q);
    }
    st4(x0 + (size_t)t * 256 + q, make_float4(v.x + pe.x, v.y + pe.y, v.z + pe.z, v.w + pe.w));
}

}  // namespace

cudaError_t launch_stem_conv(const float* img, const float* w, float* raw1, const int* img_off, const int* img_hw,
                             int nimg, int total_p1, cudaStream_t st) {
    ImgGeom g{img_off, img_hw, nimg};
    (void)total_p1;
    dim3 grid(80, nimg);             // output rows per image: H/2 <= 80 (H <= 160); rows past an image's height exit at once
    stem_conv_kernel<<<grid, 256, 0, st>>>(img, w, raw1, g);
    return cudaGetLastError();
}

cudaError_t launch_gn_stats(const float* raw, int C, int level, const int* img_off, int nimg, int nchunk,
                            double* partial, float* stats, cudaStream_t st) {
    dim3 grid(nchunk, nimg);
    switch (C) {
        case 64: gn_stats_kernel<64><<<grid, 256, 0, st>>>(raw, img_off, level, nchunk, partial); break;
        case 128: gn_stats_kernel<128><<<grid, 256, 0, st>>>(raw, img_off, level, nchunk, partial); break;
        case 256: gn_stats_kernel<256><<<grid, 256, 0, st>>>(raw, img_off, level, nchunk, partial); break;
        case 512: gn_stats_kernel<512><<<grid, 256, 0, st>>>(raw, img_off, level, nchunk, partial); break;
        case 1024: gn_stats_kernel<1024><<<grid, 256, 0, st>>>(raw, img_off, level, nchunk, partial); break;
        default: return cudaErrorInvalidValue;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    gn_finalize_kernel<<<nimg, 32, 0, st>>>(partial, img_off, level, nchunk, C / 32, stats);
    return cudaGetLastError();
}

cudaError_t launch_gn_stats_blocks(const float* raw, int C, int level, const int* img_off, int nimg, long total_rows, float* part,
                                   cudaStream_t st) {
    if (C % 64 != 0 || C > 1024 || ((C >> 5) & ((C >> 5) - 1))) return cudaErrorInvalidValue;
    const long nslots = (total_rows >> 5) + nimg + 1;
    const long warps = nslots * (C >> 5);
    gn_stats_blocks_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(raw, img_off, nimg, level, C, (int)nslots, part);
    return cudaGetLastError();
}

cudaError_t launch_gn_finalize_blocks(const float* part, int C, int level, const int* img_off, int nimg, float* stats, cudaStream_t st) {
    gn_finalize_blocks_kernel<<<nimg, 256, 0, st>>>(part, img_off, level, C / 32, stats);
    return cudaGetLastError();
}

cudaError_t launch_gn_apply(const GnApplyArgs& a, const int* img_off, int nimg, int nchunk, cudaStream_t st) {
    if (a.C % 4 != 0 || a.C / 4 > 256) return cudaErrorInvalidValue;
    dim3 grid(nchunk, nimg);
    // split-bf16 output, one raw input (the bf16 tier's backbone away from the down-sampling blocks): eight channels per thread
    if (a.out_hi && !a.out && !a.res && !a.raw2 && a.C % 64 == 0 && a.C / 8 <= 128) {
        if (a.res_hi) gn_apply8_kernel<true><<<grid, 256, 0, st>>>(a, img_off, nchunk);
        else gn_apply8_kernel<false><<<grid, 256, 0, st>>>(a, img_off, nchunk);
    } else gn_apply_kernel<<<grid, 256, 0, st>>>(a, img_off, nchunk);
    return cudaGetLastError();
}

cudaError_t launch_gn_apply_maxpool(const float* raw1, const float* stats, const float* gamma, const float* beta,
                                    float* out2, void* out_hi, void* out_lo, const int* img_off, const int* img_hw, int nimg,
                                    int total_p2, cudaStream_t st) {
    ImgGeom g{img_off, img_hw, nimg};
    gn_apply_maxpool_kernel<<<(total_p2 + 15) / 16, 256, 0, st>>>(raw1, stats, gamma, beta, out2, (bf16*)out_hi, (bf16*)out_lo, g, total_p2);
    return cudaGetLastError();
}

cudaError_t launch_im2col_split(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, const ConvGather& cg,
                                long total_out_pixels, cudaStream_t st) {
    if (cg.cin % 8 != 0) return cudaErrorInvalidValue;
    const int c8 = cg.cin / 8, taps = cg.ksz * cg.ksz;
    const long items = total_out_pixels * taps * c8;
    if (items <= 0) return cudaSuccess;
    const long blocks = (items + 255) / 256;
    if (blocks > 0x7fffffffL) return cudaErrorInvalidValue;
    im2col_split_kernel<<<(unsigned)blocks, 256, 0, st>>>((const bf16*)in_hi, (const bf16*)in_lo, (bf16*)out_hi, (bf16*)out_lo, cg,
                                                         items, c8, taps);
    return cudaGetLastError();
}

cudaError_t launch_im2col_patch(const float* img, float* cols, const int* img_off, const int* img_hw, int nimg,
                                int total_p4, cudaStream_t st) {
    ImgGeom g{img_off, img_hw, nimg};
    im2col_patch_kernel<<<(total_p4 + 3) / 4, 256, 0, st>>>(img, cols, g, total_p4);
    return cudaGetLastError();
}

cudaError_t launch_assemble_tokens(const float* proj, const float* cls, const float* pos, float* x0, const int* img_off,
                                   const int* img_hw, const int* tok_off, int nimg, int total_tok, cudaStream_t st) {
    ImgGeom g{img_off, img_hw, nimg};
    assemble_tokens_kernel<<<(total_tok + 3) / 4, 256, 0, st>>>(proj, cls, pos, x0, g, tok_off, total_tok);
    return cudaGetLastError();
}
