// FFMA GEMM  C[M,N] = epi(A[M,K] . W[N,K]^T)  -- the fp32 parity tier of every GEMM / convolution
// on the path (tcgen05 has no fp32 kind, so the 1e-4 tier needs real FFMA), and the
// fallback for shapes the tcgen05 kernel does not take.  Both operands are K-major (PyTorch Linear
// weights are [out,in]; NHWC activations are [pixel, channel]).
//
// Convolutions (model/resnet.py:58-66) are the same kernel with an implicit-im2col A operand:
// row m = output pixel of a ragged NHWC batch, k = (ky, kx, c); the TF-"SAME" padding of
// utils.py:93-123 becomes an in-range test on the gathered input pixel.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int BK = 16;

struct Params {
    const void* A; const void* W; void* C;
    int M, N, K, lda, ldw, ldc;
    const float* bias; const float* res; int ldres;
    ConvGather cv;
};

template <typename TA, typename TC, int EPI, int BM, int BN, bool GATHER>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const Params p) {
    constexpr int TM = BM / 16, TN = BN / 16;       // micro-tile per thread
    constexpr int HM = TM / 4, HN = TN / 4;         // float4 groups per thread (rows / cols)
    constexpr int LA = BM * 4 / 256, LW = BN * 4 / 256;   // float4 loads per thread per k-tile
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Ws[2][BK][BN + 4];

    const TA* __restrict__ A = reinterpret_cast<const TA*>(p.A);
    const TA* __restrict__ W = reinterpret_cast<const TA*>(p.W);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

    // ---- loader state: each thread owns LA rows of the A tile and LW rows of the W tile, one k-quad
    const int kq = tid & 3;
    int a_row[LA];
    bool a_ok[LA];
    const float* a_base[LA];     // GATHER: image base pointer
    int a_oy[LA], a_ox[LA], a_hin[LA], a_win[LA];
    const TA* a_ptr[LA];
#pragma unroll
    for (int i = 0; i < LA; ++i) {
        a_row[i] = (tid >> 2) + i * 64;
        int m = m0 + a_row[i];
        a_ok[i] = m < p.M;
        a_ptr[i] = A;
        a_base[i] = nullptr; a_oy[i] = a_ox[i] = a_hin[i] = a_win[i] = 0;
        if (GATHER) {
            if (a_ok[i]) {
                int b = find_image(p.cv.img_off, p.cv.nimg, p.cv.lout, m);
                int H = p.cv.img_hw[2 * b], Wd = p.cv.img_hw[2 * b + 1];
                int wo = Wd >> p.cv.lout;
                int local = m - (p.cv.img_off[b] >> (2 * p.cv.lout));
                a_oy[i] = local / wo; a_ox[i] = local - a_oy[i] * wo;
                a_hin[i] = H >> p.cv.lin; a_win[i] = Wd >> p.cv.lin;
                a_base[i] = reinterpret_cast<const float*>(p.A) + (size_t)(p.cv.img_off[b] >> (2 * p.cv.lin)) * p.cv.cin;
            }
        } else {
            a_ptr[i] = A + (size_t)(a_ok[i] ? m : 0) * p.lda;
        }
    }
    int w_row[LW]; bool w_ok[LW]; const TA* w_ptr[LW];
#pragma unroll
    for (int i = 0; i < LW; ++i) {
        w_row[i] = (tid >> 2) + i * 64;
        int n = n0 + w_row[i];
        w_ok[i] = n < p.N;
        w_ptr[i] = W + (size_t)(w_ok[i] ? n : 0) * p.ldw;
    }

    float4 ra[LA], rw[LW];
    auto load_tile = [&](int kt) {
        const int k = kt * BK + kq * 4;
#pragma unroll
        for (int i = 0; i < LA; ++i) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (GATHER) {
                if (a_ok[i]) {
                    int tap = k / p.cv.cin, c = k - tap * p.cv.cin;
                    int ky = tap / p.cv.ksz, kx = tap - ky * p.cv.ksz;
                    int iy = a_oy[i] * p.cv.stride + ky - p.cv.pad, ix = a_ox[i] * p.cv.stride + kx - p.cv.pad;
                    if (iy >= 0 && iy < a_hin[i] && ix >= 0 && ix < a_win[i])
                        v = ld4(a_base[i] + ((size_t)iy * a_win[i] + ix) * p.cv.cin + c);
                }
            } else {
                if (a_ok[i]) v = ld4(a_ptr[i] + k);
            }
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < LW; ++i) rw[i] = w_ok[i] ? ld4(w_ptr[i] + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int i = 0; i < LA; ++i) {
            As[buf][kq * 4 + 0][a_row[i]] = ra[i].x; As[buf][kq * 4 + 1][a_row[i]] = ra[i].y;
            As[buf][kq * 4 + 2][a_row[i]] = ra[i].z; As[buf][kq * 4 + 3][a_row[i]] = ra[i].w;
        }
#pragma unroll
        for (int i = 0; i < LW; ++i) {
            Ws[buf][kq * 4 + 0][w_row[i]] = rw[i].x; Ws[buf][kq * 4 + 1][w_row[i]] = rw[i].y;
            Ws[buf][kq * 4 + 2][w_row[i]] = rw[i].z; Ws[buf][kq * 4 + 3][w_row[i]] = rw[i].w;
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int nk = p.K / BK;
    load_tile(0);
    store_tile(0);
    __syncthreads();
    int cur = 0;
    for (int kt = 0; kt < nk; ++kt) {
        if (kt + 1 < nk) load_tile(kt + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int h = 0; h < HM; ++h) {
                float4 v = *reinterpret_cast<const float4*>(&As[cur][k][h * (BM / 2) + ty * 4]);
                a[h * 4 + 0] = v.x; a[h * 4 + 1] = v.y; a[h * 4 + 2] = v.z; a[h * 4 + 3] = v.w;
            }
#pragma unroll
            for (int h = 0; h < HN; ++h) {
                float4 v = *reinterpret_cast<const float4*>(&Ws[cur][k][h * (BN / 2) + tx * 4]);
                b[h * 4 + 0] = v.x; b[h * 4 + 1] = v.y; b[h * 4 + 2] = v.z; b[h * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            store_tile(cur ^ 1);
            __syncthreads();
            cur ^= 1;
        }
    }

    // ---- epilogue.  Thread owns rows {h*(BM/2) + ty*4 + r} and column quads {h*(BN/2) + tx*4 .. +3}.
#pragma unroll
    for (int hm = 0; hm < HM; ++hm)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int m = m0 + hm * (BM / 2) + ty * 4 + r;
            if (m >= p.M) continue;
#pragma unroll
            for (int hn = 0; hn < HN; ++hn) {
                const int n = n0 + hn * (BN / 2) + tx * 4;
                if (n >= p.N) continue;     // N % 4 == 0: a quad is fully in or out
                float v0 = acc[hm * 4 + r][hn * 4 + 0], v1 = acc[hm * 4 + r][hn * 4 + 1];
                float v2 = acc[hm * 4 + r][hn * 4 + 2], v3 = acc[hm * 4 + r][hn * 4 + 3];
                if (p.bias) {
                    float4 bb = ld4(p.bias + n);
                    v0 += bb.x; v1 += bb.y; v2 += bb.z; v3 += bb.w;
                }
                if (EPI == EPI_STORE) {
                    st4(reinterpret_cast<TC*>(p.C) + (size_t)m * p.ldc + n, make_float4(v0, v1, v2, v3));
                } else if (EPI == EPI_BIAS_RES) {
                    float4 rr = ld4(p.res + (size_t)m * p.ldres + n);
                    st4(reinterpret_cast<float*>(p.C) + (size_t)m * p.ldc + n,
                        make_float4(v0 + rr.x, v1 + rr.y, v2 + rr.z, v3 + rr.w));
                } else if (EPI == EPI_GLU_RES) {
                    const int j = n >> 1;
                    float2 rr = *reinterpret_cast<const float2*>(p.res + (size_t)m * p.ldres + j);
                    st2(reinterpret_cast<float*>(p.C) + (size_t)m * p.ldc + j, v0 * sigmoidf_(v1) + rr.x,
                        v2 * sigmoidf_(v3) + rr.y);
                } else {   // EPI_GEGLU
                    const int j = n >> 1;
                    st2(reinterpret_cast<TA*>(p.C) + (size_t)m * p.ldc + j, v0 * gelu_erf(v1), v2 * gelu_erf(v3));
                }
            }
        }
}

template <typename TA, typename TC, int EPI, bool GATHER>
cudaError_t launch_t(const Params& p, cudaStream_t st) {
    // small problems get 64x64 tiles so that the grid still covers the 148 SMs
    const long tiles128 = (long)((p.M + 127) / 128) * ((p.N + 127) / 128);
    if (tiles128 >= 2 * 148) {
        dim3 grid((p.N + 127) / 128, (p.M + 127) / 128);
        gemm_simt_kernel<TA, TC, EPI, 128, 128, GATHER><<<grid, 256, 0, st>>>(p);
    } else {
        dim3 grid((p.N + 63) / 64, (p.M + 63) / 64);
        gemm_simt_kernel<TA, TC, EPI, 64, 64, GATHER><<<grid, 256, 0, st>>>(p);
    }
    return cudaGetLastError();
}

template <typename TA>
cudaError_t launch_a(const GemmArgs& g, const Params& p, cudaStream_t st) {
    switch (g.epi) {
        case EPI_STORE:
            if (g.dt_c == DT_F32) return launch_t<TA, float, EPI_STORE, false>(p, st);
            return launch_t<TA, bf16, EPI_STORE, false>(p, st);
        case EPI_GLU_RES: return launch_t<TA, float, EPI_GLU_RES, false>(p, st);
        case EPI_GEGLU: return launch_t<TA, float, EPI_GEGLU, false>(p, st);
        case EPI_BIAS_RES: return launch_t<TA, float, EPI_BIAS_RES, false>(p, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace

cudaError_t launch_gemm_simt(const GemmArgs& g, cudaStream_t st) {
    if (g.M <= 0) return cudaSuccess;
    if (g.K % BK != 0 || g.N % 4 != 0) return cudaErrorInvalidValue;
    Params p;
    p.A = g.A; p.W = g.W; p.C = g.C; p.M = g.M; p.N = g.N; p.K = g.K;
    p.lda = g.lda; p.ldw = g.ldw; p.ldc = g.ldc; p.bias = g.bias; p.res = g.res; p.ldres = g.ldres;
    if (g.conv) {
        if (g.dt_a != DT_F32 || g.epi != EPI_STORE || g.dt_c != DT_F32 || g.conv->cin % BK != 0)
            return cudaErrorInvalidValue;
        p.cv = *g.conv;
        return launch_t<float, float, EPI_STORE, true>(p, st);
    }
    p.cv = ConvGather{};
    if (g.dt_a == DT_F32) {
        if (g.epi == EPI_STORE && g.dt_c != DT_F32) return cudaErrorInvalidValue;
        return launch_a<float>(g, p, st);
    }
    return launch_a<bf16>(g, p, st);
}
