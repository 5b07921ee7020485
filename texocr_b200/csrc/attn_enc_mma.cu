// Encoder self-attention on the legacy tensor-core path (bf16 tier): ragged, non-causal, unmasked -- the
// `mask=None` call of model/encoder.py:147 -> model/attention.py:148-173 (energy * 0.125, softmax, . v).
//
// CTA = (image, head, 64-query tile), 4 warps x 16 query rows.  K and V of the (image, head) stream through shared
// memory in 64-key chunks (row stride 144 B: conflict-free ldmatrix); S = Q.K^T and O += P.V are mma.sync.m16n8k16
// (bf16 in, fp32 accumulate), the online softmax stays in fp32 registers (FlashAttention-2 register reuse: the S
// accumulator fragments are repacked as the A operand of the second MMA).  N <= 631 keys, d = 64.
// Masked / causal attention (teacher-forced decoder) keeps the SIMT kernel in attention.cu, which carries the reference's
// -FLT_MAX fill semantics.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int QT = 64, KT = 64, HD = 64, ROWB = 144;      // bytes per staged K/V row (64 bf16 + 16 B pad)
constexpr float SCALE = 0.125f;

TX_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
TX_DEVINL void ldsm_x4(uint32_t addr, uint32_t& d0, uint32_t& d1, uint32_t& d2, uint32_t& d3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(d0), "=r"(d1), "=r"(d2), "=r"(d3) : "r"(addr));
}
TX_DEVINL void ldsm_x4_t(uint32_t addr, uint32_t& d0, uint32_t& d1, uint32_t& d2, uint32_t& d3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(d0), "=r"(d1), "=r"(d2), "=r"(d3) : "r"(addr));
}
TX_DEVINL void mma_bf16(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
TX_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(128) attn_enc_mma_kernel(AttnVarlenArgs a) {
    __shared__ __align__(16) uint8_t Ks[KT * ROWB];
    __shared__ __align__(16) uint8_t Vs[KT * ROWB];
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int q0 = a.q_off[b], nq = a.q_len ? a.q_len[b] : a.q_off[b + 1] - q0;
    const int k0 = a.k_off[b], nk = a.k_len ? a.k_len[b] : a.k_off[b + 1] - k0;
    if (qt * QT >= nq || nk <= 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const bf16* __restrict__ Q = reinterpret_cast<const bf16*>(a.q);
    const bf16* __restrict__ K = reinterpret_cast<const bf16*>(a.k);
    const bf16* __restrict__ V = reinterpret_cast<const bf16*>(a.v);

    // Q fragments of this warp's 16 rows (rows g and g+8), pre-scaled by 0.125 (exact in bf16)
    const int r0 = qt * QT + warp * 16 + g, r1 = r0 + 8;
    uint32_t qa[4][4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int col = 16 * s + 8 * half + 2 * t;
            uint32_t w0 = 0, w1 = 0;
            if (r0 < nq) w0 = *reinterpret_cast<const uint32_t*>(Q + (size_t)(q0 + r0) * a.ldq + h * HD + col);
            if (r1 < nq) w1 = *reinterpret_cast<const uint32_t*>(Q + (size_t)(q0 + r1) * a.ldq + h * HD + col);
            const float2 f0 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w0));
            const float2 f1 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w1));
            qa[s][half * 2 + 0] = pack_bf16x2(f0.x * SCALE, f0.y * SCALE);     // a0 / a2: row g
            qa[s][half * 2 + 1] = pack_bf16x2(f1.x * SCALE, f1.y * SCALE);     // a1 / a3: row g+8
        }
    }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float o[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
    const uint32_t ks = smem_u32(Ks), vs = smem_u32(Vs);
    const int lm_r = lane & 7, lm_m = lane >> 3;

    for (int kc = 0; kc < nk; kc += KT) {
        __syncthreads();
        {   // K / V chunk -> smem (rows past nk zero-filled): 64 rows x 8 chunks of 16 B, 128 threads -> 4 passes each
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int idx = tid + i * 128, r = idx >> 3, c = idx & 7;
                uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
                if (kc + r < nk) {
                    kv = *reinterpret_cast<const uint4*>(K + (size_t)(k0 + kc + r) * a.ldk + h * HD + c * 8);
                    vv = *reinterpret_cast<const uint4*>(V + (size_t)(k0 + kc + r) * a.ldv + h * HD + c * 8);
                }
                *reinterpret_cast<uint4*>(Ks + r * ROWB + c * 16) = kv;
                *reinterpret_cast<uint4*>(Vs + r * ROWB + c * 16) = vv;
            }
        }
        __syncthreads();
        // ---- S = Q.K^T : 8 n-tiles of 8 keys
        float sc[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) {
                uint32_t b0, b1, b2, b3;
                ldsm_x4(ks + (8 * j + lm_r) * ROWB + (4 * s2 + lm_m) * 16, b0, b1, b2, b3);
                mma_bf16(sc[j], qa[2 * s2][0], qa[2 * s2][1], qa[2 * s2][2], qa[2 * s2][3], b0, b1);
                mma_bf16(sc[j], qa[2 * s2 + 1][0], qa[2 * s2 + 1][1], qa[2 * s2 + 1][2], qa[2 * s2 + 1][3], b2, b3);
            }
        }
        // keys past the sequence -> -inf; row maxima (rows g: c0,c1 ; g+8: c2,c3)
        float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int key = kc + 8 * j + 2 * t;
            if (key >= nk) { sc[j][0] = -INFINITY; sc[j][2] = -INFINITY; }
            if (key + 1 >= nk) { sc[j][1] = -INFINITY; sc[j][3] = -INFINITY; }
            cm0 = fmaxf(cm0, fmaxf(sc[j][0], sc[j][1]));
            cm1 = fmaxf(cm1, fmaxf(sc[j][2], sc[j][3]));
        }
        cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1)); cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
        cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1)); cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
        const float mn0 = fmaxf(m0, cm0), mn1 = fmaxf(m1, cm1);       // finite: every chunk holds >= 1 valid key
        const float corr0 = __expf(m0 - mn0), corr1 = __expf(m1 - mn1);
        m0 = mn0; m1 = mn1;
        float rs0 = 0.f, rs1 = 0.f;
        uint32_t pa[4][4];                                           // P as A fragments: k-step = 16 keys = n-tiles 2s, 2s+1
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float p0 = __expf(sc[j][0] - mn0), p1 = __expf(sc[j][1] - mn0);
            const float p2 = __expf(sc[j][2] - mn1), p3 = __expf(sc[j][3] - mn1);
            rs0 += p0 + p1; rs1 += p2 + p3;
            pa[j >> 1][(j & 1) * 2 + 0] = pack_bf16x2(p0, p1);       // a0 (j even) / a2 (j odd): row g
            pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(p2, p3);       // a1 / a3: row g+8
        }
        l0 = l0 * corr0 + rs0; l1 = l1 * corr1 + rs1;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) { o[nt][0] *= corr0; o[nt][1] *= corr0; o[nt][2] *= corr1; o[nt][3] *= corr1; }
        // ---- O += P.V : 8 n-tiles of 8 dims x 4 k-steps of 16 keys
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int vr = 16 * s + (lane & 7) + 8 * ((lane >> 3) & 1);
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t b0, b1, b2, b3;
                ldsm_x4_t(vs + vr * ROWB + (2 * np + (lane >> 4)) * 16, b0, b1, b2, b3);
                mma_bf16(o[2 * np], pa[s][0], pa[s][1], pa[s][2], pa[s][3], b0, b1);
                mma_bf16(o[2 * np + 1], pa[s][0], pa[s][1], pa[s][2], pa[s][3], b2, b3);
            }
        }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    bf16* __restrict__ O = reinterpret_cast<bf16*>(a.o);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const int col = h * HD + 8 * nt + 2 * t;
        if (r0 < nq) *reinterpret_cast<uint32_t*>(O + (size_t)(q0 + r0) * a.ldo + col) = pack_bf16x2(o[nt][0] * i0, o[nt][1] * i0);
        if (r1 < nq) *reinterpret_cast<uint32_t*>(O + (size_t)(q0 + r1) * a.ldo + col) = pack_bf16x2(o[nt][2] * i1, o[nt][3] * i1);
    }
}

}  // namespace

bool attn_enc_mma_supported(const AttnVarlenArgs& a) {
    return a.dt == DT_BF16 && !a.causal && !a.q_mask && !a.k_mask && a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 && a.ldo % 2 == 0;
}

cudaError_t launch_attn_enc_mma(const AttnVarlenArgs& a, cudaStream_t st) {
    if (a.batch <= 0 || a.max_q <= 0) return cudaSuccess;
    dim3 grid((a.max_q + QT - 1) / QT, 8, a.batch);
    attn_enc_mma_kernel<<<grid, 128, 0, st>>>(a);
    return cudaGetLastError();
}
