// Decode-step attention for the bf16 tier: a persistent, TMA-fed streaming kernel (the HBM-bound kernel of the generate loop).
//
// Work unit = (sequence b, head pair hp): 2 heads x nk keys x (K 128 B + V 128 B).  Units are dealt round-robin to
// persistent CTAs.  Each CTA has one producer thread that issues 2-D TMA loads (box = 16 keys x 64 columns, 128-byte
// swizzle; K and V tiles of both heads) into an 8-stage shared-memory ring, running ahead across unit boundaries so the
// memory pipe never drains, and two consumer warps (one per head) that run the flash-decoding inner loop on the legacy
// tensor-core path: S = q.K^T and O += P.V as mma.sync.m16n8k16 (bf16 in, fp32 accumulate) with the single query in row 0
// of the 16-row A operand, operands fetched with ldmatrix from the swizzled tiles, online softmax in fp32 on 4 lanes.
// (A SIMT inner loop needs ~7x the instructions per byte and is issue/latency-bound at ~70 % of HBM speed; tcgen05 is
// pointless for a 1-row problem.)  Self-attention also folds in this step's own key/value (from the QKV GEMM output)
// and appends it to the cache.
//
// Reference semantics: model/attention.py:148-173 with q length 1 (energy * 0.125, softmax, . v); no masks are active in
// generate (model/decoder.py:95: mask all True; causal over a prefix == attend to everything cached).
#include <algorithm>

#include "common.cuh"
#include "tc_gemm.h"

int g_attn_full_tail = 0;
int g_attn_abs_minb = 3;     // attn_abs_kernel: 3 = 128 registers, no spills (default: 1-3 % better with batches in flight); 4 = 96 registers (small spills), 4 CTAs per SM

namespace {

constexpr int CH = 16;            // keys per stage
constexpr int NS = 4;             // ring stages (32 KB ring; an 8-stage ring was measured slower: 3 CTAs per SM instead of 4)
constexpr int TILE = CH * 256;    // bytes of the K (or V) tiles of one stage: CH keys x 2 heads x 64 x bf16
constexpr int STAGE = 2 * TILE;
constexpr float SCALE = 0.125f;

struct Args {
    const bf16* q; int ldq;
    const bf16* knew; const bf16* vnew; int ldnew;     // self only
    bf16* cache;                                        // self: element (row 0, col 0) of the matrix the tensor map covers (append target)
    const int* k_off;                                   // cross: token offsets [B+1]
    const int* step;                                    // self: keys already cached
    bf16* o; int ldo;
    int batch;
    int ld, col0, col_h, v_col, row_h, row_b;           // KvLayout (kernels.h)
    int full_tail;                                      // debug: fetch the last, partial stage with full 16-row boxes
    unsigned long long* trace; const int* trace_step; int trace_k;     // debug timeline: [3][256 steps][8 launches] of this branch
};
TX_DEVINL unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

TX_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
TX_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
TX_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
TX_DEVINL void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
TX_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// K/V is read exactly once per step: the loads carry an L2 evict-first policy so that the ~1 GB/step stream does not flush
// the decoder weights and the activations the concurrently running GEMM / LayerNorm kernels of other branches live on.
TX_DEVINL uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
TX_DEVINL void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol) : "memory");
}
TX_DEVINL void tma_load_2d_nohint(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
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
// m16n8k16 with only rows 0..7 of A in use (a1 = a3 = 0): rows 8..15 of the accumulator stay zero, so they are fed as constants
// and their results land in two scratch registers instead of occupying two live registers per n-tile
TX_DEVINL void mma_bf16_top(float& c0, float& c1, uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
    float j0, j1;
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %5}, {%7, %8}, {%0, %1, %9, %9};"
                 : "+f"(c0), "+f"(c1), "=f"(j0), "=f"(j1) : "r"(a0), "r"(0u), "r"(a2), "r"(b0), "r"(b1), "f"(0.f));
}
TX_DEVINL void sts_f2(uint32_t addr, float x, float y) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(x), "f"(y) : "memory"); }
TX_DEVINL float2 lds_f2(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
    return v;
}
TX_DEVINL void sts_f1(uint32_t addr, float x) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(x) : "memory"); }
TX_DEVINL void sts_h1(uint32_t addr, float x) {
    const unsigned short h = __bfloat16_as_ushort(__float2bfloat16_rn(x));
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(h) : "memory");
}
TX_DEVINL uint32_t lds_u1(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
TX_DEVINL void sts_u4(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
TX_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
TX_DEVINL float2 unpack_bf16x2(uint32_t w) { return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w)); }

// Stage layout: [K head0 | K head1 | V head0 | V head1], each a 16 x 64 bf16 tile (2 KB) in the TMA 128B-swizzle layout:
// element (row r, 16-byte chunk c) lives at r*128 + ((c ^ (r & 7)) << 4).
constexpr int HTILE = CH * 128;

template <bool SELF>
__global__ void __launch_bounds__(96) attn_decode_tma_kernel(const __grid_constant__ CUtensorMap tm,
                                                             const __grid_constant__ CUtensorMap tm4, const Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* ring = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + NS * STAGE);
    uint64_t* empty = full + NS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    pdl_launch_dependents();
    const unsigned long long t_entry = a.trace ? gtime() : 0ull;
    // Rows past the end of a sequence (fetched by the last box or stale in a partially filled stage) may hold anything:
    // their scores are replaced by -inf and their V fragments are cleared below, so the ring needs no initialisation.
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm4) : "memory");
        for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 2); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // Cross-attention: the producer only reads the encoder-memory K/V and the token offsets, both written before the
    // generate loop started, so its TMA stream starts while the predecessor (the query GEMM) is still running.
    if (SELF || warp != 0) pdl_wait();

    const int units = a.batch * 4;
    const int t = SELF ? ldcg_i32(a.step) : 0;
    unsigned long long* tr = nullptr;
    if (a.trace && threadIdx.x == 32) {
        const int ts = min(ldcg_i32(a.trace_step), 255);
        tr = a.trace + (size_t)ts * 8 + a.trace_k;
        atomicMin(tr, t_entry);
        atomicMin(tr + 2048, gtime());
    }

    if (warp == 0) {
        // ------------------------------------------------------------ producer
        if (lane == 0) {
            int it = 0;
            const uint64_t pol = l2_evict_first_policy();
            asm volatile("fence.proxy.async.global;" ::: "memory");     // K/V rows were appended by generic-proxy stores of earlier kernels
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int b = u >> 2, hp = u & 3;
                int row0, nk;
                if (SELF) { row0 = b * a.row_b; nk = t; }
                else { row0 = ldcg_i32(a.k_off + b); nk = ldcg_i32(a.k_off + b + 1) - row0; }
                const int nchunk = (nk + CH - 1) / CH;
                const int h0 = 2 * hp, h1 = h0 + 1;
                const int kc0 = a.col0 + h0 * a.col_h, kc1 = a.col0 + h1 * a.col_h;
                const int r0 = row0 + h0 * a.row_h, r1 = row0 + h1 * a.row_h;
                for (int c = 0; c < nchunk; ++c, ++it) {
                    const int s = it % NS, ph = (it / NS) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* st = ring + s * STAGE;
                    const int left = nk - c * CH;
                    const int rc = c * CH;
                    if (left >= CH || (a.full_tail & 1)) {
                        mbar_expect_tx(&full[s], STAGE);
                        tma_load_2d(&tm, &full[s], st, kc0, r0 + rc, pol);
                        tma_load_2d(&tm, &full[s], st + HTILE, kc1, r1 + rc, pol);
                        tma_load_2d(&tm, &full[s], st + 2 * HTILE, kc0 + a.v_col, r0 + rc, pol);
                        tma_load_2d(&tm, &full[s], st + 3 * HTILE, kc1 + a.v_col, r1 + rc, pol);
                    } else {               // tail: 4-row boxes, at most 3 rows fetched beyond the sequence
                        const int n4 = (left + 3) >> 2;
                        mbar_expect_tx(&full[s], n4 * 4 * 512);
                        for (int j = 0; j < n4; ++j) {
                            tma_load_2d(&tm4, &full[s], st + j * 512, kc0, r0 + rc + 4 * j, pol);
                            tma_load_2d(&tm4, &full[s], st + HTILE + j * 512, kc1, r1 + rc + 4 * j, pol);
                            tma_load_2d(&tm4, &full[s], st + 2 * HTILE + j * 512, kc0 + a.v_col, r0 + rc + 4 * j, pol);
                            tma_load_2d(&tm4, &full[s], st + 3 * HTILE + j * 512, kc1 + a.v_col, r1 + rc + 4 * j, pol);
                        }
                    }
                }
            }
        }
        return;
    }
    // ---------------------------------------------------------------- consumers: warp 1 -> head 2*hp, warp 2 -> head 2*hp+1
    const int hd = warp - 1;
    const int g = lane >> 2, tq = lane & 3;             // mma fragment coordinates: row group / thread-in-group
    const bool row0_lane = g == 0;                       // the single query lives in row 0 of the 16-row A operand
    // ldmatrix address pieces (thread i supplies row i&7 of matrix i>>3)
    const int lm_r = lane & 7, lm_m = lane >> 3;
    int it = 0;
    // header: q, this step's k and v for (b, head); lanes 0..3 hold the words of their fragment positions
    uint32_t q_w[8], kn_w[8], vn_w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { q_w[i] = 0; kn_w[i] = 0; vn_w[i] = 0; }
    auto load_header = [&](int u) {
        const int b = u >> 2, h = (u & 3) * 2 + hd;
        if (row0_lane) {
            const uint32_t* qp = reinterpret_cast<const uint32_t*>(a.q + (size_t)b * a.ldq + h * 64);
#pragma unroll
            for (int s = 0; s < 4; ++s) { q_w[2 * s] = ldcg_u32(qp + 8 * s + tq); q_w[2 * s + 1] = ldcg_u32(qp + 8 * s + 4 + tq); }   // dims 16s+2t, 16s+8+2t
            if (SELF) {
                const uint32_t* kp = reinterpret_cast<const uint32_t*>(a.knew + (size_t)b * a.ldnew + h * 64);
                const uint32_t* vp = reinterpret_cast<const uint32_t*>(a.vnew + (size_t)b * a.ldnew + h * 64);
#pragma unroll
                for (int s = 0; s < 4; ++s) { kn_w[2 * s] = ldcg_u32(kp + 8 * s + tq); kn_w[2 * s + 1] = ldcg_u32(kp + 8 * s + 4 + tq); }
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) vn_w[nt] = ldcg_u32(vp + 4 * nt + tq);                            // dims 8nt+2t, +1
            }
        }
    };
    if ((int)blockIdx.x < units) load_header(blockIdx.x);
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int b = u >> 2, h = (u & 3) * 2 + hd;
        // A fragments of q (scaled by 0.125, exact in bf16): a0 = (row g, k 2t..), a2 = (row g, k 2t+8..); rows 8..15 zero
        uint32_t qa[8];
        float kn_f[16], vn_f[16], q_f[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 f = unpack_bf16x2(q_w[i]);
            q_f[2 * i] = f.x * SCALE; q_f[2 * i + 1] = f.y * SCALE;
            qa[i] = row0_lane ? pack_bf16x2(q_f[2 * i], q_f[2 * i + 1]) : 0u;
            if (SELF) {
                const float2 kf = unpack_bf16x2(kn_w[i]), vf = unpack_bf16x2(vn_w[i]);
                kn_f[2 * i] = kf.x; kn_f[2 * i + 1] = kf.y; vn_f[2 * i] = vf.x; vn_f[2 * i + 1] = vf.y;
            }
        }
        if (SELF && lane < 8) {      // append this step's k / v row to the cache (16 B per lane), K at h*64, V at 512 + h*64
            const uint4 kr = ldcg_u4(a.knew + (size_t)b * a.ldnew + h * 64 + lane * 8);
            const uint4 vr = ldcg_u4(a.vnew + (size_t)b * a.ldnew + h * 64 + lane * 8);
            bf16* row = a.cache + ((size_t)b * a.row_b + (size_t)h * a.row_h + t) * a.ld + a.col0 + h * a.col_h + lane * 8;
            *reinterpret_cast<uint4*>(row) = kr;
            *reinterpret_cast<uint4*>(row + a.v_col) = vr;
            // the next step reads this row through the async proxy (TMA): order the generic-proxy stores against it
            asm volatile("fence.proxy.async.global;" ::: "memory");
        }
        const int un = u + gridDim.x;
        if (un < units) load_header(un);          // prefetch the next unit's header while this one streams
        int nk;
        if (SELF) nk = t; else nk = ldcg_i32(a.k_off + b + 1) - ldcg_i32(a.k_off + b);
        const int nchunk = (nk + CH - 1) / CH;
        float m = -INFINITY, l = 0.f;
        float o[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
        for (int c = 0; c < nchunk; ++c, ++it) {
            const int s = it % NS, ph = (it / NS) & 1;
            mbar_wait(&full[s], ph);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const uint32_t kt = smem_u32(ring + s * STAGE + hd * HTILE);
            const uint32_t vt = kt + 2 * HTILE;
            // ---- S = q.K^T : 2 n-tiles of 8 keys, 4 k-steps of 16 dims
            float sc[2][4];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
                const int r = 8 * j + lm_r;
#pragma unroll
                for (int s2 = 0; s2 < 2; ++s2) {
                    uint32_t b0, b1, b2, b3;
                    ldsm_x4(kt + r * 128 + (((4 * s2 + lm_m) ^ (r & 7)) << 4), b0, b1, b2, b3);
                    mma_bf16(sc[j], qa[4 * s2], 0u, qa[4 * s2 + 1], 0u, b0, b1);
                    mma_bf16(sc[j], qa[4 * s2 + 2], 0u, qa[4 * s2 + 3], 0u, b2, b3);
                }
            }
            // row 0 scores: lane tq holds keys 8j+2tq, 8j+2tq+1
            const int kbase = c * CH + 2 * tq;
            float p[2][2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                p[j][0] = (kbase + 8 * j < nk) ? sc[j][0] : -INFINITY;
                p[j][1] = (kbase + 8 * j + 1 < nk) ? sc[j][1] : -INFINITY;
            }
            float cm = fmaxf(fmaxf(p[0][0], p[0][1]), fmaxf(p[1][0], p[1][1]));
            cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 1));
            cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 2));
            const float mn = fmaxf(m, cm);                 // finite on row 0: every stage holds at least one valid key
            const float corr = __expf(m - mn);
#pragma unroll
            for (int j = 0; j < 2; ++j) { p[j][0] = __expf(p[j][0] - mn); p[j][1] = __expf(p[j][1] - mn); }
            l = l * corr + (p[0][0] + p[0][1]) + (p[1][0] + p[1][1]);
            m = mn;
            const uint32_t pa0 = row0_lane ? pack_bf16x2(p[0][0], p[0][1]) : 0u;     // keys 2t, 2t+1
            const uint32_t pa2 = row0_lane ? pack_bf16x2(p[1][0], p[1][1]) : 0u;     // keys 8+2t, 9+2t
            // ---- O = O*corr + P.V : 8 n-tiles of 8 dims, one k-step of 16 keys
            const int vr = (lane & 7) + 8 * ((lane >> 3) & 1);
            // V rows past the end of the sequence were fetched (<= 3 of them) or are stale: they carry p = 0, but 0 * NaN = NaN,
            // so their halves of the B fragments (keys 2t, 2t+1 | 2t+8, 2t+9) are cleared in the last, partial stage.
            uint32_t vm_lo = 0xffffffffu, vm_hi = 0xffffffffu;
            if (nk - c * CH < CH) {
                const int k0 = c * CH + 2 * tq;
                vm_lo = (k0 < nk ? 0x0000ffffu : 0u) | (k0 + 1 < nk ? 0xffff0000u : 0u);
                vm_hi = (k0 + 8 < nk ? 0x0000ffffu : 0u) | (k0 + 9 < nk ? 0xffff0000u : 0u);
            }
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t b0, b1, b2, b3;
                ldsm_x4_t(vt + vr * 128 + (((2 * np + (lane >> 4)) ^ (vr & 7)) << 4), b0, b1, b2, b3);
                b0 &= vm_lo; b2 &= vm_lo; b1 &= vm_hi; b3 &= vm_hi;
                o[2 * np][0] *= corr; o[2 * np][1] *= corr;
                o[2 * np + 1][0] *= corr; o[2 * np + 1][1] *= corr;
                mma_bf16(o[2 * np], pa0, 0u, pa2, 0u, b0, b1);
                mma_bf16(o[2 * np + 1], pa0, 0u, pa2, 0u, b2, b3);
            }
            // The stage is about to be handed back to the TMA producer: order this warp's generic-proxy reads (ldmatrix) before
            // the async-proxy overwrite.  Without the cross-proxy fence the refill of a reused stage occasionally overtook the
            // last V reads (1-3 % error in one head of one sequence, a few times per thousand launches under concurrency).
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        if (SELF) {
            // this step's own key / value: dot over the lane's 16 dims, folded over the 4 lanes of row 0
            float d = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) d = fmaf(q_f[i], kn_f[i], d);
            d += __shfl_xor_sync(0xffffffffu, d, 1);
            d += __shfl_xor_sync(0xffffffffu, d, 2);
            const float mn = fmaxf(m, d);
            const float corr = __expf(m - mn), pn = __expf(d - mn);
            l = l * corr;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                o[nt][0] = fmaf(pn, vn_f[2 * nt], o[nt][0] * corr);
                o[nt][1] = fmaf(pn, vn_f[2 * nt + 1], o[nt][1] * corr);
            }
            l += (tq == 0) ? pn : 0.f;         // l is a per-lane partial, summed over the 4 lanes below
        }
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        if (row0_lane) {
            const float inv = 1.0f / l;
            uint32_t* op = reinterpret_cast<uint32_t*>(a.o + (size_t)b * a.ldo + h * 64);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) op[4 * nt + tq] = pack_bf16x2(o[nt][0] * inv, o[nt][1] * inv);
        }
    }
    if (tr) atomicMax(tr + 4096, gtime());
}


// ------------------------------------------------------------------------------------------------ absorbed ("latent") attention
// Decode-step attention of the bf16 generate loop with the key / value projections absorbed into the query and output
// projections.  For a head h with key / value weights Wk_h, Wv_h [64 x 256] and latent rows z_j [256] (cross-attention:
// z_j = encoder memory token j; self-attention: z_j = the layer's LayerNorm'd input at position j, i.e. what the reference
// feeds to to_k / to_v, model/attention.py:114-126):
//   q_h . K_h[j] = q_h . (Wk_h z_j) = (q_h Wk_h) . z_j                   Q'_h = q_h Wk_h       (256 wide; folded query GEMM)
//   sum_j p_h[j] V_h[j] = (sum_j p_h[j] z_j) Wv_h^T                      C_h = P_h . Z         (256 wide; folded out-projection)
// so all 8 heads of a sequence stream the SAME [n, 256] bf16 latent rows instead of their own [n, 64 | 64] K/V slices:
// 512 bytes per key and layer instead of 2,048 -- 4x less HBM traffic for the loop's dominant stream, and the self-attention
// cache holds 256 instead of 1,024 values per position (model/attention.py:148-173 computes the left-hand sides).
// Work unit = sequence.  The 8 heads are rows 0..7 of the 16-row mma.sync A operand.  Stage = 16 latent rows x 256 columns as
// four 64-column TMA boxes (128B swizzle), 8 KB.  Four consumer warps, warp w owns column block w: it computes the partial
// scores Q'[:, 64w..64w+63] . Z[:, 64w..]^T (8 MMAs, k = 64), the partials are summed through shared memory (one named barrier per
// stage, double-buffered), every warp runs the identical online softmax, and accumulates C[:, 64w..64w+63] += P . Z[:, 64w..] (8 MMAs).
// Self-attention: this step's own latent row (position t) is written into the last stage by the consumers (it is key t of
// that stage) and appended to the cache for the following steps.
constexpr int AW = 4;                       // consumer warps = 64-column blocks of a latent row
constexpr int ANS = 5;                      // ring stages of the absorbed kernel (40 KB)
constexpr int XROW = 40;                    // floats per (warp, head) row of the score exchange: 32 keys, padded so that a half-warp's float2 accesses hit 32 distinct banks
struct AbsArgs {
    const bf16* q; int ldq;            // [batch, ldq]: head h at h*256 (absorbed query, unscaled)
    const int* k_off;                  // cross: token offsets [batch + 1] into the latent matrix the tensor maps cover
    int uni_nk;                        // cross: > 0 = every sequence has this many memory tokens and sequence b starts at row b * uni_nk (k_off is not read)
    const bf16* znew; int ldz;         // self: this step's latent rows [batch, ldz]
    bf16* cache; int tcap;             // self: latent cache [batch][tcap][256] (row b*tcap + j); the tensor maps cover it
    const int* step;                   // self: positions already cached (= index of this step's row)
    bf16* o; int ldo;                  // [batch, ldo]: head h at h*256
    int batch;
    unsigned long long* trace; const int* trace_step; int trace_k;
    unsigned long long* dbg;           // debug: sums over CTAs of [wait for predecessor, first data, stage loop, epilogue] ns + count
};

template <bool SELF, int MINB>
__global__ void __launch_bounds__(32 * (AW + 1), MINB) attn_abs_kernel(const __grid_constant__ CUtensorMap tm,
                                                                 const __grid_constant__ CUtensorMap tm4, const AbsArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* ring = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    float* xbuf = reinterpret_cast<float*>(ring + ANS * STAGE);           // [2][AW][8 heads][XROW] partial scores of two stages (32 keys)
    uint64_t* full = reinterpret_cast<uint64_t*>(xbuf + 2 * AW * 8 * XROW);
    uint64_t* empty = full + ANS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    pdl_launch_dependents();
    const unsigned long long t_entry = a.trace ? gtime() : 0ull;
    const unsigned long long t_entry0 = a.dbg ? gtime() : 0ull;
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm4) : "memory");
        for (int s = 0; s < ANS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], AW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // cross: the producer only reads the encoder memory and the token offsets, both written before the generate loop started
    if (SELF || warp != 0) pdl_wait();
    const int units = a.batch;
    const int t = SELF ? ldcg_i32(a.step) : 0;
    unsigned long long* tr = nullptr;
    if (a.trace && threadIdx.x == 32) {
        const int ts = min(ldcg_i32(a.trace_step), 255);
        tr = a.trace + (size_t)ts * 8 + a.trace_k;
        atomicMin(tr, t_entry);
        atomicMin(tr + 2048, gtime());
    }

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            if (SELF) asm volatile("fence.proxy.async.global;" ::: "memory");     // cache rows were appended by generic-proxy stores of earlier steps
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                int row0, nc;      // nc = rows to fetch; self: the t cached rows (this step's own row is added by the consumers)
                if (SELF) { row0 = u * a.tcap; nc = t; }
                else if (a.uni_nk > 0) { row0 = u * a.uni_nk; nc = a.uni_nk; }      // equal memory lengths: no offset loads in front of the first TMA
                else { row0 = ldcg_i32(a.k_off + u); nc = ldcg_i32(a.k_off + u + 1) - row0; }
                const int nchunk = ((SELF ? nc + 1 : nc) + CH - 1) / CH;
                for (int c = 0; c < nchunk; ++c, ++it) {
                    const int s = it % ANS, ph = (it / ANS) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* st = ring + s * STAGE;
                    const int left = nc - c * CH, r = row0 + c * CH;
                    if (left >= CH) {
                        mbar_expect_tx(&full[s], STAGE);
#pragma unroll
                        for (int cb = 0; cb < 4; ++cb) tma_load_2d_nohint(&tm, &full[s], st + cb * HTILE, 64 * cb, r);
                    } else {               // tail: 4-row boxes (none at all when only this step's own row is left)
                        const int n4 = left > 0 ? (left + 3) >> 2 : 0;
                        mbar_expect_tx(&full[s], n4 * 4 * 512);
                        for (int j = 0; j < n4; ++j)
#pragma unroll
                            for (int cb = 0; cb < 4; ++cb) tma_load_2d_nohint(&tm4, &full[s], st + cb * HTILE + j * 512, 64 * cb, r + 4 * j);
                    }
                }
            }
        }
        return;
    }
    // ---------------------------------------------------------------- consumers: warp w+1 owns column block w
    const int cw = warp - 1;
    const int g = lane >> 2, tq = lane & 3;               // fragment row (= head) / thread-in-group
    const int lm_r = lane & 7, lm_m = lane >> 3;
    const uint32_t ring_u32 = smem_u32(ring), xbuf_u32 = smem_u32(xbuf);
    // Both contractions put the KEYS / latent columns on the 16-row M side of mma.sync.m16n8k16 and the 8 heads on its 8-wide N
    // side, so no MMA row is padding:  S^T[16 keys x 8 heads] = Z[16 keys x 16 cols] . Q'^T[16 cols x 8 heads]  (A = Z by ldmatrix),
    //                                  C^T[16 cols x 8 heads] += Z^T[16 cols x 16 keys] . P^T[16 keys x 8 heads] (A = Z by ldmatrix.trans).
    // Swizzled ldmatrix offsets inside a 16 x 64 tile, the same for every stage.  Matrix m of an x4 load is addressed by lanes 8m..8m+7:
    //   scores, k-step ks: m0 = keys 0-7 / chunk 2ks, m1 = keys 8-15 / chunk 2ks, m2 = keys 0-7 / chunk 2ks+1, m3 = keys 8-15 / chunk 2ks+1  (= a0..a3)
    //   P.Z, m-tile mt   : the same four 8 x 8 blocks, transposed on load: (m0, m2, m1, m3) = (a0, a1, a2, a3)
    uint32_t off_z[4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) { const int r = (lm_m & 1) * 8 + lm_r; off_z[ks] = r * 128 + (((2 * ks + (lm_m >> 1)) ^ (r & 7)) << 4); }
    int it = 0;
    uint32_t qn[8];
    auto load_header = [&](int u) {       // words of row g, columns 64cw..: k-step s holds dims 16s+2t,+1 and 16s+8+2t,+1
        const uint32_t* qp = reinterpret_cast<const uint32_t*>(a.q + (size_t)u * a.ldq + g * 256 + cw * 64);
#pragma unroll
        for (int s = 0; s < 4; ++s) { qn[2 * s] = ldcg_u32(qp + 8 * s + tq); qn[2 * s + 1] = ldcg_u32(qp + 8 * s + 4 + tq); }
    };
    if ((int)blockIdx.x < units) load_header(blockIdx.x);
    int xiter = 0;          // score-exchange buffer parity: runs across units (a CTA may own several), never reset
    const unsigned long long t_ready = a.dbg ? gtime() : 0ull;
    unsigned long long t_first = 0ull;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
        uint32_t qa[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 f = unpack_bf16x2(qn[i]);
            qa[i] = pack_bf16x2(f.x * SCALE, f.y * SCALE);          // exact: a power of two
        }
        uint4 zrow = make_uint4(0u, 0u, 0u, 0u);
        if (SELF && lane < 8) {      // this step's latent row, columns 64cw + 8*lane .. +7: append to the cache, keep for the last stage
            zrow = ldcg_u4(a.znew + (size_t)u * a.ldz + cw * 64 + lane * 8);
            *reinterpret_cast<uint4*>(a.cache + ((size_t)u * a.tcap + t) * 256 + cw * 64 + lane * 8) = zrow;
            asm volatile("fence.proxy.async.global;" ::: "memory");      // later steps read the row through the async proxy (TMA)
        }
        const int un = u + gridDim.x;
        if (un < units) load_header(un);
        int nk;
        if (SELF) nk = t + 1;
        else if (a.uni_nk > 0) nk = a.uni_nk;
        else nk = ldcg_i32(a.k_off + u + 1) - ldcg_i32(a.k_off + u);
        const int nchunk = (nk + CH - 1) / CH;
        float m = -INFINITY, l = 0.f;
        float o[4][4];       // m-tile mt (columns 16mt + g, 16mt + g + 8 of the warp's block) x heads (2tq, 2tq + 1)
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) { o[mt][0] = o[mt][1] = o[mt][2] = o[mt][3] = 0.f; }
        // Two stages (32 keys) per iteration: one score exchange / barrier / softmax update for both, two independent MMA chains.
        for (int c = 0; c < nchunk; c += 2) {
            const bool two = c + 1 < nchunk;
            const bool last = c + 2 >= nchunk;
            uint32_t kts[2];
            int sidx[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (q == 0 || two) {
                    const int iq = it + q;
                    sidx[q] = iq % ANS;
                    mbar_wait(&full[sidx[q]], (iq / ANS) & 1);
                    kts[q] = ring_u32 + sidx[q] * STAGE + cw * HTILE;
                } else { sidx[q] = 0; kts[q] = 0; }
            }
            // (no proxy fence here: observing the mbarrier phase makes the TMA writes of the stage visible to this thread's ldmatrix)
            if (a.dbg && t_first == 0ull) t_first = gtime();
            if (SELF && last) {      // key t of the sequence = this step's own row, row t % 16 of the last stage
                const int r = t & (CH - 1);
                if (lane < 8) sts_u4(kts[two ? 1 : 0] + r * 128 + ((lane ^ (r & 7)) << 4), zrow);
                __syncwarp();
            }
            // ---- partial S^T = Z[:, block] . Q'[:, block]^T : per stage one 16-key m-tile, 4 k-steps of 16 columns, N = the 8 heads
            float sc[2][4];      // (key g, head 2tq), (key g, head 2tq+1), (key g+8, head 2tq), (key g+8, head 2tq+1)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                sc[q][0] = sc[q][1] = sc[q][2] = sc[q][3] = 0.f;
                if (q == 0 || two) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        uint32_t a0, a1, a2, a3;
                        ldsm_x4(kts[q] + off_z[ks], a0, a1, a2, a3);
                        mma_bf16(sc[q], a0, a1, a2, a3, qa[2 * ks], qa[2 * ks + 1]);
                    }
                }
            }
            // exchange buffer: [warp][head][key of the iteration]; this thread READS row g (= head g), keys 16q+8j+2tq, +1 ...
            const uint32_t xbase = xbuf_u32 + (xiter & 1) * (AW * 8 * XROW * 4);
            const uint32_t xb = xbase + (g * XROW + 2 * tq) * 4;
            ++xiter;
            // ... and WRITES its partials for heads 2tq, 2tq+1 and keys g, g+8 of both stages
            const uint32_t xw = xbase + cw * (8 * XROW * 4) + (2 * tq * XROW + g) * 4;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                sts_f1(xw + (16 * q) * 4, sc[q][0]);
                sts_f1(xw + (XROW + 16 * q) * 4, sc[q][1]);
                sts_f1(xw + (16 * q + 8) * 4, sc[q][2]);
                sts_f1(xw + (XROW + 16 * q + 8) * 4, sc[q][3]);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * AW) : "memory");
            float p[2][2][2];
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    float2 acc = lds_f2(xb + (16 * q + 8 * j) * 4);
#pragma unroll
                    for (int w2 = 1; w2 < AW; ++w2) {
                        const float2 v = lds_f2(xb + w2 * (8 * XROW * 4) + (16 * q + 8 * j) * 4);
                        acc.x += v.x; acc.y += v.y;
                    }
                    p[q][j][0] = acc.x; p[q][j][1] = acc.y;
                }
            uint32_t vm_lo[2] = {0xffffffffu, 0xffffffffu}, vm_hi[2] = {0xffffffffu, 0xffffffffu};
            if (last) {      // keys past the end of the sequence: score -> -inf; their Z rows carry p = 0, but 0 * NaN = NaN -> cleared below
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int k0 = (c + q) * CH + 2 * tq;
                    if (k0 >= nk) p[q][0][0] = -INFINITY;
                    if (k0 + 1 >= nk) p[q][0][1] = -INFINITY;
                    if (k0 + 8 >= nk) p[q][1][0] = -INFINITY;
                    if (k0 + 9 >= nk) p[q][1][1] = -INFINITY;
                    vm_lo[q] = (k0 < nk ? 0x0000ffffu : 0u) | (k0 + 1 < nk ? 0xffff0000u : 0u);
                    vm_hi[q] = (k0 + 8 < nk ? 0x0000ffffu : 0u) | (k0 + 9 < nk ? 0xffff0000u : 0u);
                }
            }
            float cm = fmaxf(fmaxf(fmaxf(p[0][0][0], p[0][0][1]), fmaxf(p[0][1][0], p[0][1][1])),
                             fmaxf(fmaxf(p[1][0][0], p[1][0][1]), fmaxf(p[1][1][0], p[1][1][1])));
            cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 1));
            cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 2));
            const float mn = fmaxf(m, cm);                 // finite: the first stage of an iteration holds at least one valid key
            const float corr = __expf(m - mn);
            float ps = 0.f;
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    p[q][j][0] = __expf(p[q][j][0] - mn); p[q][j][1] = __expf(p[q][j][1] - mn);
                    ps += p[q][j][0] + p[q][j][1];
                }
            l = l * corr + ps;
            m = mn;
            {      // the accumulators belong to heads 2tq, 2tq+1; their softmax state lives in the lanes of rows g = 2tq, 2tq+1
                const float c0 = __shfl_sync(0xffffffffu, corr, 8 * tq), c1 = __shfl_sync(0xffffffffu, corr, 8 * tq + 4);
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) { o[mt][0] *= c0; o[mt][1] *= c1; o[mt][2] *= c0; o[mt][3] *= c1; }
            }
            // ---- C^T[block, :] += Z[:, block]^T . P^T : per stage 4 m-tiles of 16 columns, one k-step of 16 keys, N = the 8 heads
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (q == 0 || two) {
                    const uint32_t pb0 = pack_bf16x2(p[q][0][0], p[q][0][1]);     // (keys 2t, 2t+1; head g)
                    const uint32_t pb1 = pack_bf16x2(p[q][1][0], p[q][1][1]);     // (keys 8+2t, 9+2t; head g)
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) {
                        uint32_t z0, z1, z2, z3;        // transposed blocks: keys 0-7 / chunk 2mt, keys 8-15 / 2mt, keys 0-7 / 2mt+1, keys 8-15 / 2mt+1
                        ldsm_x4_t(kts[q] + off_z[mt], z0, z1, z2, z3);
                        if (last) { z0 &= vm_lo[q]; z2 &= vm_lo[q]; z1 &= vm_hi[q]; z3 &= vm_hi[q]; }
                        mma_bf16(o[mt], z0, z2, z1, z3, pb0, pb1);
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) { mbar_arrive(&empty[sidx[0]]); if (two) mbar_arrive(&empty[sidx[1]]); }
            it += two ? 2 : 1;
        }
        const unsigned long long t_loop = a.dbg ? gtime() : 0ull;
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        const float inv = 1.0f / l;
        {   // normalise, transpose through this warp's own slice of the exchange buffer the NEXT iteration would use (every warp is past
            // the barrier that followed its last reads of it), then store 128-byte row segments: out[u][head * 256 + 64cw ..]
            const float i0 = __shfl_sync(0xffffffffu, inv, 8 * tq), i1 = __shfl_sync(0xffffffffu, inv, 8 * tq + 4);
            const uint32_t ow = xbuf_u32 + (xiter & 1) * (AW * 8 * XROW * 4) + cw * (8 * XROW * 4);      // [8 heads][64 cols] bf16, head rows 136 B apart (conflict-free 16-bit stores)
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
                sts_h1(ow + (2 * tq) * 136 + (16 * mt + g) * 2, o[mt][0] * i0);
                sts_h1(ow + (2 * tq + 1) * 136 + (16 * mt + g) * 2, o[mt][1] * i1);
                sts_h1(ow + (2 * tq) * 136 + (16 * mt + g + 8) * 2, o[mt][2] * i0);
                sts_h1(ow + (2 * tq + 1) * 136 + (16 * mt + g + 8) * 2, o[mt][3] * i1);
            }
            __syncwarp();
            uint32_t* op = reinterpret_cast<uint32_t*>(a.o + (size_t)u * a.ldo + g * 256 + cw * 64);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) op[4 * nt + tq] = lds_u1(ow + g * 136 + (4 * nt + tq) * 4);
            __syncwarp();        // the next unit's first exchange writes this slice again
        }
        if (a.dbg && threadIdx.x == 32 && u == (int)blockIdx.x) {
            atomicAdd(a.dbg + 0, t_ready - t_entry0); atomicAdd(a.dbg + 1, t_first - t_ready); atomicAdd(a.dbg + 2, t_loop - t_first);
            atomicAdd(a.dbg + 3, gtime() - t_loop); atomicAdd(a.dbg + 4, 1ull);
        }
    }
    if (tr) atomicMax(tr + 4096, gtime());
    if (a.dbg && threadIdx.x == 32) { atomicAdd(a.dbg + 5, gtime() - t_entry0); atomicAdd(a.dbg + 6, 1ull); }      // CTA residency, all units
}

int g_smem_set = 0;

}  // namespace

bool attn_decode_tma_supported(const AttnDecodeArgs& a) {
    if (a.dt != DT_BF16 || a.ldkv % 8 != 0 || a.ldq % 8 != 0 || a.ldo % 8 != 0) return false;
    if (a.knew && a.ldnew % 8 != 0) return false;
    return true;
}

cudaError_t launch_attn_decode_tma(const AttnDecodeArgs& a, const KvLayout& lay, int max_ctas, cudaStream_t st) {
    if (a.batch <= 0) return cudaSuccess;
    CUtensorMap tm, tm4;
    cudaError_t e = tma_map_2d_bf16(lay.map_base, lay.map_rows, lay.map_cols, lay.ld, CH, 64, 1, &tm);
    if (e != cudaSuccess) return e;
    if ((e = tma_map_2d_bf16(lay.map_base, lay.map_rows, lay.map_cols, lay.ld, 4, 64, 1, &tm4)) != cudaSuccess) return e;
    const size_t smem = (size_t)NS * STAGE + 1024 + 2 * NS * 8 + 64;
    if (!g_smem_set) {
        if ((e = cudaFuncSetAttribute(attn_decode_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(attn_decode_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        g_smem_set = 1;
    }
    Args k{};
    k.q = (const bf16*)a.q; k.ldq = a.ldq; k.knew = (const bf16*)a.knew; k.vnew = (const bf16*)a.vnew; k.ldnew = a.ldnew;
    k.cache = (bf16*)const_cast<void*>(lay.map_base); k.k_off = a.k_off; k.step = a.step; k.o = (bf16*)a.o; k.ldo = a.ldo;
    k.full_tail = g_attn_full_tail;
    k.trace = a.trace; k.trace_step = a.trace_step; k.trace_k = a.trace_k;
    k.batch = a.batch; k.ld = lay.ld; k.col0 = lay.col0; k.col_h = lay.col_h; k.v_col = lay.v_col; k.row_h = lay.row_h; k.row_b = lay.row_b;
    const int units = a.batch * 4;
    const int grid = units < max_ctas ? units : max_ctas;
    if (a.knew) return launch_pdl(PDL_ATTN_TMA, attn_decode_tma_kernel<true>, dim3(grid), dim3(96), smem, st, tm, tm4, k);
    return launch_pdl(PDL_ATTN_TMA, attn_decode_tma_kernel<false>, dim3(grid), dim3(96), smem, st, tm, tm4, k);
}


cudaError_t launch_attn_abs(const AttnAbsArgs& a, int max_ctas, cudaStream_t st) {
    if (a.batch <= 0) return cudaSuccess;
    if (a.ldq % 8 != 0 || a.ldo % 8 != 0 || (a.znew && a.ldz % 8 != 0)) return cudaErrorInvalidValue;
    CUtensorMap tm, tm4;
    cudaError_t e = tma_map_2d_bf16(a.latent, a.latent_rows, 256, 256, CH, 64, 1, &tm);
    if (e != cudaSuccess) return e;
    if ((e = tma_map_2d_bf16(a.latent, a.latent_rows, 256, 256, 4, 64, 1, &tm4)) != cudaSuccess) return e;
    const size_t smem = (size_t)ANS * STAGE + 1024 + 2 * AW * 8 * XROW * 4 + 2 * ANS * 8 + 64;
    static int smem_set = 0;
    if (!smem_set) {
        if ((e = cudaFuncSetAttribute(attn_abs_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(attn_abs_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(attn_abs_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(attn_abs_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        smem_set = 1;
    }
    AbsArgs k{};
    k.q = (const bf16*)a.q; k.ldq = a.ldq; k.k_off = a.k_off; k.uni_nk = a.znew ? 0 : a.uni_nk; k.o = (bf16*)a.o; k.ldo = a.ldo; k.batch = a.batch;
    k.znew = (const bf16*)a.znew; k.ldz = a.ldz; k.cache = (bf16*)const_cast<void*>(a.latent); k.tcap = a.tcap; k.step = a.step;
    k.trace = a.trace; k.trace_step = a.trace_step; k.trace_k = a.trace_k; k.dbg = a.dbg;
    // persistent grid: never more CTAs than can be resident (the rest would only queue behind them without the cross-unit prefetch)
    static int occ[2][2] = {{0, 0}, {0, 0}};
    static int sms = 0;
    const int vi = g_attn_abs_minb == 3 ? 0 : 1, si = a.znew ? 1 : 0;
    if (!occ[vi][si]) {
        int n = 0, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const void* fn = vi == 0 ? (si ? (const void*)attn_abs_kernel<true, 3> : (const void*)attn_abs_kernel<false, 3>)
                                 : (si ? (const void*)attn_abs_kernel<true, 4> : (const void*)attn_abs_kernel<false, 4>);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, 32 * (AW + 1), smem) != cudaSuccess || n < 1) n = 1;
        occ[vi][si] = n;
    }
    const int cap = std::min(max_ctas, occ[vi][si] * (sms > 0 ? sms : 148));
    const int grid = a.batch < cap ? a.batch : cap;
    const dim3 block(32 * (AW + 1));
    if (g_attn_abs_minb == 3) {
        if (a.znew) return launch_pdl(PDL_ATTN_TMA, attn_abs_kernel<true, 3>, dim3(grid), block, smem, st, tm, tm4, k);
        return launch_pdl(PDL_ATTN_TMA, attn_abs_kernel<false, 3>, dim3(grid), block, smem, st, tm, tm4, k);
    }
    if (a.znew) return launch_pdl(PDL_ATTN_TMA, attn_abs_kernel<true, 4>, dim3(grid), block, smem, st, tm, tm4, k);
    return launch_pdl(PDL_ATTN_TMA, attn_abs_kernel<false, 4>, dim3(grid), block, smem, st, tm, tm4, k);
}
