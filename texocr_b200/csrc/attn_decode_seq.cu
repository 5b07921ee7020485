// Decode-step attention of the bf16 generate loop, absorbed ("latent") formulation, ONE WARP PER SEQUENCE.
//
// Same function and interface as attn_abs_kernel (attn_decode_tma.cu; model/attention.py:148-173 with a query length of 1 and
// the key / value projections folded into the query / output projections): for every sequence, all 8 heads attend over the
// same [n, 256] bf16 latent rows (self: the layer's cached LayerNorm'd inputs plus this step's own row, which is appended;
// cross: the encoder memory), C_h = softmax(Q'_h . Z^T / 8) . Z.
//
// Why another kernel: with several batches in flight the decode is bound by SM residency.  The team kernel
// gives a sequence a whole CTA (4 consumer warps that split the 256 columns and exchange partial scores through shared memory at
// a CTA barrier every 32 keys, plus a producer warp): 1/3 of an SM for ~8 us.  Here a sequence is one warp's private job -- its own
// TMA ring, its own mbarriers, no CTA-level synchronisation at all -- so a CTA of 4 warps carries 4 sequences and a sequence costs
// 1/8 of an SM.  Both contractions keep the keys / latent columns on the 16-row M side of mma.sync.m16n8k16 and the 8 heads on the
// N side (no padded rows): per 16-key stage a warp issues 16 + 16 MMAs and as many ldmatrix; the probabilities go from the
// accumulator layout of the score MMA to the B-operand layout of the P.Z MMA with four warp shuffles.
#include <algorithm>

#include "common.cuh"
#include "tc_gemm.h"

int g_attn_seq = 1;          // texocr_set_option("attn_seq"): 1 = this kernel for the absorbed decode attention, 0 = the team kernel

namespace {

constexpr int CH = 16;                 // keys per stage
constexpr int STAGE = CH * 512;        // 16 latent rows x 256 bf16
constexpr int HTILE = CH * 128;        // one 64-column block of a stage, TMA 128B-swizzle layout: (row r, 16-byte chunk c) at r*128 + ((c ^ (r&7)) << 4)
constexpr int SW = 4;                  // warps = sequences in flight per CTA
constexpr int SNS = 2;                 // ring stages per warp (16 KB in flight per sequence; 3 stages measured no better)
constexpr int SEQ_MINB = 3;            // CTAs per SM: 168 registers (no spills), 74 KB of shared memory -> 12 sequences in flight per SM
constexpr int OSTG_ROW = 272;          // output staging: 128 columns + 8 pad, bf16 (conflict-free 16-bit stores)
constexpr int OSTG = 8 * OSTG_ROW;     // bytes per warp: 8 heads x half of the 256 columns
constexpr float SCALE = 0.125f;

struct SeqArgs {
    const bf16* q; int ldq;            // [batch, ldq]: head h at h*256 (absorbed query, unscaled)
    const int* k_off; int uni_nk;      // cross: token offsets [batch + 1], or uni_nk > 0: sequence b = rows [b*uni_nk, (b+1)*uni_nk)
    const bf16* znew; int ldz;         // self: this step's latent rows [batch, ldz]
    bf16* cache; int tcap;             // self: latent cache [batch][tcap][256]; the tensor maps cover it
    const int* step;                   // self: positions already cached (= index of this step's row)
    bf16* o; int ldo;                  // [batch, ldo]: head h at h*256
    int batch;
    int mid_trigger;                   // programmatic dependent released when the warp's last chunk has been issued instead of at entry
    unsigned long long* dbg;           // debug: [0] wait, [1] first stage, [2] loop, [3] epilogue ns of every warp's first unit, [4] count; [5] warp residency, [6] warps
};

TX_DEVINL unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
TX_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
TX_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
TX_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
TX_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "SEQ_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra SEQ_WAIT_DONE;\n\t"
        "bra SEQ_WAIT_LOOP;\n\t"
        "SEQ_WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
TX_DEVINL void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
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
TX_DEVINL void stsm_x4_t(uint32_t addr, uint32_t d0, uint32_t d1, uint32_t d2, uint32_t d3) {
    asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(d0), "r"(d1), "r"(d2), "r"(d3) : "memory");
}
TX_DEVINL void mma_bf16(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
TX_DEVINL void sts_u4(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
TX_DEVINL uint4 lds_u4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
TX_DEVINL void sts_h1(uint32_t addr, float x) {
    const unsigned short h = __bfloat16_as_ushort(__float2bfloat16_rn(x));
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(h) : "memory");
}
TX_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
TX_DEVINL float2 unpack_bf16x2(uint32_t w) { return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w)); }

template <bool SELF>
__global__ void __launch_bounds__(32 * SW, SEQ_MINB) attn_seq_kernel(const __grid_constant__ CUtensorMap tm,
                                                            const __grid_constant__ CUtensorMap tm4, const SeqArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* ring = base + warp * (SNS * STAGE);                                   // this warp's ring
    const uint32_t ostg = smem_u32(base + SW * SNS * STAGE + warp * OSTG);          // this warp's output staging
    uint64_t* full = reinterpret_cast<uint64_t*>(base + SW * SNS * STAGE + SW * OSTG) + warp * SNS;

    if (!a.mid_trigger) pdl_launch_dependents();
    const unsigned long long t_entry = a.dbg ? gtime() : 0ull;
    if (lane == 0) {
        if (warp == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm4) : "memory");
        }
        for (int s = 0; s < SNS; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    const int units = a.batch;
    const int gw = blockIdx.x * SW + warp, nw = gridDim.x * SW;      // sequences are dealt to warps
    int t = 0;
    // rows of unit u: first row in the latent matrix, rows to fetch (self: the t cached rows; this step's own row is added below)
    auto unit_rows = [&](int u, int& row0, int& nc) {
        if (SELF) { row0 = u * a.tcap; nc = t; }
        else if (a.uni_nk > 0) { row0 = u * a.uni_nk; nc = a.uni_nk; }
        else { row0 = ldcg_i32(a.k_off + u); nc = ldcg_i32(a.k_off + u + 1) - row0; }
    };
    // lane 0 streams this warp's chunks: chunk c of a unit -> ring slot k % SNS (k = running chunk number of the warp)
    auto issue = [&](int k, int row0, int nc, int c) {
        const int s = k % SNS;
        uint8_t* st = ring + s * STAGE;
        const int left = nc - c * CH, r = row0 + c * CH;
        if (left >= CH) {
            mbar_expect_tx(&full[s], STAGE);
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) tma_load_2d(&tm, &full[s], st + cb * HTILE, 64 * cb, r);
        } else {               // tail: 4-row boxes (none at all when only this step's own row is left)
            const int n4 = left > 0 ? (left + 3) >> 2 : 0;
            mbar_expect_tx(&full[s], n4 * 4 * 512);
            for (int j = 0; j < n4; ++j)
#pragma unroll
                for (int cb = 0; cb < 4; ++cb) tma_load_2d(&tm4, &full[s], st + cb * HTILE + j * 512, 64 * cb, r + 4 * j);
        }
    };
    int iu = gw, ic = 0, i_row0 = 0, i_nc = 0, i_nch = 0, issued = 0;       // issue cursor (every lane tracks it, lane 0 acts)
    bool released = !a.mid_trigger;
    auto refill = [&](int consumed) {       // chunks below `consumed` are released: up to SNS chunks may be in the ring
        while (iu < units && issued < consumed + SNS) {
            if (ic == 0) { unit_rows(iu, i_row0, i_nc); i_nch = ((SELF ? i_nc + 1 : i_nc) + CH - 1) / CH; }
            if (lane == 0) issue(issued, i_row0, i_nc, ic);
            ++issued;
            if (++ic == i_nch) { ic = 0; iu += nw; }
        }
        // the warp's last chunk is on its way: at most SNS stages and the output are left -- about the dependent's launch + prologue
        if (!released && iu >= units && consumed > 0) { pdl_launch_dependents(); released = true; }
    };
    // cross: the memory rows and the token offsets were written before the generate loop started -- stream before the wait
    if (!SELF) refill(0);
    pdl_wait();
    if (SELF) {
        t = ldcg_i32(a.step);
        if (lane == 0) asm volatile("fence.proxy.async.global;" ::: "memory");     // cache rows were appended by generic-proxy stores of earlier steps
        refill(0);
    }
    const unsigned long long t_ready = a.dbg ? gtime() : 0ull;
    unsigned long long t_first = 0ull;

    const int g = lane >> 2, tq = lane & 3;
    const int lm_r = lane & 7, lm_m = lane >> 3;
    // ldmatrix offsets inside a 16 x 64 tile (see attn_abs_kernel): k-step / m-tile j: m0 = keys 0-7 / chunk 2j, m1 = keys 8-15 / 2j,
    // m2 = keys 0-7 / 2j+1, m3 = keys 8-15 / 2j+1
    uint32_t off_z[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { const int r = (lm_m & 1) * 8 + lm_r; off_z[j] = r * 128 + (((2 * j + (lm_m >> 1)) ^ (r & 7)) << 4); }
    // shuffle sources of the probability transpose: (key 2tq [+1], heads 2(g/2), 2(g/2)+1) live in lane (2tq [+1]) * 4 + g/2
    const int src_a = (2 * tq) * 4 + (g >> 1), src_b = src_a + 4;
    const bool odd_head = g & 1;

    int it = 0;
    for (int u = gw; u < units; u += nw) {
        // B fragments of Q'^T, scaled by 1/8 (exact): k-step ks (16 columns): (cols 16ks+2tq,+1; head g), (cols 16ks+8+2tq,+1; head g)
        // The 8 x 256 query block (4 KB) comes in with 16-byte coalesced loads, goes through the warp's staging tile (two halves of 128
        // columns, head rows 272 bytes apart) and is picked up by ldmatrix in fragment layout: one L2 round trip instead of 32 scattered ones.
        uint32_t qa[32];
        {
            const bf16* qp = a.q + (size_t)u * a.ldq;
            uint4 qv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {      // chunk id = i*32 + lane of 256 chunks: head = id >> 5, 16-byte chunk of the head row = id & 31
                const int id = i * 32 + lane;
                qv[i] = ldcg_u4(qp + (id >> 5) * 256 + (id & 31) * 8);
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int id = i * 32 + lane, hd = id >> 5, ch = id & 31;
                    if ((ch >> 4) == half) sts_u4(ostg + hd * OSTG_ROW + (ch & 15) * 16, qv[i]);
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 4; ++j) {      // x4: the 8 x 8 blocks of columns 32j .. 32j+31 of this half = k-steps 8*half + 2j, +1
                    const int kk = 8 * half + 2 * j;
                    ldsm_x4(ostg + lm_r * OSTG_ROW + (4 * j + lm_m) * 16, qa[2 * kk], qa[2 * kk + 1], qa[2 * kk + 2], qa[2 * kk + 3]);
                }
                __syncwarp();
            }
        }
        uint4 zrow = make_uint4(0u, 0u, 0u, 0u);
        if (SELF) {      // this step's latent row, columns 8*lane .. +7: append to the cache, keep for the last stage
            zrow = ldcg_u4(a.znew + (size_t)u * a.ldz + lane * 8);
            *reinterpret_cast<uint4*>(a.cache + ((size_t)u * a.tcap + t) * 256 + lane * 8) = zrow;
            asm volatile("fence.proxy.async.global;" ::: "memory");      // later steps read the row through the async proxy (TMA)
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float2 f = unpack_bf16x2(qa[i]);
            qa[i] = pack_bf16x2(f.x * SCALE, f.y * SCALE);
        }
        int nk;
        if (SELF) nk = t + 1;
        else if (a.uni_nk > 0) nk = a.uni_nk;
        else nk = ldcg_i32(a.k_off + u + 1) - ldcg_i32(a.k_off + u);
        const int nchunk = (nk + CH - 1) / CH;
        float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;       // heads 2tq, 2tq+1 (max: warp-uniform per head; sums: this lane's keys only)
        float o[16][4];      // m-tile mt (columns 16mt + g, 16mt + g + 8) x heads (2tq, 2tq + 1)
#pragma unroll
        for (int mt = 0; mt < 16; ++mt) { o[mt][0] = o[mt][1] = o[mt][2] = o[mt][3] = 0.f; }
        for (int c = 0; c < nchunk; ++c) {
            const int s = it % SNS;
            mbar_wait(&full[s], (it / SNS) & 1);
            const uint32_t st = smem_u32(ring + s * STAGE);
            if (a.dbg && t_first == 0ull) t_first = gtime();
            const bool last = c == nchunk - 1;
            if (SELF && last) {      // key t of the sequence = this step's own row, row t % 16 of the last stage; lane = 16-byte chunk of the 512-byte row
                const int r = t & (CH - 1);
                sts_u4(st + (lane >> 3) * HTILE + r * 128 + (((lane & 7) ^ (r & 7)) << 4), zrow);
                __syncwarp();
            }
            // ---- S^T[16 keys x 8 heads] = Z . Q'^T : 16 k-steps of 16 columns, four independent accumulation chains (one per column block)
            float sc[4][4];
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) {
                sc[cb][0] = sc[cb][1] = sc[cb][2] = sc[cb][3] = 0.f;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    uint32_t a0, a1, a2, a3;
                    ldsm_x4(st + cb * HTILE + off_z[ks], a0, a1, a2, a3);
                    mma_bf16(sc[cb], a0, a1, a2, a3, qa[2 * (4 * cb + ks)], qa[2 * (4 * cb + ks) + 1]);
                }
            }
            float s0 = (sc[0][0] + sc[1][0]) + (sc[2][0] + sc[3][0]);       // (key g,   head 2tq)
            float s1 = (sc[0][1] + sc[1][1]) + (sc[2][1] + sc[3][1]);       // (key g,   head 2tq+1)
            float s2 = (sc[0][2] + sc[1][2]) + (sc[2][2] + sc[3][2]);       // (key g+8, head 2tq)
            float s3 = (sc[0][3] + sc[1][3]) + (sc[2][3] + sc[3][3]);       // (key g+8, head 2tq+1)
            uint32_t vm_lo = 0xffffffffu, vm_hi = 0xffffffffu;      // masks of the transposed Z fragments: keys (2tq, 2tq+1) of 0-7 / of 8-15
            if (last) {      // keys past the end of the sequence: score -> -inf; their Z rows carry p = 0, but 0 * NaN = NaN -> cleared below
                const int kg = c * CH + g;
                if (kg >= nk) { s0 = -INFINITY; s1 = -INFINITY; }
                if (kg + 8 >= nk) { s2 = -INFINITY; s3 = -INFINITY; }
                const int k0 = c * CH + 2 * tq;
                vm_lo = (k0 < nk ? 0x0000ffffu : 0u) | (k0 + 1 < nk ? 0xffff0000u : 0u);
                vm_hi = (k0 + 8 < nk ? 0x0000ffffu : 0u) | (k0 + 9 < nk ? 0xffff0000u : 0u);
            }
            // per-head maximum over the 16 keys: over this lane's two keys, then over the 8 lanes that share tq
            float x0 = fmaxf(s0, s2), x1 = fmaxf(s1, s3);
#pragma unroll
            for (int sh = 4; sh < 32; sh <<= 1) {
                x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, sh));
                x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, sh));
            }
            const float n0 = fmaxf(m0, x0), n1 = fmaxf(m1, x1);      // finite: every stage holds at least one valid key
            const float c0 = __expf(m0 - n0), c1 = __expf(m1 - n1);
            m0 = n0; m1 = n1;
            const float p0 = __expf(s0 - n0), p1 = __expf(s1 - n1), p2 = __expf(s2 - n0), p3 = __expf(s3 - n1);
            l0 = l0 * c0 + (p0 + p2);
            l1 = l1 * c1 + (p1 + p3);
            if (c0 != 1.f || c1 != 1.f) {      // warp-uniform per tq group is not guaranteed: plain per-lane rescale of the lane's own heads
#pragma unroll
                for (int mt = 0; mt < 16; ++mt) { o[mt][0] *= c0; o[mt][1] *= c1; o[mt][2] *= c0; o[mt][3] *= c1; }
            }
            // probabilities: accumulator layout (key g | g+8; heads 2tq, 2tq+1) -> B operand layout (keys 2tq, 2tq+1 | +8, +9; head g)
            const uint32_t w_lo = pack_bf16x2(p0, p1), w_hi = pack_bf16x2(p2, p3);      // (head 2tq | head 2tq+1) of key g / key g+8
            const uint32_t a_lo = __shfl_sync(0xffffffffu, w_lo, src_a), b_lo = __shfl_sync(0xffffffffu, w_lo, src_b);
            const uint32_t a_hi = __shfl_sync(0xffffffffu, w_hi, src_a), b_hi = __shfl_sync(0xffffffffu, w_hi, src_b);
            const uint32_t pb0 = odd_head ? ((a_lo >> 16) | (b_lo & 0xffff0000u)) : ((a_lo & 0xffffu) | (b_lo << 16));      // (keys 2tq, 2tq+1; head g)
            const uint32_t pb1 = odd_head ? ((a_hi >> 16) | (b_hi & 0xffff0000u)) : ((a_hi & 0xffffu) | (b_hi << 16));      // (keys 2tq+8, +9; head g)
            // ---- C^T[256 cols x 8 heads] += Z^T . P^T : 16 m-tiles of 16 columns, one k-step of 16 keys
#pragma unroll
            for (int mt = 0; mt < 16; ++mt) {
                uint32_t z0, z1, z2, z3;        // transposed blocks: keys 0-7 / chunk 2j, keys 8-15 / 2j, keys 0-7 / 2j+1, keys 8-15 / 2j+1
                ldsm_x4_t(st + (mt >> 2) * HTILE + off_z[mt & 3], z0, z1, z2, z3);
                if (last) { z0 &= vm_lo; z2 &= vm_lo; z1 &= vm_hi; z3 &= vm_hi; }
                mma_bf16(o[mt], z0, z2, z1, z3, pb0, pb1);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy reads of the stage before its async-proxy refill
            __syncwarp();
            ++it;
            refill(it);
        }
        const unsigned long long t_loop = a.dbg ? gtime() : 0ull;
        // softmax denominators: sum the lanes' partial sums over the 8 lanes that share tq
#pragma unroll
        for (int sh = 4; sh < 32; sh <<= 1) {
            l0 += __shfl_xor_sync(0xffffffffu, l0, sh);
            l1 += __shfl_xor_sync(0xffffffffu, l1, sh);
        }
        const float i0 = 1.0f / l0, i1 = 1.0f / l1;
        // output: normalise, transpose through the warp's staging tile with stmatrix.trans (the accumulator fragment of an m-tile is an
        // 8 x 8 block [column][head]; stored transposed it is 8 head rows of 8 columns = 16 bytes), two halves of 128 columns, then
        // 16 bytes per lane to out[u][head * 256 + col]
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int m2 = 0; m2 < 4; ++m2) {      // two m-tiles (32 columns) per stmatrix.x4
                const int mt = half * 8 + 2 * m2;
                const uint32_t r0 = pack_bf16x2(o[mt][0] * i0, o[mt][1] * i1), r1 = pack_bf16x2(o[mt][2] * i0, o[mt][3] * i1);
                const uint32_t r2 = pack_bf16x2(o[mt + 1][0] * i0, o[mt + 1][1] * i1), r3 = pack_bf16x2(o[mt + 1][2] * i0, o[mt + 1][3] * i1);
                // matrix i (lanes 8i .. 8i+7 give the addresses of its 8 head rows): columns 32*m2 + 8i .. +7 of this half
                stsm_x4_t(ostg + lm_r * OSTG_ROW + (4 * m2 + lm_m) * 16, r0, r1, r2, r3);
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {      // 8 heads x 256 bytes = 128 chunks of 16 bytes
                const int id = i * 32 + lane, hd = id >> 4, ch = id & 15;
                const uint4 v = lds_u4(ostg + hd * OSTG_ROW + ch * 16);
                *reinterpret_cast<uint4*>(a.o + (size_t)u * a.ldo + hd * 256 + half * 128 + ch * 8) = v;
            }
            __syncwarp();
        }
        if (a.dbg && lane == 0 && u == gw) {
            atomicAdd(a.dbg + 0, t_ready - t_entry); atomicAdd(a.dbg + 1, t_first - t_ready); atomicAdd(a.dbg + 2, t_loop - t_first);
            atomicAdd(a.dbg + 3, gtime() - t_loop); atomicAdd(a.dbg + 4, 1ull);
        }
    }
    if (a.dbg && lane == 0 && gw < units) { atomicAdd(a.dbg + 5, gtime() - t_entry); atomicAdd(a.dbg + 6, 1ull); }      // warp residency, all its units
}

}  // namespace

// Same contract as launch_attn_abs (kernels.h).  max_ctas bounds the persistent grid; every warp of a CTA owns a sequence.
cudaError_t launch_attn_seq(const AttnAbsArgs& a, int max_ctas, cudaStream_t st) {
    if (a.batch <= 0) return cudaSuccess;
    if (a.ldq % 8 != 0 || a.ldo % 8 != 0 || (a.znew && a.ldz % 8 != 0)) return cudaErrorInvalidValue;
    CUtensorMap tm, tm4;
    cudaError_t e = tma_map_2d_bf16(a.latent, a.latent_rows, 256, 256, CH, 64, 1, &tm);
    if (e != cudaSuccess) return e;
    if ((e = tma_map_2d_bf16(a.latent, a.latent_rows, 256, 256, 4, 64, 1, &tm4)) != cudaSuccess) return e;
    const size_t smem = 1024 + (size_t)SW * SNS * STAGE + (size_t)SW * OSTG + (size_t)SW * SNS * 8 + 64;
    typedef void (*Kern)(CUtensorMap, CUtensorMap, SeqArgs);
    static const Kern kerns[2] = {attn_seq_kernel<false>, attn_seq_kernel<true>};
    static int occ[2] = {0, 0};
    static int sms = 0;
    const int si = a.znew ? 1 : 0;
    if (!occ[si]) {
        int n = 0, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if ((e = cudaFuncSetAttribute(kerns[si], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kerns[si], 32 * SW, smem) != cudaSuccess || n < 1) n = 1;
        occ[si] = n;
    }
    SeqArgs k{};
    k.q = (const bf16*)a.q; k.ldq = a.ldq; k.k_off = a.k_off; k.uni_nk = a.znew ? 0 : a.uni_nk; k.o = (bf16*)a.o; k.ldo = a.ldo; k.batch = a.batch;
    k.znew = (const bf16*)a.znew; k.ldz = a.ldz; k.cache = (bf16*)const_cast<void*>(a.latent); k.tcap = a.tcap; k.step = a.step;
    k.dbg = a.dbg; k.mid_trigger = (g_texocr_pdl_mid >> 1) & 1;
    const int want = (a.batch + SW - 1) / SW;
    const int cap = std::max(1, std::min(max_ctas, occ[si] * (sms > 0 ? sms : 148)));
    const int grid = want < cap ? want : cap;
    return launch_pdl(PDL_ATTN_TMA, kerns[si], dim3(grid), dim3(32 * SW), smem, st, tm, tm4, k);
}
