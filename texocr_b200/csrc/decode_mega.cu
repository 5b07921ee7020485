// Cluster-persistent greedy decode loop of the bf16 tier: ONE kernel runs `nsteps` complete decoder steps
// (model/decoder.py:84-118 in the temperature -> 0 limit; model/attention.py:148-259 per layer).
//
// The batch is cut into G groups of at most 80 sequences; every group is owned by one 16-CTA thread-block cluster for the
// whole launch (a B200 holds 7 such clusters at one CTA per SM), so groups never synchronise with each other and the only
// synchronisation inside a step is the hardware cluster barrier between its 34 phases.  Per layer:
//   [LN,LN -> QKV GEMM] | self-attention | [out GEMM, GLU, +res] | [LN,LN -> Q GEMM] | cross-attention | [out GEMM, GLU, +res]
//   | [LN,LN -> FF1 GEMM, GeGLU] | [FF2 GEMM, +res]            then [final LN -> logits GEMM -> partial argmax] | argmax.
// GEMM phases: the 16 CTAs split the N dimension of the weight matrix; every CTA holds all rows of the group's activation
// (bf16, K chunks of 256 in two shared-memory buffers; LayerNorm-fed GEMMs rebuild the cheap row-wise LayerNorm locally)
// and its [N/16][K] weight slice (cp.async, XOR-swizzled 16-byte chunks; the slice of the NEXT phase is prefetched while
// the current epilogue and the cluster barrier run), and computes with mma.sync.m16n8k16 (bf16 in, fp32 accumulate) -- M
// is at most 80 and the work per CTA is ~3 MFLOP, nothing a tcgen05 tile would help.  Activations travel between phases
// through L2 (ld/st .cg; the cluster barrier is the release/acquire point).
// Attention phases: CTA c streams the K/V of sequences c, c+16, ... of its group for all 8 heads: lane 0 of warp 0
// issues 3-D TMA loads (16 keys x 64 dims x 8 heads per box, 128-byte swizzle) into a 3 x 32 KB ring, warp h consumes
// head h with the flash-decoding inner loop of attn_decode_tma.cu.  This is the HBM-bound part: 2 KB per cached key per
// sequence per layer.
#include <algorithm>

#include "common.cuh"
#include "decode_mega.h"
#include "tc_gemm.h"

namespace {

constexpr int CS = MEGA_CLUSTER, RG = MEGA_ROWS_PER_GROUP;      // 16 CTAs, at most 80 rows
constexpr int MT_ALL = RG / 16;           // 5 m-tiles
constexpr int MAXU = RG / CS;             // sequences per CTA in the attention phases
constexpr int NTHR = 256;                 // 8 warps (= heads); 2 per SM sub-partition, so up to 255 registers per thread
constexpr int CH = 16, NS = 4;            // keys per ring stage, ring stages (3 in the ring + activation buffer 1)
constexpr int HT = CH * 128;              // one head's K (or V) tile of a stage: 16 rows x 128 B
constexpr int HALF = 8 * HT;              // the K (or V) tiles of the 8 heads
constexpr int STAGE = 2 * HALF;           // 32 KB
constexpr int RING = 3 * STAGE;
constexpr int WB = 48 * 1024;             // weight slice (the second half of the FF1 slice lives in activation buffer 1)
constexpr int ABUF = RG * 512;            // one activation K chunk: 80 rows x 256 bf16
constexpr int AB = 2 * ABUF;
constexpr int MISC = 512;
constexpr int SMEM_BYTES = RING + WB + AB + MISC + 1024;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
static_assert(ABUF >= STAGE && ABUF % 1024 == 0 && WB % 1024 == 0, "activation buffer 1 doubles as the fourth K/V stage");
constexpr float SCALE = 0.125f, LN_EPS = 1e-5f;
enum { EPI_STORE16, EPI_GLU, EPI_GEGLU, EPI_RES, EPI_LOGITS };
enum { PRO_EMBED, PRO_LN2, PRO_FIN };

struct Params {
    CUtensorMap tm_self, tm_self4, tm_cross, tm_cross4;
    MegaArgs a;
};

// ------------------------------------------------------------------------------------------------ PTX helpers
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
TX_DEVINL void tma_load_3d(const CUtensorMap* map, uint64_t* bar, uint32_t dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
TX_DEVINL void ldsm_x4(uint32_t addr, uint32_t& d0, uint32_t& d1, uint32_t& d2, uint32_t& d3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(d0), "=r"(d1), "=r"(d2), "=r"(d3) : "r"(addr));
}
TX_DEVINL void ldsm_x2(uint32_t addr, uint32_t& d0, uint32_t& d1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(d0), "=r"(d1) : "r"(addr));
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
TX_DEVINL float2 unpack_bf16x2(uint32_t w) { return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w)); }
TX_DEVINL void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
TX_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> TX_DEVINL void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
TX_DEVINL void st_shared_zero16(uint32_t dst) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
}
TX_DEVINL void st_shared16(uint32_t dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
TX_DEVINL uint32_t ld_shared32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
TX_DEVINL uint4 ld_shared128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
// every thread of every CTA of the cluster; orders the global-memory traffic of the phases on both sides
TX_DEVINL void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
TX_DEVINL unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// phase timing (debug, MegaArgs::dbg_time): 0 qkv, 1 out_s, 2 q_c, 3 out_c, 4 ff1, 5 ff2, 6 logits, 7 self-attn, 8 cross-attn,
// 9 argmax, 10 cluster barriers, 11 whole kernel, 12 prologues, 13 operand waits, 14 MMA loops (0..8 then hold the epilogues)
#define TICK(idx)                                                                                                        \
    do {                                                                                                                 \
        if (a.dbg_time && threadIdx.x == 0) { const unsigned long long n_ = gtime(); tacc[idx] += n_ - tacc[16]; tacc[16] = n_; } \
    } while (0)

// ------------------------------------------------------------------------------------------------ weight slices
struct WDesc { const bf16* w; int n0, nrows, nvalid, K; };

// slice of the gi-th GEMM of a step (6 per layer, then the logits) owned by CTA `crank`
__device__ __noinline__ WDesc w_desc(const MegaArgs& a, int gi, int crank) {
    WDesc d;
    if (gi >= 6 * a.L) {
        d.w = (const bf16*)a.w_logits; d.n0 = crank * 64; d.nrows = 64; d.K = 256;
        const int nv = a.V - d.n0;
        d.nvalid = nv < 0 ? 0 : (nv > 64 ? 64 : nv);
        return d;
    }
    const MegaLayerW& w = a.layer[gi / 6];
    switch (gi % 6) {
        case 0: d.w = (const bf16*)w.wqkv; d.nrows = 96; d.K = 256; break;
        case 1: d.w = (const bf16*)w.wo_s; d.nrows = 32; d.K = 512; break;
        case 2: d.w = (const bf16*)w.wq_c; d.nrows = 32; d.K = 256; break;
        case 3: d.w = (const bf16*)w.wo_c; d.nrows = 32; d.K = 512; break;
        case 4: d.w = (const bf16*)w.w1; d.nrows = 128; d.K = 256; break;
        default: d.w = (const bf16*)w.w2; d.nrows = 16; d.K = 1024; break;
    }
    d.n0 = crank * d.nrows; d.nvalid = d.nrows;
    return d;
}

// rows [r0, r1) of the slice -> dst (row r at (r - r0) * 2K bytes, 16-byte chunk c at chunk position c ^ (r & 7))
TX_DEVINL void load_w_rows(uint8_t* dst_buf, const WDesc& d, int r0, int r1, int tid) {
    const int sh = d.K == 256 ? 5 : (d.K == 512 ? 6 : 7);       // log2(chunks per row)
    const int total = (r1 - r0) << sh;
    const uint32_t base = smem_u32(dst_buf);
#pragma unroll 1
    for (int i = tid; i < total; i += NTHR) {
        const int rr = i >> sh, ch = i & ((1 << sh) - 1);
        const uint32_t dst = base + rr * (d.K * 2) + ((ch ^ (rr & 7)) << 4);
        if (r0 + rr < d.nvalid) cp_async16(dst, d.w + (size_t)(d.n0 + r0 + rr) * d.K + ch * 8);
        else st_shared_zero16(dst);
    }
}
// whole slice: into wbuf; the 128-row FF1 slice puts rows 64..127 into activation buffer 1 (free while K = 256)
TX_DEVINL void load_w(uint8_t* wbuf, uint8_t* abuf1, const WDesc& d, int tid) {
    if (d.nrows <= 96) load_w_rows(wbuf, d, 0, d.nrows, tid);
    else { load_w_rows(wbuf, d, 0, 64, tid); load_w_rows(abuf1, d, 64, d.nrows, tid); }
    cp_async_commit();
}

// bf16 activation rows x 256 columns starting at src -> one activation buffer (same swizzle); rows past the group zero
TX_DEVINL void copy_a_chunk(uint8_t* abuf, const bf16* src, int ld, int row0, int rows, int mrows, int tid) {
    const uint32_t base = smem_u32(abuf);
#pragma unroll 1
    for (int i = tid; i < mrows * 32; i += NTHR) {
        const int r = i >> 5, ch = i & 31;
        const uint32_t dst = base + r * 512 + ((ch ^ (r & 7)) << 4);
        if (r < rows) cp_async16(dst, src + (size_t)(row0 + r) * ld + ch * 8);
        else st_shared_zero16(dst);
    }
    cp_async_commit();
}

// ------------------------------------------------------------------------------------------------ shared memory / group geometry
// The kernel is a sequence of ~35 phases per step, each executed once: its instruction footprint has to stay small (the
// first version, with every phase inlined and unrolled, was 350 KB of SASS and spent most of its time in instruction-cache
// misses).  Every phase is therefore ONE non-inlined function with run-time parameters; each recomputes this layout.
struct Sm {
    uint8_t* ring; uint8_t* wbuf; uint8_t* abuf;
    uint64_t* full; uint64_t* empty;
    struct UnitMeta* meta;
    unsigned long long* tacc;
};
struct UnitMeta { int key0, nk, z, pad; };
TX_DEVINL Sm sm_layout() {
    extern __shared__ uint8_t smem_raw[];
    Sm s;
    s.ring = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    s.wbuf = s.ring + RING;
    s.abuf = s.wbuf + WB;
    s.full = reinterpret_cast<uint64_t*>(s.abuf + AB);
    s.empty = s.full + NS;
    s.meta = reinterpret_cast<UnitMeta*>(s.empty + NS + 2);          // [MAXU]
    s.tacc = reinterpret_cast<unsigned long long*>(s.meta + 8);      // [18]
    return s;
}
struct Grp { int crank, grp, row0, rows, mrows, mtiles; };
TX_DEVINL Grp grp_of(const MegaArgs& a) {
    Grp g;
    g.crank = blockIdx.x % CS; g.grp = blockIdx.x / CS;
    g.row0 = g.grp * a.B / a.G;                  // B < 2^15
    g.rows = (g.grp + 1) * a.B / a.G - g.row0;
    g.mtiles = (g.rows + 15) >> 4; g.mrows = g.mtiles * 16;
    return g;
}
constexpr int CST_LD = 136;          // fp32 accumulator staging in the (idle) K/V ring: [mrows][CST_LD], up to 128 columns

// ------------------------------------------------------------------------------------------------ prologues
// LayerNorm of 4 rows at once (one row per warp pass, 8 elements per lane): the four shuffle-reduction chains are
// independent, which is what hides their latency at 2 warps per scheduler.
TX_DEVINL void ln8x4(float (&v)[4][8], const float* g, const float* b) {
    float s[4], q[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        s[i] = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s[i] += v[i][k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        s[i] *= (1.0f / 256);
        q[i] = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { const float d = v[i][k] - s[i]; q[i] = fmaf(d, d, q[i]); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int i = 0; i < 4; ++i) q[i] += __shfl_xor_sync(0xffffffffu, q[i], o);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float rstd = 1.0f / sqrtf(q[i] * (1.0f / 256) + LN_EPS);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[i][k] = (v[i][k] - s[i]) * rstd * g[k] + b[k];
    }
}

// A operand of the LayerNorm-fed GEMMs, rebuilt by every CTA for all rows of the group (model/attention.py:242-259:
// x = LN(s) is the residual input, xn = LN(x) feeds the sub-layer; the first block sees x = embedding, xn = LN(x)).
// CTA `crank` also publishes the fp32 x of rows r with r % 16 == crank for the residual add two phases later.
__device__ __noinline__ void prologue_rows(const Params& p, int mode, int t) {
    const MegaArgs& a = p.a;
    const Sm sm = sm_layout();
    const Grp gr = grp_of(a);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int col = lane * 8;
    // fp32 rows -> the (idle) K/V ring, asynchronously: [rows][256]
    const uint32_t stg = smem_u32(sm.ring);
    if (mode == PRO_EMBED) {
#pragma unroll 1
        for (int r = warp; r < gr.rows; r += 8) {
            long id = (long)__ldcg(reinterpret_cast<const long long*>(a.cur_tok) + gr.row0 + r);
            id = id < 0 ? 0 : (id >= a.V ? a.V - 1 : id);
            const float* src = a.tok_emb + (size_t)id * 256;
            cp_async16(stg + r * 1024 + lane * 16, src + lane * 4);
            cp_async16(stg + r * 1024 + 512 + lane * 16, src + 128 + lane * 4);
        }
    } else {
#pragma unroll 1
        for (int i = tid; i < gr.rows * 64; i += NTHR) {
            const int r = i >> 6, ch = i & 63;
            cp_async16(stg + r * 1024 + ch * 16, a.s + (size_t)(gr.row0 + r) * 256 + ch * 4);
        }
    }
    cp_async_commit();
    float g[8], b[8], pe[8];
    ld8((mode == PRO_FIN ? a.fin_g : a.ln_g) + col, g);
    ld8((mode == PRO_FIN ? a.fin_b : a.ln_b) + col, b);
    if (mode == PRO_EMBED) ld8(a.pos_emb + (size_t)t * 256 + col, pe);
    cp_async_wait<0>();
    __syncthreads();
    const uint32_t base = smem_u32(sm.abuf);
#pragma unroll 1
    for (int rb = warp; rb < gr.mrows; rb += 32) {       // rows rb, rb + 8, rb + 16, rb + 24 (zeros past the group)
        float v[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = rb + 8 * i;
            uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
            if (r < gr.rows) { lo = ld_shared128(stg + r * 1024 + lane * 32); hi = ld_shared128(stg + r * 1024 + lane * 32 + 16); }
            v[i][0] = __uint_as_float(lo.x); v[i][1] = __uint_as_float(lo.y); v[i][2] = __uint_as_float(lo.z); v[i][3] = __uint_as_float(lo.w);
            v[i][4] = __uint_as_float(hi.x); v[i][5] = __uint_as_float(hi.y); v[i][6] = __uint_as_float(hi.z); v[i][7] = __uint_as_float(hi.w);
            if (mode == PRO_EMBED) {
#pragma unroll
                for (int k = 0; k < 8; ++k) v[i][k] += pe[k];
            }
        }
        if (mode == PRO_LN2) ln8x4(v, g, b);
        if (mode != PRO_FIN) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = rb + 8 * i;
                if (r < gr.rows && (r % CS) == gr.crank) {
                    float* xp = a.x + (size_t)(gr.row0 + r) * 256 + col;
                    st4(xp, make_float4(v[i][0], v[i][1], v[i][2], v[i][3]));
                    st4(xp + 4, make_float4(v[i][4], v[i][5], v[i][6], v[i][7]));
                }
            }
        }
        ln8x4(v, g, b);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = rb + 8 * i;
            if (r < gr.mrows) {
                const bool ok = r < gr.rows;
                st_shared16(base + r * 512 + ((lane ^ (r & 7)) << 4), ok ? pack_bf16x2(v[i][0], v[i][1]) : 0u, ok ? pack_bf16x2(v[i][2], v[i][3]) : 0u,
                            ok ? pack_bf16x2(v[i][4], v[i][5]) : 0u, ok ? pack_bf16x2(v[i][6], v[i][7]) : 0u);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ GEMM core
// One routine for every GEMM of the step (code size matters more than the last bit of MMA efficiency here): the CTA's
// weight slice is processed in passes of 32 weight rows; warp (mg, ng) = (warp & 1, warp >> 1) owns the m-tiles mg, mg+2,
// mg+4 and the n-tile ng of the pass.  Multi-pass GEMMs (QKV 3, FF1 4, logits 2) have K = 256 (one activation chunk,
// resident in buffer 0); the K = 512 / 1024 GEMMs are single-pass and stream their activation chunks through both
// buffers.  FF2's slice has only 16 rows: n-tiles 2, 3 of its pass are idle.
// A chunk 0 is already in activation buffer 0 (prologue) when copy_src == nullptr.  The weight slice `wd` is in flight
// (cp.async); `next` is prefetched after the MMAs.  The fp32 accumulators are left in the K/V ring (idle during GEMM
// phases) as cst[row][column of the slice] for the rolled epilogue.
__device__ __noinline__ void gemm_mma(const Params& p, WDesc wd, WDesc next, int have_next, const bf16* copy_src, int copy_ld) {
    const MegaArgs& a = p.a;
    const Sm sm = sm_layout();
    const Grp gr = grp_of(a);
    unsigned long long* tacc = sm.tacc;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int mg = warp & 1, ng = warp >> 1;
    const int nchunk = wd.K >> 8, npass = (wd.nrows + 31) >> 5;
    const int w_stride = wd.K * 2;
    float* cst = reinterpret_cast<float*>(sm.ring);
    const int g = lane >> 2, tq = lane & 3;
    const int a_r = lane & 15, a_hi = lane >> 4, b_hi = (lane >> 3) & 1;
    if (copy_src) {
        copy_a_chunk(sm.abuf, copy_src, copy_ld, gr.row0, gr.rows, gr.mrows, tid);
        if (nchunk > 1) copy_a_chunk(sm.abuf + ABUF, copy_src + 256, copy_ld, gr.row0, gr.rows, gr.mrows, tid);
    }
    float acc[3][4];
#pragma unroll 1
    for (int q = 0; q < npass; ++q) {
        const int wr = 32 * q + 8 * ng;                       // first weight row of this warp's n-tile inside the slice
        const bool active = wr < wd.nrows;
        // rows 64.. of the 128-row FF1 slice live in activation buffer 1
        const uint32_t w_base = (wd.nrows > 96 && wr >= 64) ? smem_u32(sm.abuf + ABUF) + (wr - 64) * w_stride : smem_u32(sm.wbuf) + wr * w_stride;
        const int br = lane & 7;                              // (wr & 7) == 0: the row's swizzle phase is lane & 7
        const uint32_t w_row = w_base + br * w_stride;
#pragma unroll
        for (int i = 0; i < 3; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll 1
        for (int c = 0; c < nchunk; ++c) {
            if (q == 0) {
                if (c + 1 < nchunk) cp_async_wait<1>(); else cp_async_wait<0>();      // chunk c (and the weights) have landed
                __syncthreads();
                TICK(13);
            }
            const uint32_t ab = smem_u32(sm.abuf + (c & 1) * ABUF);
            if (active) {
#pragma unroll 4
                for (int ks = 0; ks < 16; ++ks) {
                    uint32_t b0, b1;
                    ldsm_x2(w_row + (((c * 32 + 2 * ks + b_hi) ^ br) << 4), b0, b1);
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const int mt = mg + 2 * i;
                        if (mt < gr.mtiles) {
                            const int ar = mt * 16 + a_r;
                            uint32_t a0, a1, a2, a3;
                            ldsm_x4(ab + ar * 512 + (((2 * ks + a_hi) ^ (ar & 7)) << 4), a0, a1, a2, a3);
                            mma_bf16(acc[i], a0, a1, a2, a3, b0, b1);
                        }
                    }
                }
            }
            if (c + 2 < nchunk) {
                __syncthreads();                                  // every warp is done with buffer c & 1
                copy_a_chunk(sm.abuf + (c & 1) * ABUF, copy_src + (c + 2) * 256, copy_ld, gr.row0, gr.rows, gr.mrows, tid);
            }
        }
        // accumulators -> cst: c0,c1 = (row 16 mt + g, columns n, n+1), c2,c3 = row + 8
        if (active) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int mt = mg + 2 * i;
                if (mt < gr.mtiles) {
                    float* d = cst + (mt * 16 + g) * CST_LD + wr + 2 * tq;
                    *reinterpret_cast<float2*>(d) = make_float2(acc[i][0], acc[i][1]);
                    *reinterpret_cast<float2*>(d + 8 * CST_LD) = make_float2(acc[i][2], acc[i][3]);
                }
            }
        }
    }
    __syncthreads();                                          // weight / activation buffers free, accumulators staged
    TICK(14);
    if (have_next) load_w(sm.wbuf, sm.abuf + ABUF, next, tid);
}

// Rolled epilogue over the staged accumulators of the CTA's [rows][nrows] slice (weight rows n0 .. n0 + nrows).
__device__ __noinline__ void epilogue(const Params& p, int epi, int n0, int nrows, const float* bias, bf16* out16, int ld16) {
    const MegaArgs& a = p.a;
    const Sm sm = sm_layout();
    const Grp gr = grp_of(a);
    const float* cst = reinterpret_cast<const float*>(sm.ring);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    __syncthreads();                                          // staged accumulators visible
    if (epi == EPI_LOGITS) {
        // per-row argmax over the CTA's 64 vocabulary columns: lane -> columns 2 lane, 2 lane + 1; first maximum wins (torch argmax)
        for (int r = warp; r < gr.rows; r += 8) {
            const float2 c = *reinterpret_cast<const float2*>(cst + r * CST_LD + 2 * lane);
            const int n = n0 + 2 * lane;
            float best = -INFINITY;
            int bi = 0x7fffffff;
            if (n < a.V) { const float v = c.x + bias[n]; if (v > best) { best = v; bi = n; } }
            if (n + 1 < a.V) { const float v = c.y + bias[n + 1]; if (v > best) { best = v; bi = n + 1; } }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) {
                a.part_val[(size_t)(gr.row0 + r) * CS + gr.crank] = best;
                a.part_idx[(size_t)(gr.row0 + r) * CS + gr.crank] = bi;
            }
        }
    } else {
        const int ppr = nrows >> 1;                           // column pairs per row
        const int total = gr.rows * ppr;
#pragma unroll 1
        for (int i0 = tid; i0 < total; i0 += 4 * NTHR) {      // 4 items per thread and iteration: all loads first
            float2 c[4], xr[4], bb[4];
            int n[4];
            size_t grow[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = i0 + k * NTHR;
                const int r = i < total ? i / ppr : 0, pj = i < total ? i - r * ppr : 0;
                c[k] = *reinterpret_cast<const float2*>(cst + r * CST_LD + 2 * pj);
                n[k] = n0 + 2 * pj;
                grow[k] = (size_t)(gr.row0 + r);
                bb[k] = bias ? *reinterpret_cast<const float2*>(bias + n[k]) : make_float2(0.f, 0.f);
                xr[k] = make_float2(0.f, 0.f);
                if (epi == EPI_RES) xr[k] = __ldcg(reinterpret_cast<const float2*>(a.x + grow[k] * 256 + n[k]));
                else if (epi == EPI_GLU) xr[k].x = __ldcg(a.x + grow[k] * 256 + (n[k] >> 1));
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (i0 + k * NTHR >= total) break;
                const float v0 = c[k].x + bb[k].x, v1 = c[k].y + bb[k].y;
                const int oc = n[k] >> 1;                     // GLU / GeGLU: pair (value n, gate n+1) -> output column n/2
                if (epi == EPI_STORE16) *reinterpret_cast<uint32_t*>(out16 + grow[k] * ld16 + n[k]) = pack_bf16x2(v0, v1);
                else if (epi == EPI_RES) st2(a.s + grow[k] * 256 + n[k], v0 + xr[k].x, v1 + xr[k].y);
                else if (epi == EPI_GLU) a.s[grow[k] * 256 + oc] = v0 * sigmoidf_(v1) + xr[k].x;
                else out16[grow[k] * ld16 + oc] = __float2bfloat16_rn(v0 * gelu_erf(v1));
            }
        }
    }
    // the ring goes back to the TMA (async proxy) in the next attention phase
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------ attention phase
// K (or V) element (row r, 16-byte chunk ch) of head h inside a stage half.  Full stage: one 8 x 16 x 64 box -> [head][16][128 B];
// tail stage: 4-row boxes -> [row/4][head][4][128 B].  The 128-byte swizzle XORs the chunk with address bits 7..9.
TX_DEVINL uint32_t kv_off(bool tail, int h, int r, int ch) {
    if (!tail) return h * HT + r * 128 + ((ch ^ (r & 7)) << 4);
    return (r >> 2) * 4096 + h * 512 + (r & 3) * 128 + ((ch ^ (((h & 1) << 2) | (r & 3))) << 4);
}

// One attention phase of layer l (self: causal over the cache + this step's key, appends k / v; else: encoder memory).
// `it` = ring position (same sequence in the producer and in every consumer warp); returns the new position.
// Activation buffer 0 stages this step's q | k | v rows (3 x 1 KB per sequence) of the CTA's sequences.
__device__ __noinline__ int attn_phase(const Params& p, int self, int l, int t, int it) {
    const MegaArgs& a = p.a;
    const Sm sm = sm_layout();
    const Grp gr = grp_of(a);
    uint8_t* ring = sm.ring;
    uint64_t* full = sm.full;
    uint64_t* empty = sm.empty;
    UnitMeta* meta = sm.meta;
    const int row0 = gr.row0, rows = gr.rows, crank = gr.crank;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bf16* qkv = (const bf16*)a.qkv;
    const int nu = crank < rows ? (rows - crank + CS - 1) / CS : 0;
    if (tid < nu) {
        const int b = row0 + crank + CS * tid;
        UnitMeta m;
        if (self) { m.key0 = 0; m.nk = t; m.z = (l * a.B + b) * 8; }
        else { m.key0 = ldcg_i32(a.enc_off + b); m.nk = ldcg_i32(a.enc_off + b + 1) - m.key0; m.z = l * 8; }
        m.pad = 0;
        meta[tid] = m;
    }
    const uint32_t hb = smem_u32(sm.abuf);
    {   // q (| k | v) rows of the CTA's sequences -> shared memory
        const int per = self ? 192 : 64;          // 16-byte chunks per sequence
        for (int i = tid; i < nu * per; i += NTHR) {
            const int u = i / per, ch = i - u * per;
            cp_async16(hb + u * 3072 + ch * 16, qkv + (size_t)(row0 + crank + CS * u) * 1536 + ch * 8);
        }
        cp_async_commit();
    }
    __syncthreads();                               // meta visible
    // ---- producer state (lane 0 of warp 0): next chunk to issue = chunk pc of unit pu, global chunk counter pi
    int pu = 0, pc = 0, pi = it;
    const CUtensorMap* tm = self ? &p.tm_self : &p.tm_cross;
    const CUtensorMap* tm4 = self ? &p.tm_self4 : &p.tm_cross4;
    auto produce = [&](int limit) {
        while (pi < limit) {
            while (pu < nu && pc * CH >= meta[pu].nk) { ++pu; pc = 0; }
            if (pu >= nu) break;
            const UnitMeta m = meta[pu];
            const int s = pi % NS, ph = (pi / NS) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            const uint32_t st = smem_u32(s < 3 ? ring + s * STAGE : sm.abuf + ABUF);
            const int left = m.nk - pc * CH, k = m.key0 + pc * CH;
            if (left >= CH) {
                mbar_expect_tx(&full[s], STAGE);
                tma_load_3d(tm, &full[s], st, 0, k, m.z);
                tma_load_3d(tm, &full[s], st + HALF, 64, k, m.z);
            } else {               // tail: 4-row boxes, at most 3 rows fetched beyond the sequence
                const int n4 = (left + 3) >> 2;
                mbar_expect_tx(&full[s], n4 * 2 * 4096);
                for (int j = 0; j < n4; ++j) {
                    tma_load_3d(tm4, &full[s], st + j * 4096, 0, k + 4 * j, m.z);
                    tma_load_3d(tm4, &full[s], st + HALF + j * 4096, 64, k + 4 * j, m.z);
                }
            }
            ++pc; ++pi;
        }
    };
    if (tid == 0) {
        asm volatile("fence.proxy.async.global;" ::: "memory");      // cache rows appended by generic-proxy stores of earlier steps
        produce(it + NS);
    }
    cp_async_wait<0>();
    __syncthreads();                               // headers visible
    const int h = warp;
    if (self) {
        // append this step's k / v rows of head h to the cache: lanes 0..7 the K row, 8..15 the V row (16 B each)
        if (lane < 16) {
            for (int u = 0; u < nu; ++u) {
                const size_t b = (size_t)(row0 + crank + CS * u);
                const uint4 val = ld_shared128(hb + u * 3072 + 1024 + (lane >> 3) * 1024 + h * 128 + (lane & 7) * 16);
                bf16* row = (bf16*)a.kv_self + ((((size_t)l * a.B + b) * 8 + h) * a.tcap + t) * 128 + (lane >> 3) * 64 + (lane & 7) * 8;
                *reinterpret_cast<uint4*>(row) = val;
            }
            asm volatile("fence.proxy.async.global;" ::: "memory");     // read back through TMA from the next step on
        }
        __syncwarp();
    }
    const int g = lane >> 2, tq = lane & 3;
    const bool row0_lane = g == 0;                       // the single query lives in row 0 of the 16-row A operand
    const int lm_r = lane & 7, lm_m = lane >> 3;
#pragma unroll 1
    for (int u = 0; u < nu; ++u) {
        const size_t b = (size_t)(row0 + crank + CS * u);
        const uint32_t hu = hb + u * 3072 + h * 128;
        uint32_t qa[8];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            // fragment words: dims 16s+2tq,+1 and 16s+8+2tq,+1; q * 0.125 is exact in bf16
            const float2 f0 = unpack_bf16x2(ld_shared32(hu + (8 * s + tq) * 4)), f1 = unpack_bf16x2(ld_shared32(hu + (8 * s + 4 + tq) * 4));
            qa[2 * s] = row0_lane ? pack_bf16x2(f0.x * SCALE, f0.y * SCALE) : 0u;
            qa[2 * s + 1] = row0_lane ? pack_bf16x2(f1.x * SCALE, f1.y * SCALE) : 0u;
        }
        const int nk = meta[u].nk;
        const int nchunk = (nk + CH - 1) / CH;
        float m = -INFINITY, lsum = 0.f;
        float o[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
#pragma unroll 1
        for (int c = 0; c < nchunk; ++c, ++it) {
            if (warp == 0) {
                if (lane == 0) produce(it + NS);
                __syncwarp();
            }
            const int s = it % NS, ph = (it / NS) & 1;
            const bool tail = nk - c * CH < CH;
            mbar_wait(&full[s], ph);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const uint32_t kt = smem_u32(s < 3 ? ring + s * STAGE : sm.abuf + ABUF);
            const uint32_t vt = kt + HALF;
            float sc[2][4];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
                const int kr = 8 * j + lm_r;
#pragma unroll
                for (int s2 = 0; s2 < 2; ++s2) {
                    uint32_t b0, b1, b2, b3;
                    ldsm_x4(kt + kv_off(tail, h, kr, 4 * s2 + lm_m), b0, b1, b2, b3);
                    mma_bf16(sc[j], qa[4 * s2], 0u, qa[4 * s2 + 1], 0u, b0, b1);
                    mma_bf16(sc[j], qa[4 * s2 + 2], 0u, qa[4 * s2 + 3], 0u, b2, b3);
                }
            }
            const int kbase = c * CH + 2 * tq;
            float pr[2][2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                pr[j][0] = (kbase + 8 * j < nk) ? sc[j][0] : -INFINITY;
                pr[j][1] = (kbase + 8 * j + 1 < nk) ? sc[j][1] : -INFINITY;
            }
            float cm = fmaxf(fmaxf(pr[0][0], pr[0][1]), fmaxf(pr[1][0], pr[1][1]));
            cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 1));
            cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 2));
            const float mn = fmaxf(m, cm);                 // finite on row 0: every stage holds at least one valid key
            const float corr = __expf(m - mn);
#pragma unroll
            for (int j = 0; j < 2; ++j) { pr[j][0] = __expf(pr[j][0] - mn); pr[j][1] = __expf(pr[j][1] - mn); }
            lsum = lsum * corr + (pr[0][0] + pr[0][1]) + (pr[1][0] + pr[1][1]);
            m = mn;
            const uint32_t pa0 = row0_lane ? pack_bf16x2(pr[0][0], pr[0][1]) : 0u;
            const uint32_t pa2 = row0_lane ? pack_bf16x2(pr[1][0], pr[1][1]) : 0u;
            const int vr = (lane & 7) + 8 * ((lane >> 3) & 1);
            // rows past the end of the sequence carry p = 0, but 0 * NaN = NaN: clear their halves of the V fragments
            uint32_t vm_lo = 0xffffffffu, vm_hi = 0xffffffffu;
            if (tail) {
                const int k0 = c * CH + 2 * tq;
                vm_lo = (k0 < nk ? 0x0000ffffu : 0u) | (k0 + 1 < nk ? 0xffff0000u : 0u);
                vm_hi = (k0 + 8 < nk ? 0x0000ffffu : 0u) | (k0 + 9 < nk ? 0xffff0000u : 0u);
            }
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t b0, b1, b2, b3;
                ldsm_x4_t(vt + kv_off(tail, h, vr, 2 * np + (lane >> 4)), b0, b1, b2, b3);
                b0 &= vm_lo; b2 &= vm_lo; b1 &= vm_hi; b3 &= vm_hi;
                o[2 * np][0] *= corr; o[2 * np][1] *= corr;
                o[2 * np + 1][0] *= corr; o[2 * np + 1][1] *= corr;
                mma_bf16(o[2 * np], pa0, 0u, pa2, 0u, b0, b1);
                mma_bf16(o[2 * np + 1], pa0, 0u, pa2, 0u, b2, b3);
            }
            // order this warp's generic-proxy reads (ldmatrix) before the async-proxy refill of the stage
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        if (self) {
            // this step's own key / value (from the staged header): the lane's 16 of the 64 dims, folded over the 4 lanes of row 0
            float d = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const int word = 8 * (w >> 1) + 4 * (w & 1) + tq;
                const float2 qf = unpack_bf16x2(ld_shared32(hu + word * 4)), kf = unpack_bf16x2(ld_shared32(hu + 1024 + word * 4));
                d = fmaf(qf.x * SCALE, kf.x, d);
                d = fmaf(qf.y * SCALE, kf.y, d);
            }
            d += __shfl_xor_sync(0xffffffffu, d, 1);
            d += __shfl_xor_sync(0xffffffffu, d, 2);
            const float mn = fmaxf(m, d);
            const float corr = __expf(m - mn), pn = __expf(d - mn);
            lsum = lsum * corr;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {             // dims 8nt+2tq, +1
                const float2 vf = unpack_bf16x2(ld_shared32(hu + 2048 + (4 * nt + tq) * 4));
                o[nt][0] = fmaf(pn, vf.x, o[nt][0] * corr);
                o[nt][1] = fmaf(pn, vf.y, o[nt][1] * corr);
            }
            lsum += (tq == 0) ? pn : 0.f;
        }
        lsum += __shfl_xor_sync(0xffffffffu, lsum, 1);
        lsum += __shfl_xor_sync(0xffffffffu, lsum, 2);
        if (row0_lane) {
            const float inv = 1.0f / lsum;
            uint32_t* op = reinterpret_cast<uint32_t*>((bf16*)a.o + b * 512 + h * 64);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) op[4 * nt + tq] = pack_bf16x2(o[nt][0] * inv, o[nt][1] * inv);
        }
    }
    return it;
}

// ------------------------------------------------------------------------------------------------ the kernel
#define CSYNC()         \
    do {                \
        cluster_sync(); \
        TICK(10);       \
    } while (0)

__global__ void __launch_bounds__(NTHR, 1) decode_mega_kernel(const __grid_constant__ Params p) {
    const MegaArgs& a = p.a;
    const Sm sm = sm_layout();
    const Grp gr = grp_of(a);
    unsigned long long* tacc = sm.tacc;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int crank = gr.crank, grp = gr.grp, row0 = gr.row0, rows = gr.rows;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (a.dbg_time) { for (int i = 0; i < 18; ++i) tacc[i] = 0; tacc[16] = gtime(); tacc[17] = tacc[16]; }
    }
    __syncthreads();

    int t = ldcg_i32(a.step + grp);
    int it = 0;
    const int ngemm = 6 * a.L + 1;
    WDesc cur = w_desc(a, 0, crank);
    load_w(sm.wbuf, sm.abuf + ABUF, cur, tid);
    const int t_end = (t + a.nsteps < a.tcap) ? t + a.nsteps : a.tcap;
#pragma unroll 1
    for (; t < t_end; ++t) {
        // GEMM gi of the step: layer gi / 6, kind gi % 6 = 0 QKV, 1 self out-proj, 2 cross Q, 3 cross out-proj, 4 FF1, 5 FF2; last: logits
#pragma unroll 1
        for (int gi = 0; gi < ngemm; ++gi) {
            const int l = gi / 6, kind = gi + 1 == ngemm ? 6 : gi - 6 * l;
            const MegaLayerW& w = a.layer[l < a.L ? l : 0];
            const bool have_next = gi + 1 < ngemm || t + 1 < t_end;
            const WDesc nxt = w_desc(a, gi + 1 < ngemm ? gi + 1 : 0, crank);
            if (!(kind & 1)) {            // LayerNorm-fed GEMMs rebuild their A operand from the residual stream
                prologue_rows(p, gi == 0 ? PRO_EMBED : (kind == 6 ? PRO_FIN : PRO_LN2), t);
                TICK(12);
            }
            const bf16* csrc = nullptr;
            int cld = 0, epi = EPI_STORE16, ld16 = 1536;
            const float* bias = nullptr;
            bf16* out16 = (bf16*)a.qkv;
            if (kind == 1 || kind == 3) { csrc = (const bf16*)a.o; cld = 512; epi = EPI_GLU; bias = kind == 1 ? w.bo_s : w.bo_c; }
            else if (kind == 4) { epi = EPI_GEGLU; bias = w.b1; out16 = (bf16*)a.hid; ld16 = 1024; }
            else if (kind == 5) { csrc = (const bf16*)a.hid; cld = 1024; epi = EPI_RES; bias = w.b2; }
            else if (kind == 6) { epi = EPI_LOGITS; bias = a.b_logits; }
            gemm_mma(p, cur, nxt, have_next, csrc, cld);
            epilogue(p, epi, cur.n0, cur.nrows, bias, out16, ld16);
            cur = nxt;
            TICK(kind);
            CSYNC();
            if (kind == 0 || kind == 2) {     // attention (self: causal over the cache + this step's key, appends k / v)
                it = attn_phase(p, kind == 0, l, t, it);
                TICK(kind == 0 ? 7 : 8);
                CSYNC();
            }
        }
        // ---- argmax over the 16 partials; token bookkeeping (model/decoder.py:103-116)
        if (warp < MAXU) {
            const int r = crank + CS * warp;
            if (r < rows) {
                float v = -INFINITY;
                int bi = 0x7fffffff;
                if (lane < CS) {
                    v = __ldcg(a.part_val + (size_t)(row0 + r) * CS + lane);
                    bi = __ldcg(a.part_idx + (size_t)(row0 + r) * CS + lane);
                }
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > v || (ov == v && oi < bi)) { v = ov; bi = oi; }
                }
                if (lane == 0) {
                    if (bi == 0x7fffffff) bi = 0;
                    a.out_ids[(size_t)(row0 + r) * a.tcap + t] = bi;
                    a.cur_tok[row0 + r] = bi;
                    if (a.eos >= 0 && bi == a.eos) a.seen[row0 + r] = 1;
                }
            }
        }
        TICK(9);
        CSYNC();
        if (crank == 0 && warp == 7) {       // has every sequence of the group produced an EOS?
            int all = 1;
            for (int r = lane; r < rows; r += 32) all &= (ldcg_i32(a.seen + row0 + r) != 0);
            all = __all_sync(0xffffffffu, all);
            if (lane == 0 && all && ldcg_i32(a.done_step + grp) == 0) a.done_step[grp] = t + 1;
        }
    }
    if (crank == 0 && tid == 0) a.step[grp] = t;
    cp_async_wait<0>();
    if (a.dbg_time && tid == 0) {
        tacc[11] = gtime() - tacc[17];
        for (int i = 0; i < 16; ++i) atomicAdd(a.dbg_time + i, tacc[i]);
    }
}

int g_clusters = -1;

}  // namespace

int decode_mega_active_clusters() {
    if (g_clusters < 0) {
        g_clusters = 0;
        if (cudaFuncSetAttribute(decode_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) == cudaSuccess &&
            cudaFuncSetAttribute(decode_mega_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(CS * 8); cfg.blockDim = dim3(NTHR); cfg.dynamicSmemBytes = SMEM_BYTES;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, decode_mega_kernel, &cfg) == cudaSuccess) g_clusters = n;
        }
        cudaGetLastError();
    }
    return g_clusters;
}

// Groups of a batch: as many clusters as the device holds at once (fewer for small batches: >= 16 rows per group), more
// only when a group would exceed 80 rows (those run as a second wave).
int decode_mega_groups(int B) {
    const int nact = std::max(1, decode_mega_active_clusters());
    int G = std::min(nact, (B + 15) / 16);
    G = std::max(G, 1);
    if ((B + G - 1) / G > RG) G = (B + RG - 1) / RG;
    return G;
}

bool decode_mega_supported(int B, int L, int V, int* why) {
    int w = 0;
    if (L < 1 || L > MEGA_MAX_LAYERS) w = 1;
    else if (V < 1 || V > MEGA_CLUSTER * 64) w = 2;
    else if (B < 1) w = 3;
    else if (decode_mega_active_clusters() < 1) w = 4;
    if (why) *why = w;
    return w == 0;
}

cudaError_t launch_decode_mega(const MegaArgs& a, cudaStream_t st) {
    if (a.G < 1 || (a.B + a.G - 1) / a.G > RG) return cudaErrorInvalidValue;
    Params p;
    p.a = a;
    cudaError_t e;
    const long zs = (long)a.L * a.B * 8;
    if ((e = tma_map_3d_bf16(a.kv_self, 128, a.tcap, zs, 256, (long)a.tcap * 256, 64, CH, 8, &p.tm_self)) != cudaSuccess) return e;
    if ((e = tma_map_3d_bf16(a.kv_self, 128, a.tcap, zs, 256, (long)a.tcap * 256, 64, 4, 8, &p.tm_self4)) != cudaSuccess) return e;
    if ((e = tma_map_3d_bf16(a.kv_cross, 128, a.ntok, (long)a.L * 8, 256, a.ntok * 256, 64, CH, 8, &p.tm_cross)) != cudaSuccess) return e;
    if ((e = tma_map_3d_bf16(a.kv_cross, 128, a.ntok, (long)a.L * 8, 256, a.ntok * 256, 64, 4, 8, &p.tm_cross4)) != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(a.G * CS); cfg.blockDim = dim3(NTHR); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, decode_mega_kernel, p);
}
