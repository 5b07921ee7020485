// tcgen05 GEMM for sm_100a:  C[M,N] = epi(A[M,K] . W[N,K]^T), bf16 operands (both K-major), fp32 accumulation in TMEM.
//
//   TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ring (STAGES x {A 128x64, W BNx64})
//   -> tcgen05.mma.cta_group::1.kind::f16 issued by one thread, accumulator 128 lanes x BN columns in TMEM
//   -> tcgen05.ld (32x32b) by four or eight epilogue warps, one accumulator row per thread, fused epilogue
//      (bias / GLU+residual / GeGLU / bias+residual / GroupNorm partial sums), vectorised row-segment stores.
//
// Warp roles (64 + 128 * EW threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2.. = epilogue
// (warp w may only touch TMEM lanes 32*(w%4)..+31, so four consecutive warps cover the 128 rows).  EW = 2: a second
// epilogue warpgroup (warps 6..9) takes every other column chunk of the tile (halves of the single chunk when BN = 32) --
// the epilogue, not the MMA loop, bounds the K <= 1024 GEMMs of this model.
//
// SPLIT = 3 is the "bf16x3" mode used for the ill-conditioned ResNet backbone: both operands are
// given as hi + lo bf16 pairs and every k-slice issues hi.hi + hi.lo + lo.hi into the same accumulator, which
// carries ~16 mantissa bits per operand through the tensor cores.
#include <cuda.h>

#include <map>
#include <mutex>
#include <tuple>
#include <type_traits>

#include "common.cuh"
#include "gn_block.cuh"
#include "tc_gemm.h"

int g_tc_persistent = 1;          // texocr_set_option("gemm_persistent"): persistent double-buffered kernel for GEMMs of >= 296 tiles
int g_tc_min_ctas = 120;          // tile width rule: narrow the N tile (128 -> 64 -> 32) while the grid would have fewer CTAs than this
int g_tc_persistent_stages = 0;   // 0 = as many ring stages as fit in 200 KB; n > 0 caps them (leaves shared memory to co-resident kernels)
int g_tc_bn256 = 1;               // texocr_set_option("gemm_bn256"): 128 x 256 tiles for the wide split-operand convolutions
int g_tc_epi_warps = 8;           // texocr_set_option("gemm_epi_warps"): 4 or 8 epilogue warps per CTA

namespace {

constexpr int BM = 128, BK = 64, UMMA_K = 16;

struct TcParams {
    void* C; int M, N, K, ldc;
    const float* bias; const float* res; int ldres;
    int late_trigger;
    // implicit-GEMM convolution (TMA im2col loads of A): cpk = 64-channel chunks per filter tap (0 = plain GEMM)
    int cv_cpk, cv_ksz, cv_stride, cv_lower, cv_Wo, cv_Ho;
    int stages;       // persistent kernel: ring stages actually used (<= SmemP::STAGES)
    int a_block_k;    // block-diagonal GEMM (one-tile-per-CTA kernel, BN = 64): n-tile j reads A columns from j * a_block_k
    unsigned long long* dbg;      // debug residency sums (GemmArgs::dbg)
    // GroupNorm partial sums of the fp32 output (EPI_STORE, M rows = same-size images of gn_rpi rows each, gn_rpi % 32 == 0):
    // gn_part[(row >> 5) + image][32][2], see gn_block.cuh.  null = off
    float* gn_part; int gn_cpg, gn_rpi;
    const int* gn_img_off; int gn_nimg, gn_level;      // gn_rpi == 0: images of different sizes, each a multiple of 32 rows at this level
    // ragged implicit-GEMM convolution (tc_conv_gather_kernel): the split-bf16 NHWC activation and the batch geometry
    const bf16* g_hi; const bf16* g_lo; const int* g_img_off; const int* g_img_hw; int g_nimg, g_lin, g_lout, g_pad;
};
TX_DEVINL unsigned long long gtime_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// ------------------------------------------------------------------------------------------------ PTX wrappers
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
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
TX_DEVINL void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// im2col-mode TMA load: 128 consecutive output pixels (w fastest, then h, then n, stepping by the map's traversal stride and
// wrapping inside its bounding box) x 64 channels of the input pixel at (base + filter offset); out-of-image taps are zero
TX_DEVINL void tma_load_im2col(const CUtensorMap* map, uint64_t* bar, void* dst, int c, int w, int h, int n, int off_w, int off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"((unsigned short)off_w), "h"((unsigned short)off_h)
        : "memory");
}
TX_DEVINL void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
TX_DEVINL void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
TX_DEVINL void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
TX_DEVINL void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
TX_DEVINL void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
TX_DEVINL void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

TX_DEVINL void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major operand tile, 128-byte swizzle: rows of 64 bf16 (128 B), 8-row groups 1024 B apart (SBO), LBO unused (=1).
// Bit layout: cute::UMMA::SmemDescriptor (start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout [61,64)).
TX_DEVINL uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;       // SWIZZLE_128B
    return d;
}
// cute::UMMA::InstrDescriptor: c_format f32 (1<<4), a/b format bf16 (1<<7, 1<<10), K-major both, N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t make_idesc(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24); }

// Epilogue shared by the GEMM kernels.  Thread = accumulator row (TMEM lane), 32 columns per tcgen05.ld.  Global traffic
// goes through a per-warp staging tile in shared memory (the pipeline stages are idle once the accumulator is complete),
// so that residual loads and output stores are whole 64..128-byte row segments per quarter-warp instead of one row per
// lane: the epilogue, not the MMA loop, bounds these K <= 1024 GEMMs.
constexpr int STG_STRIDE = 144;                 // bytes per staged row (128 + 16: conflict-free 16-byte accesses)
constexpr int STG_WARP = 32 * STG_STRIDE;       // staging bytes per epilogue warp

// copy a [32 rows x RB bytes] tile between the warp's staging area and global memory, 16 bytes per lane
template <int RB, bool TO_GLOBAL>
TX_DEVINL void stage_copy(uint8_t* stg, uint8_t* gptr, size_t grow_bytes, int lane, int rows_ok, int bytes_ok) {
    constexpr int CPR = RB / 16, RPI = 32 / CPR;
    const int rr = lane / CPR, ch = lane % CPR;
#pragma unroll
    for (int it = 0; it < CPR; ++it) {
        const int row = it * RPI + rr;
        if (row < rows_ok && ch * 16 < bytes_ok) {
            uint4* sp = reinterpret_cast<uint4*>(stg + row * STG_STRIDE + ch * 16);
            uint4* gp = reinterpret_cast<uint4*>(gptr + (size_t)row * grow_bytes + ch * 16);
            if (TO_GLOBAL) *gp = *sp; else *sp = __ldcg(gp);
        }
    }
}

// [32 rows x 128 bytes] of fp32 outputs, staging -> global like stage_copy<128, true>, and on the way the lane's column sums /
// sums of squares over its 8 rows (gn_block.cuh)
TX_DEVINL void stage_copy_gn(uint8_t* stg, uint8_t* gptr, size_t grow_bytes, int lane, int rows_ok, float* s, float* q) {
    const int rr = lane >> 3, ch = lane & 7;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int row = it * 4 + rr;
        if (row < rows_ok) {
            const float4 f = *reinterpret_cast<const float4*>(stg + row * STG_STRIDE + ch * 16);
            *reinterpret_cast<float4*>(gptr + (size_t)row * grow_bytes + ch * 16) = f;
            gn_block_acc(s, q, f);
        }
    }
}

// the two halves of stage_copy<RB, false>: global -> registers (issued early, the latency overlaps other work), registers -> staging
template <int RB>
TX_DEVINL void stage_fetch(const uint8_t* gptr, size_t grow_bytes, int lane, int rows_ok, int bytes_ok, uint4* rq) {
    constexpr int CPR = RB / 16, RPI = 32 / CPR;
    const int rr = lane / CPR, ch = lane % CPR;
#pragma unroll
    for (int it = 0; it < CPR; ++it) {
        const int row = it * RPI + rr;
        rq[it] = make_uint4(0, 0, 0, 0);
        if (row < rows_ok && ch * 16 < bytes_ok) rq[it] = __ldcg(reinterpret_cast<const uint4*>(gptr + (size_t)row * grow_bytes + ch * 16));
    }
}
template <int RB>
TX_DEVINL void stage_put(uint8_t* stg, int lane, const uint4* rq) {
    constexpr int CPR = RB / 16, RPI = 32 / CPR;
    const int rr = lane / CPR, ch = lane % CPR;
#pragma unroll
    for (int it = 0; it < CPR; ++it) *reinterpret_cast<uint4*>(stg + (it * RPI + rr) * STG_STRIDE + ch * 16) = rq[it];
}

template <int BN, int EPI, typename TC, int EW, bool AHEAD = false>
TX_DEVINL void epilogue_tile(uint64_t* tmem_full, uint32_t tmem_base, int warp, int lane, int m0, int n0, const TcParams& p,
                             uint8_t* smem_idle, uint32_t parity = 0, unsigned long long dbg_t0 = 0ull, int* gn_hint = nullptr) {
    constexpr int CW = (EW == 2 && BN == 32) ? 16 : 32;         // columns per tcgen05.ld chunk
    const int q = warp & 3;                                     // TMEM lane quarter of this warp
    const int ew = warp - 2, eg = ew >> 2;                      // epilogue warp / warpgroup index
    uint8_t* stg = smem_idle + ew * STG_WARP;
    uint8_t* my = stg + lane * STG_STRIDE;                      // this thread's staged row
    const int mrow0 = m0 + q * 32;
    const int rows_ok = min(32, p.M - mrow0);                   // <= 0: nothing of this warp's rows is inside the matrix
    constexpr bool PAIRS = (EPI == EPI_GLU_RES || EPI == EPI_GEGLU);
    constexpr int NOUT = PAIRS ? CW / 2 : CW;
    using TO = typename std::conditional<EPI == EPI_STORE, TC, typename std::conditional<EPI == EPI_GEGLU, bf16, float>::type>::type;
    constexpr int RB = NOUT * (int)sizeof(TO);
    const int cfirst = eg * CW;                                 // this warp's first chunk
    // Narrow tiles (the latency-bound decode GEMMs): fetch the thread's residual row while the MMAs are still running --
    // it does not depend on the accumulator, and its L2 round trip would otherwise sit behind the tmem_full wait.
    constexpr bool PRE_RES = BN == 32 && (EPI == EPI_BIAS_RES || EPI == EPI_GLU_RES);
    float rpre[PRE_RES ? NOUT : 1];
    if (PRE_RES) {
        const int nf = n0 + cfirst;
        const int ncol = PAIRS ? (nf >> 1) : nf, nvalid_out = PAIRS ? (max(0, min(CW, p.N - nf)) >> 1) : max(0, min(CW, p.N - nf));
        const float* rp = p.res + (size_t)(mrow0 + lane) * p.ldres + ncol;
#pragma unroll
        for (int i = 0; i < (PRE_RES ? NOUT : 0); i += 4) {
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lane < rows_ok && i < nvalid_out) t = __ldcg(reinterpret_cast<const float4*>(rp + i));
            rpre[i] = t.x; rpre[i + 1] = t.y; rpre[i + 2] = t.z; rpre[i + 3] = t.w;
        }
    }
    // bias of the warp's first chunk: a weight, fetched before the wait as well (narrow tiles only: 32 more live registers)
    constexpr bool PRE_BIAS = BN == 32 || EW == 1;
    float bpre[PRE_BIAS ? CW : 4];
#pragma unroll
    for (int i = 0; i < (PRE_BIAS ? CW : 0); i += 4) {
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias && n0 + cfirst + i < p.N) b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + cfirst + i));
        bpre[i] = b.x; bpre[i + 1] = b.y; bpre[i + 2] = b.z; bpre[i + 3] = b.w;
    }
    // Wide tiles with a residual: the residual tile of a chunk is fetched (coalesced, into registers) one chunk ahead -- the first one
    // before the accumulator wait, the next one before the current chunk's math and stores -- so that its L2 / HBM round trip is
    // never exposed (ncu: the 128 x 128 GLU + residual epilogue of the ViT blocks took 9.6 us per tile against 0.5 us of MMAs)
    constexpr bool RES_AHEAD = AHEAD && !PRE_RES && (EPI == EPI_BIAS_RES || EPI == EPI_GLU_RES);      // persistent kernel (register budget of one CTA per SM)
    uint4 rq[RES_AHEAD ? NOUT * 4 / 16 : 1];
    auto res_fetch = [&](int c0) {
        const int n = n0 + c0;
        if (rows_ok <= 0 || n >= p.N) return;
        const int nv = min(CW, p.N - n);
        stage_fetch<NOUT * 4>(reinterpret_cast<const uint8_t*>(p.res + (size_t)mrow0 * p.ldres + (PAIRS ? (n >> 1) : n)), (size_t)p.ldres * 4, lane,
                              rows_ok, (PAIRS ? (nv >> 1) : nv) * 4, rq);
    };
    if (RES_AHEAD) res_fetch(cfirst);
    // GroupNorm partials: the slot of this warp's row block (image index by division, or by a search of the offset table for ragged
    // batches), looked up while the MMAs are still running
    int gn_slot = 0;
    if constexpr (EPI == EPI_STORE && std::is_same<TC, float>::value && CW == 32) {
        if (p.gn_part && rows_ok > 0) {
            int bimg;
            if (p.gn_rpi > 0) bimg = mrow0 / p.gn_rpi;
            else {
                bimg = find_image_from(p.gn_img_off, p.gn_nimg, p.gn_level, mrow0, gn_hint ? *gn_hint : -1);
                if (gn_hint) *gn_hint = bimg;
            }
            gn_slot = (mrow0 >> 5) + bimg;
        }
    }
    mbar_wait(tmem_full, parity);
    if (p.late_trigger == 2) pdl_launch_dependents();           // one-tile kernel: what is left of this CTA is about as long as the dependent's launch + prologue
    if (dbg_t0) atomicAdd(p.dbg + 6, gtime_ns() - dbg_t0);      // debug: epilogue warp entry -> accumulator complete
    tcgen05_fence_after();
    float am_best = -INFINITY;                                  // EPI_ARGMAX: running maximum of this thread's row over the tile
    int am_idx = 0x7fffffff;
#pragma unroll 1
    for (int c0 = cfirst; c0 < BN; c0 += EW * CW) {
        uint32_t raw[CW];
        if constexpr (CW == 32) tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, raw);
        else tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, raw);
        if (dbg_t0 && c0 == cfirst) atomicAdd(p.dbg + 7, gtime_ns() - dbg_t0);      // debug: -> first accumulator chunk in registers
        const int n = n0 + c0;
        if (rows_ok <= 0 || n >= p.N) continue;                 // warp-uniform
        const int nvalid = min(CW, p.N - n);                    // multiple of 8
        const int ncol = PAIRS ? (n >> 1) : n, nvalid_out = PAIRS ? (nvalid >> 1) : nvalid;
        float v[CW];
#pragma unroll
        for (int i = 0; i < CW; ++i) v[i] = __uint_as_float(raw[i]);
        if (p.bias) {
            if (PRE_BIAS && c0 == cfirst) {
#pragma unroll
                for (int i = 0; i < CW; ++i) v[i] += bpre[PRE_BIAS ? i : 0];
            } else {
#pragma unroll
                for (int i = 0; i < CW; i += 4) {
                    if (i < nvalid) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n + i));
                        v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
                    }
                }
            }
        }
        if constexpr (EPI == EPI_ARGMAX) {                      // ascending scan, strict >: the first maximum wins (torch argmax)
#pragma unroll
            for (int i = 0; i < CW; ++i)
                if (i < nvalid && v[i] > am_best) { am_best = v[i]; am_idx = n + i; }
            continue;
        }
        float o[NOUT];
        if (PRE_RES) {
#pragma unroll
            for (int i = 0; i < NOUT; ++i)
                o[i] = (EPI == EPI_BIAS_RES) ? v[i] + rpre[PRE_RES ? i : 0] : v[2 * i] * sigmoid_fast(v[2 * i + 1]) + rpre[PRE_RES ? i : 0];
        } else if (EPI == EPI_BIAS_RES || EPI == EPI_GLU_RES) {
            // residual tile (fetched a chunk ahead, or now) -> staging, then every thread picks up its own row
            if (RES_AHEAD) stage_put<NOUT * 4>(stg, lane, rq);
            else stage_copy<NOUT * 4, false>(stg, reinterpret_cast<uint8_t*>(const_cast<float*>(p.res) + (size_t)mrow0 * p.ldres + ncol),
                                             (size_t)p.ldres * 4, lane, rows_ok, nvalid_out * 4);
            __syncwarp();
            float r[NOUT];
#pragma unroll
            for (int i = 0; i < NOUT; i += 4) {
                const float4 t = *reinterpret_cast<const float4*>(my + i * 4);
                r[i] = t.x; r[i + 1] = t.y; r[i + 2] = t.z; r[i + 3] = t.w;
            }
            __syncwarp();
            if (RES_AHEAD && c0 + EW * CW < BN) res_fetch(c0 + EW * CW);
#pragma unroll
            for (int i = 0; i < NOUT; ++i)
                o[i] = (EPI == EPI_BIAS_RES) ? v[i] + r[i] : v[2 * i] * sigmoid_fast(v[2 * i + 1]) + r[i];
        } else if (EPI == EPI_GEGLU) {
#pragma unroll
            for (int i = 0; i < NOUT; ++i) o[i] = v[2 * i] * gelu_erf_fast(v[2 * i + 1]);
        } else {
#pragma unroll
            for (int i = 0; i < NOUT; ++i) o[i] = v[i];
        }
        // outputs -> staging (own row) -> coalesced global stores
        TO* mine = reinterpret_cast<TO*>(my);
#pragma unroll
        for (int i = 0; i < NOUT; i += 4) st4(mine + i, make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]));
        if (dbg_t0 && c0 == cfirst) atomicAdd(p.dbg + 8, gtime_ns() - dbg_t0);      // debug: -> first chunk computed and staged
        __syncwarp();
        uint8_t* gdst = reinterpret_cast<uint8_t*>(reinterpret_cast<TO*>(p.C) + (size_t)mrow0 * p.ldc + ncol);
        bool done = false;
        if constexpr (EPI == EPI_STORE && std::is_same<TC, float>::value && CW == 32) {
            if (p.gn_part && nvalid == 32) {                    // convolution output: GroupNorm partial sums of the tile on its way out
                float gs[4] = {0.f, 0.f, 0.f, 0.f}, gq[4] = {0.f, 0.f, 0.f, 0.f};
                stage_copy_gn(stg, gdst, (size_t)p.ldc * 4, lane, rows_ok, gs, gq);
                gn_block_finish(gs, gq, lane, p.gn_cpg, n, p.gn_part + (size_t)gn_slot * 64);
                done = true;
            }
        }
        if (!done) stage_copy<RB, true>(stg, gdst, (size_t)p.ldc * sizeof(TO), lane, rows_ok, nvalid_out * (int)sizeof(TO));
        __syncwarp();
    }
    if constexpr (EPI == EPI_ARGMAX) {
        if (lane < rows_ok)
            reinterpret_cast<float2*>(p.C)[(size_t)(mrow0 + lane) * p.ldc + n0 / BN] = make_float2(am_best, __int_as_float(am_idx));
    }
}

// NSTG = 0: as many stages as fit (deep prefetch for the latency-bound small-M decode GEMMs); NSTG = 2: two stages, so that
// 2-3 CTAs share an SM and one CTA's epilogue / ramp-up overlaps another's MMA loop (large-M encoder GEMMs).
template <int BN, int SPLIT, int NSTG = 0> struct Smem {
    static constexpr int A_BYTES = BM * BK * 2, W_BYTES = BN * BK * 2;
    static constexpr int NOPS = SPLIT == 3 ? 2 : 1;                 // hi (+ lo) copies per operand
    static constexpr int STAGE = NOPS * (A_BYTES + W_BYTES);
    static constexpr int STAGES = NSTG ? NSTG : ((STAGE * 4 <= 160 * 1024) ? 4 : (STAGE * 3 <= 200 * 1024 ? 3 : 2));
    static constexpr int BARS = 256;
    static constexpr int TOTAL = STAGES * STAGE + 1024 /*align slack*/ + BARS;
};

template <int BN, int EPI, typename TC, int SPLIT, int NSTG, int EW>
__global__ void __launch_bounds__(64 + 128 * EW, EW)         // EW = 2: <= 102 registers, so that two 320-thread CTAs still share an SM
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmW2, const TcParams p) {
    using S = Smem<BN, SPLIT, NSTG>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::STAGES * S::STAGE);
    uint64_t* empty = full + S::STAGES;
    uint64_t* tmem_full = empty + S::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int nkb = p.K / BK;

    if (!p.late_trigger) pdl_launch_dependents();
    unsigned long long dbg_t0 = 0ull, dbg_t1 = 0ull;
    if (p.dbg && threadIdx.x == 0) dbg_t0 = gtime_ns();
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
        for (int s = 0; s < S::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the previous kernel, and so do the
    // weight tiles of the first ring round: W never depends on the predecessor, only the activations (A) do.
    if (warp == 0) {
        if (lane == 0) {
            const int npre = nkb < S::STAGES ? nkb : S::STAGES;
            for (int kb = 0; kb < npre; ++kb) {
                uint8_t* st = smem + kb * S::STAGE;
                mbar_expect_tx(&full[kb], S::STAGE);
                tma_load_2d(&tmW, &full[kb], st + S::NOPS * S::A_BYTES, kb * BK, n0);
                if (SPLIT == 3) tma_load_2d(&tmW2, &full[kb], st + S::NOPS * S::A_BYTES + S::W_BYTES, kb * BK, n0);
            }
            pdl_wait();
            if (p.dbg) dbg_t1 = gtime_ns();
            // convolution: base input pixel of the tile's first output pixel m0 = (n, oh, ow)
            int cv_w = 0, cv_h = 0, cv_n = 0;
            if (p.cv_cpk) {
                const int per = p.cv_Wo * p.cv_Ho;
                cv_n = m0 / per;
                const int rem = m0 - cv_n * per, oh = rem / p.cv_Wo;
                cv_h = oh * p.cv_stride + p.cv_lower;
                cv_w = (rem - oh * p.cv_Wo) * p.cv_stride + p.cv_lower;
            }
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % S::STAGES, ph = (kb / S::STAGES) & 1;
                uint8_t* st = smem + s * S::STAGE;
                if (kb >= npre) {
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], S::STAGE);
                    tma_load_2d(&tmW, &full[s], st + S::NOPS * S::A_BYTES, kb * BK, n0);
                    if (SPLIT == 3) tma_load_2d(&tmW2, &full[s], st + S::NOPS * S::A_BYTES + S::W_BYTES, kb * BK, n0);
                }
                if (p.cv_cpk) {
                    const int tap = kb / p.cv_cpk, cc = (kb - tap * p.cv_cpk) * BK;
                    const int ky = tap / p.cv_ksz, kx = tap - ky * p.cv_ksz;
                    tma_load_im2col(&tmA, &full[s], st, cc, cv_w, cv_h, cv_n, kx, ky);
                    if (SPLIT == 3) tma_load_im2col(&tmA2, &full[s], st + S::A_BYTES, cc, cv_w, cv_h, cv_n, kx, ky);
                } else {
                    const int ak = kb * BK + (int)blockIdx.x * p.a_block_k;
                    tma_load_2d(&tmA, &full[s], st, ak, m0);
                    if (SPLIT == 3) tma_load_2d(&tmA2, &full[s], st + S::A_BYTES, ak, m0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN);
            const unsigned long long m_t0 = p.dbg ? gtime_ns() : 0ull;
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % S::STAGES, ph = (kb / S::STAGES) & 1;
                mbar_wait(&full[s], ph);
                if (p.dbg && kb == 0) atomicAdd(p.dbg + 3, gtime_ns() - m_t0);           // MMA thread: entry -> first k-block landed
                if (p.dbg && kb == nkb - 1) atomicAdd(p.dbg + 4, gtime_ns() - m_t0);     // ... -> last k-block landed
                tcgen05_fence_after();
                const uint32_t a_hi = smem_u32(smem + s * S::STAGE);
                const uint32_t w_hi = a_hi + S::NOPS * S::A_BYTES;
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint32_t koff = k * UMMA_K * 2;
                    umma_bf16(tmem_base, make_smem_desc(a_hi + koff), make_smem_desc(w_hi + koff), idesc, (kb | k) != 0);
                    if (SPLIT == 3) {
                        umma_bf16(tmem_base, make_smem_desc(a_hi + koff), make_smem_desc(w_hi + S::W_BYTES + koff), idesc, 1);
                        umma_bf16(tmem_base, make_smem_desc(a_hi + S::A_BYTES + koff), make_smem_desc(w_hi + koff), idesc, 1);
                    }
                }
                umma_commit(&empty[s]);        // frees the smem slot once these MMAs have read it
            }
            umma_commit(tmem_full);            // accumulator complete
        }
    } else {
        const unsigned long long e_t0 = (p.dbg && threadIdx.x == 64) ? gtime_ns() : 0ull;
        pdl_wait();        // the epilogue reads the residual stream
        epilogue_tile<BN, EPI, TC, EW>(tmem_full, tmem_base, warp, lane, m0, n0, p, smem, 0, e_t0);
        if (p.dbg && threadIdx.x == 64) atomicAdd(p.dbg + 5, gtime_ns() - e_t0);          // epilogue warp: entry -> its rows stored
    }
    if (p.late_trigger == 2 && warp < 2) { mbar_wait(tmem_full, 0); pdl_launch_dependents(); }      // every thread of the CTA releases at the same point
    tcgen05_fence_before();
    __syncthreads();
    if (p.dbg && threadIdx.x == 0) {
        const unsigned long long t2 = gtime_ns();
        atomicAdd(p.dbg, dbg_t1 - dbg_t0); atomicAdd(p.dbg + 1, t2 - dbg_t1); atomicAdd(p.dbg + 2, 1ull);
    }
    if (p.late_trigger == 1) pdl_launch_dependents();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ persistent GEMM
// Large-M GEMMs (backbone convolutions, encoder blocks, cross-K/V): one CTA per SM walks the output tiles (n fastest, so
// neighbouring CTAs share an A tile in L2) with TWO accumulator buffers in TMEM: the MMA warp fills buffer (i + 1) & 1
// while the four epilogue warps drain buffer i & 1, and the TMA ring keeps running across tile boundaries.  In the
// one-tile-per-CTA kernel above the tensor pipe sat idle during every epilogue (ncu: 20 % tensor-pipe active on the conv
// GEMMs); here the epilogue is off the critical path as long as it is shorter than a tile's MMA loop.
template <int BN, int SPLIT, int EW> struct SmemP {
    static constexpr int A_BYTES = BM * BK * 2, W_BYTES = BN * BK * 2;
    static constexpr int NOPS = SPLIT == 3 ? 2 : 1;
    static constexpr int STAGE = NOPS * (A_BYTES + W_BYTES);
    static constexpr int STG = 4 * EW * STG_WARP;                              // dedicated epilogue staging (the ring never idles)
    static constexpr int BUDGET = (BN == 256 ? 225 : 200) * 1024;             // 128 x 256 tiles of split operands: two 96 KB stages
    static constexpr int STAGES = (BUDGET - STG) / STAGE > 8 ? 8 : (BUDGET - STG) / STAGE;
    static_assert(STAGES >= 2, "the ring needs two stages");
    static constexpr int BARS = 256;
    static constexpr int TOTAL = STAGES * STAGE + STG + 1024 + BARS;
};

template <int BN, int EPI, typename TC, int SPLIT, int EW>
__global__ void __launch_bounds__(64 + 128 * EW, 1)
tc_gemm_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                          const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmW2, const TcParams p) {
    using S = SmemP<BN, SPLIT, EW>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int nst = p.stages;
    uint8_t* stg = smem + nst * S::STAGE;
    uint64_t* full = reinterpret_cast<uint64_t*>(stg + S::STG);
    uint64_t* empty = full + S::STAGES;
    uint64_t* tmem_full = empty + S::STAGES;       // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = p.K / BK;
    const int n_tiles = (p.N + BN - 1) / BN, m_tiles = (p.M + BM - 1) / BM;
    const int tiles = n_tiles * m_tiles;

    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
        for (int s = 0; s < nst; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 4 * EW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
                int cv_w = 0, cv_h = 0, cv_n = 0;
                if (p.cv_cpk) {
                    const int per = p.cv_Wo * p.cv_Ho;
                    cv_n = m0 / per;
                    const int rem = m0 - cv_n * per, oh = rem / p.cv_Wo;
                    cv_h = oh * p.cv_stride + p.cv_lower;
                    cv_w = (rem - oh * p.cv_Wo) * p.cv_stride + p.cv_lower;
                }
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % nst, ph = (it / nst) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* st = smem + s * S::STAGE;
                    mbar_expect_tx(&full[s], S::STAGE);
                    if (p.cv_cpk) {
                        const int tap = kb / p.cv_cpk, cc = (kb - tap * p.cv_cpk) * BK;
                        const int ky = tap / p.cv_ksz, kx = tap - ky * p.cv_ksz;
                        tma_load_im2col(&tmA, &full[s], st, cc, cv_w, cv_h, cv_n, kx, ky);
                        if (SPLIT == 3) tma_load_im2col(&tmA2, &full[s], st + S::A_BYTES, cc, cv_w, cv_h, cv_n, kx, ky);
                    } else {
                        tma_load_2d(&tmA, &full[s], st, kb * BK, m0);
                        if (SPLIT == 3) tma_load_2d(&tmA2, &full[s], st + S::A_BYTES, kb * BK, m0);
                    }
                    tma_load_2d(&tmW, &full[s], st + S::NOPS * S::A_BYTES, kb * BK, n0);
                    if (SPLIT == 3) tma_load_2d(&tmW2, &full[s], st + S::NOPS * S::A_BYTES + S::W_BYTES, kb * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN);
            int it = 0, i = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++i) {
                const int buf = i & 1;
                mbar_wait(&tmem_empty[buf], ((i >> 1) & 1) ^ 1);        // the epilogue has drained this buffer (first use passes)
                tcgen05_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)(buf * BN);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % nst, ph = (it / nst) & 1;
                    mbar_wait(&full[s], ph);
                    tcgen05_fence_after();
                    const uint32_t a_hi = smem_u32(smem + s * S::STAGE);
                    const uint32_t w_hi = a_hi + S::NOPS * S::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint32_t koff = k * UMMA_K * 2;
                        umma_bf16(acc, make_smem_desc(a_hi + koff), make_smem_desc(w_hi + koff), idesc, (kb | k) != 0);
                        if (SPLIT == 3) {
                            umma_bf16(acc, make_smem_desc(a_hi + koff), make_smem_desc(w_hi + S::W_BYTES + koff), idesc, 1);
                            umma_bf16(acc, make_smem_desc(a_hi + S::A_BYTES + koff), make_smem_desc(w_hi + koff), idesc, 1);
                        }
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&tmem_full[buf]);
            }
        }
    } else {
        int i = 0, gn_hint = -1;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++i) {
            const int buf = i & 1;
            const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
            epilogue_tile<BN, EPI, TC, EW, true>(&tmem_full[buf], tmem_base + (uint32_t)(buf * BN), warp, lane, m0, n0, p, stg, (uint32_t)((i >> 1) & 1), 0ull, &gn_hint);
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[buf]);              // all of this warp's tcgen05.ld of the buffer have completed
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ ragged implicit-GEMM convolution
// Batches of different-size images cannot use the TMA im2col loads (one regular [N][H][W][C] box per tensor map).  Same persistent
// kernel, bf16x3 operands, fp32 store epilogue -- but the A tiles are gathered by four producer warps.  Thread r owns output pixel r
// of the tile: it locates the pixel's image once per tile (a forward step from the previous tile's image) and computes, per k-block
// (filter tap, 64-channel chunk), the source offset of the pixel's 128-byte row (or "outside the image": TF-SAME padding = zero
// fill).  The copies themselves are issued transposed: eight neighbouring lanes fetch the eight 16-byte chunks of ONE row (offsets by
// shuffle from the owner lane), so a quarter warp reads one full 128-byte line, straight into the 128-byte-swizzled layout the MMA
// descriptors expect.  A stage is published STAGES - 2 k-blocks late (cp.async.wait_group -> fence.proxy.async -> mbarrier arrive):
// the copies of the following k-blocks are in flight while it lands, and one stage of slack stays between the MMA issuer and the
// producers.  W tiles still come by TMA.  Replaces the explicit im2col buffer (72 bytes of traffic per input element of a 3x3
// convolution) for ragged batches.
template <int BN> struct SmemG {
    static constexpr int A_BYTES = BM * BK * 2, W_BYTES = BN * BK * 2;
    static constexpr int STAGE = 2 * (A_BYTES + W_BYTES);
    static constexpr int STG = 4 * STG_WARP;
    static constexpr int STAGES = (225 * 1024 - STG) / STAGE > 6 ? 6 : (225 * 1024 - STG) / STAGE;
    static constexpr int BARS = 256;
    static constexpr int TOTAL = STAGES * STAGE + STG + 1024 + BARS;
};

TX_DEVINL void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

template <int BN>
__global__ void __launch_bounds__(320, 1)
tc_conv_gather_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmW2, const TcParams p) {
    using S = SmemG<BN>;
    constexpr int NST = S::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* stg = smem + NST * S::STAGE;
    uint64_t* full = reinterpret_cast<uint64_t*>(stg + S::STG);
    uint64_t* empty = full + NST;
    uint64_t* tmem_full = empty + NST;             // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = p.K / BK;
    const int n_tiles = (p.N + BN - 1) / BN, m_tiles = (p.M + BM - 1) / BM;
    const int tiles = n_tiles * m_tiles;

    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1 + 128); mbar_init(&empty[s], 1); }      // W producer + 128 gather threads
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                const int n0 = (tile % n_tiles) * BN;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % NST, ph = (it / NST) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* st = smem + s * S::STAGE;
                    mbar_expect_tx(&full[s], 2 * S::W_BYTES);
                    tma_load_2d(&tmW, &full[s], st + 2 * S::A_BYTES, kb * BK, n0);
                    tma_load_2d(&tmW2, &full[s], st + 2 * S::A_BYTES + S::W_BYTES, kb * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN);
            int it = 0, i = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++i) {
                const int buf = i & 1;
                mbar_wait(&tmem_empty[buf], ((i >> 1) & 1) ^ 1);
                tcgen05_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)(buf * BN);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % NST, ph = (it / NST) & 1;
                    mbar_wait(&full[s], ph);
                    tcgen05_fence_after();
                    const uint32_t a_hi = smem_u32(smem + s * S::STAGE);
                    const uint32_t w_hi = a_hi + 2 * S::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint32_t koff = k * UMMA_K * 2;
                        umma_bf16(acc, make_smem_desc(a_hi + koff), make_smem_desc(w_hi + koff), idesc, (kb | k) != 0);
                        umma_bf16(acc, make_smem_desc(a_hi + koff), make_smem_desc(w_hi + S::W_BYTES + koff), idesc, 1);
                        umma_bf16(acc, make_smem_desc(a_hi + S::A_BYTES + koff), make_smem_desc(w_hi + koff), idesc, 1);
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&tmem_full[buf]);
            }
        }
    } else if (warp < 6) {
        int i = 0, gn_hint = -1;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++i) {
            const int buf = i & 1;
            const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
            epilogue_tile<BN, EPI_STORE, float, 1, true>(&tmem_full[buf], tmem_base + (uint32_t)(buf * BN), warp, lane, m0, n0, p, stg, (uint32_t)((i >> 1) & 1), 0ull, &gn_hint);
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[buf]);
        }
    } else {
        // ---- gather producers: thread r owns row r of every A tile
        const int r = (int)threadIdx.x - 192;
        const int cin = p.cv_cpk * BK, ksz = p.cv_ksz, stride = p.cv_stride;
        // k-blocks between issue and publication.  NST - 1 would keep every stage in flight but couples the MMA of k-block j to the
        // release of k-block j - 1 (measured: 6.5 ms of convolutions per config-2 batch against 5.7 ms with one stage of slack)
        constexpr int LAG = NST > 2 ? NST - 2 : 1;
        int it = 0, b_hint = -1;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const int m = (tile / n_tiles) * BM + r;
            int oy = 0, ox = 0, hin = 0, win = 0;
            size_t base = 0;
            const bool row_ok = m < p.M;
            if (row_ok) {
                const int b = find_image_from(p.g_img_off, p.g_nimg, p.g_lout, m, b_hint);
                b_hint = b;
                const int H = p.g_img_hw[2 * b], W = p.g_img_hw[2 * b + 1];
                const int wo = W >> p.g_lout;
                hin = H >> p.g_lin; win = W >> p.g_lin;
                const int local = m - (p.g_img_off[b] >> (2 * p.g_lout));
                oy = local / wo; ox = local - oy * wo;
                base = (size_t)(p.g_img_off[b] >> (2 * p.g_lin)) * cin;
            }
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const int s = it % NST, ph = (it / NST) & 1;
                const int tap = kb / p.cv_cpk, cc = (kb - tap * p.cv_cpk) * BK;
                const int ky = tap / ksz, kx = tap - ky * ksz;
                const int iy = oy * stride + ky - p.g_pad, ix = ox * stride + kx - p.g_pad;
                const bool ok = row_ok && iy >= 0 && iy < hin && ix >= 0 && ix < win;
                const size_t src = ok ? base + ((size_t)iy * win + ix) * cin + cc : 0;
                const uint32_t nbytes = ok ? 16u : 0u;
                mbar_wait(&empty[s], ph ^ 1);
                // the copies are issued transposed: eight neighbouring lanes fetch the eight 16-byte chunks of ONE row (one 128-byte line per
                // quarter warp instead of 32 half-used sectors per instruction); the row's source offset comes from its owner lane by shuffle
                const uint32_t d_base = smem_u32(smem + s * S::STAGE) + (uint32_t)(r & ~31) * 128u;
                const uint32_t ch = (uint32_t)lane & 7u;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = 4 * i + (lane >> 3);
                    const unsigned long long s_i = __shfl_sync(0xffffffffu, (unsigned long long)src, rr);
                    const uint32_t nb_i = __shfl_sync(0xffffffffu, nbytes, rr);
                    const uint32_t dst = d_base + (uint32_t)rr * 128u + ((ch ^ ((uint32_t)rr & 7u)) << 4);
                    cp_async16(dst, p.g_hi + s_i + ch * 8, nb_i);
                    cp_async16(dst + S::A_BYTES, p.g_lo + s_i + ch * 8, nb_i);
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                if (it >= LAG) {       // publish k-block it - LAG: its copies have landed once at most LAG groups are pending
                    asm volatile("cp.async.wait_group %0;" ::"n"(LAG) : "memory");
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive(&full[(it - LAG) % NST]);
                }
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");      // the last LAG k-blocks
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int j = it > LAG ? it - LAG : 0; j < it; ++j) mbar_arrive(&full[j % NST]);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ stem convolution on the tensor cores
// model/resnet.py:219 (StdConv 7x7 / stride 2, one input channel, TF-SAME padding (2,3), 64 output channels) as an implicit GEMM
// with K = 49 taps padded to 64, bf16x3: the standardised filter bank (8 + 8 KB of hi / lo bf16) stays in shared memory for the whole
// kernel, eight producer warps build the A tiles -- thread = output pixel: 49 image loads, split into hi / lo bf16, eight 16-byte
// stores per half into the 128-byte-swizzled rows -- and the usual MMA issuer / four epilogue warps follow; the epilogue leaves the
// GroupNorm partial sums of the 64-channel output (every image has a multiple of 64 level-1 pixels, so this works for any batch).
// The FFMA kernel it replaces in the bf16 tier ran at 43 % of the fp32 SIMT peak (0.63 ms per 512 images) plus a statistics pass.
struct SmemStem {
    static constexpr int A_BYTES = BM * BK * 2, W_BYTES = 64 * BK * 2;
    static constexpr int STAGE = 2 * A_BYTES, STAGES = 4;
    static constexpr int STG = 4 * STG_WARP;
    static constexpr int TOTAL = 2 * W_BYTES + STAGES * STAGE + STG + 1024 + 256;
};
struct StemParams {
    const float* img; const int* img_off; const int* img_hw; int nimg;
    int rpi;           // > 0: every image has this many level-1 pixels (same-size batch)
};

__global__ void __launch_bounds__(448, 1)
tc_stem_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmW2, const TcParams p, const StemParams sp) {
    using S = SmemStem;
    constexpr int NST = S::STAGES, BN = 64;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* wsm = smem;                                   // W_hi | W_lo
    uint8_t* ring = smem + 2 * S::W_BYTES;
    uint8_t* stg = ring + NST * S::STAGE;
    uint64_t* full = reinterpret_cast<uint64_t*>(stg + S::STG);
    uint64_t* empty = full + NST;
    uint64_t* tmem_full = empty + NST;             // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint64_t* wbar = tmem_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles = (p.M + BM - 1) / BM;

    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 128); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 4); }
        mbar_init(wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 0 && lane == 0) {          // the filter bank is a weight: fetched before the dependency wait
        mbar_expect_tx(wbar, 2 * S::W_BYTES);
        tma_load_2d(&tmW, wbar, wsm, 0, 0);
        tma_load_2d(&tmW2, wbar, wsm + S::W_BYTES, 0, 0);
    }
    pdl_wait();

    if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN);
            mbar_wait(wbar, 0);
            const uint32_t w_hi = smem_u32(wsm);
            int i = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++i) {
                const int buf = i & 1, s = i % NST, ph = (i / NST) & 1;
                mbar_wait(&tmem_empty[buf], ((i >> 1) & 1) ^ 1);
                mbar_wait(&full[s], ph);
                tcgen05_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)(buf * BN);
                const uint32_t a_hi = smem_u32(ring + s * S::STAGE);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint32_t koff = k * UMMA_K * 2;
                    umma_bf16(acc, make_smem_desc(a_hi + koff), make_smem_desc(w_hi + koff), idesc, k != 0);
                    umma_bf16(acc, make_smem_desc(a_hi + koff), make_smem_desc(w_hi + S::W_BYTES + koff), idesc, 1);
                    umma_bf16(acc, make_smem_desc(a_hi + S::A_BYTES + koff), make_smem_desc(w_hi + koff), idesc, 1);
                }
                umma_commit(&empty[s]);
                umma_commit(&tmem_full[buf]);
            }
        }
    } else if (warp >= 2 && warp < 6) {
        int i = 0, gn_hint = -1;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++i) {
            const int buf = i & 1;
            epilogue_tile<BN, EPI_STORE, float, 1, true>(&tmem_full[buf], tmem_base + (uint32_t)(buf * BN), warp, lane, tile * BM, 0, p, stg, (uint32_t)((i >> 1) & 1), 0ull, &gn_hint);
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[buf]);
        }
    } else if (warp >= 6) {
        // ---- producers: two groups of four warps take alternate tiles of this CTA (a thread builds its row start to finish, so one
        // group alone is one tile per ~2.5 us and SM); thread r of a group builds row r (one output pixel) of its A tiles
        const int grp = (warp - 6) >> 2;
        const int r = ((int)threadIdx.x - 192) & 127;
        const uint32_t sw = (uint32_t)(r & 7);
        int i = grp, b_hint = -1;
        for (int tile = blockIdx.x + grp * gridDim.x; tile < tiles; tile += 2 * gridDim.x, i += 2) {
            const int s = i % NST, ph = (i / NST) & 1;
            const int m = tile * BM + r;
            float v[56];
#pragma unroll
            for (int t = 0; t < 56; ++t) v[t] = 0.f;
            if (m < p.M) {
                const int b = sp.rpi > 0 ? m / sp.rpi : find_image_from(sp.img_off, sp.nimg, 1, m, b_hint);      // same-size batch: a division
                b_hint = b;
                const int H = sp.img_hw[2 * b], W = sp.img_hw[2 * b + 1], W1 = W >> 1;
                const int local = m - (sp.img_off[b] >> 2);
                const int oy = local / W1, ox = local - oy * W1;
                const float* im = sp.img + sp.img_off[b];
#pragma unroll
                for (int ky = 0; ky < 7; ++ky) {
                    const int iy = 2 * oy + ky - 2;
                    const bool rowok = iy >= 0 && iy < H;
#pragma unroll
                    for (int kx = 0; kx < 7; ++kx) {
                        const int ix = 2 * ox + kx - 2;
                        if (rowok && ix >= 0 && ix < W) v[ky * 7 + kx] = __ldg(im + (size_t)iy * W + ix);
                    }
                }
            }
            mbar_wait(&empty[s], ph ^ 1);
            uint8_t* a_hi = ring + s * S::STAGE + r * 128;
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j) {
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float x0 = j < 7 ? v[(j < 7 ? j : 0) * 8 + 2 * e] : 0.f, x1 = j < 7 ? v[(j < 7 ? j : 0) * 8 + 2 * e + 1] : 0.f;
                    const float h0 = __bfloat162float(__float2bfloat16_rn(x0)), h1 = __bfloat162float(__float2bfloat16_rn(x1));
                    const __nv_bfloat162 hh = __floats2bfloat162_rn(h0, h1), ll = __floats2bfloat162_rn(x0 - h0, x1 - h1);
                    hw[e] = *reinterpret_cast<const uint32_t*>(&hh); lw[e] = *reinterpret_cast<const uint32_t*>(&ll);
                }
                *reinterpret_cast<uint4*>(a_hi + ((j ^ sw) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4*>(a_hi + S::A_BYTES + ((j ^ sw) << 4)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the MMA's async-proxy reads
            mbar_arrive(&full[s]);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::mutex g_mu;
std::map<std::tuple<const void*, long, int, long, int, int, int>, CUtensorMap> g_maps;

cudaError_t get_encode() {
    if (g_encode) return cudaSuccess;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess) return e;
    if (qres != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
    g_encode = (EncodeTiledFn)fn;
    return cudaSuccess;
}

// GEMM operand map: box = box_rows x 64 columns, 128B swizzle
cudaError_t get_map(const void* ptr, int rows, int cols, int ld, int box_rows, CUtensorMap* out) {
    return tma_map_2d_bf16(ptr, rows, cols, ld, box_rows, BK, 1, out);
}

}  // namespace

// 2-D bf16 tensor [rows, cols] with row stride ld (elements); zero OOB fill; cached per (ptr, shape, box).
cudaError_t tma_map_2d_bf16(const void* ptr, long rows, int cols, long ld, int box_rows, int box_cols, int swizzle128,
                            CUtensorMap* out) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto key = std::make_tuple(ptr, rows, cols, ld, box_rows, box_cols, swizzle128);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return cudaSuccess; }
    cudaError_t e = get_encode();
    if (e != cudaSuccess) return e;
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps[key] = m;
    *out = m;
    return cudaSuccess;
}

// 3-D bf16 tensor [d2][d1][d0] (d0 contiguous; strides of d1 / d2 in bytes); box = b2 x b1 x b0, 128-byte swizzle, zero OOB fill.
cudaError_t tma_map_3d_bf16(const void* ptr, int d0, long d1, long d2, long stride1_bytes, long stride2_bytes, int b0, int b1, int b2,
                            CUtensorMap* out) {
    std::lock_guard<std::mutex> lk(g_mu);
    cudaError_t e = get_encode();
    if (e != cudaSuccess) return e;
    cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
    cuuint64_t strides[2] = {(cuuint64_t)stride1_bytes, (cuuint64_t)stride2_bytes};
    cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// im2col-mode map of an NHWC bf16 activation [N][H][W][C]: box = 128 output pixels x 64 channels, 128-byte swizzle
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*,
                                   const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeIm2colFn g_encode_im2col = nullptr;
static cudaError_t tma_map_im2col_bf16(const void* ptr, const GemmArgs::Im2col& c, CUtensorMap* out) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_encode_im2col) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess) return e;
        if (qres != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        g_encode_im2col = (EncodeIm2colFn)fn;
    }
    cuuint64_t dims[4] = {(cuuint64_t)c.C, (cuuint64_t)c.W, (cuuint64_t)c.H, (cuuint64_t)c.N};
    cuuint64_t strides[3] = {(cuuint64_t)c.C * 2, (cuuint64_t)c.W * c.C * 2, (cuuint64_t)c.H * c.W * c.C * 2};
    // fprop corners (see CUTLASS conv/collective/detail.hpp): lower = -pad_before, upper = pad_after - (ksz - 1)
    const int pad_total = std::max((c.Wo - 1) * c.stride + c.ksz - c.W, 0);
    const int pad_total_h = std::max((c.Ho - 1) * c.stride + c.ksz - c.H, 0);
    int lower[2] = {-c.pad_lo, -c.pad_lo};
    int upper[2] = {(pad_total - c.pad_lo) - (c.ksz - 1), (pad_total_h - c.pad_lo) - (c.ksz - 1)};
    cuuint32_t estr[4] = {1, (cuuint32_t)c.stride, (cuuint32_t)c.stride, 1};
    CUresult r = g_encode_im2col(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, lower, upper, BK, BM, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

namespace {

template <int BN, int EPI, typename TC, int SPLIT, int NSTG, int EW>
cudaError_t launch_cfg3(const CUtensorMap& a, const CUtensorMap& w, const CUtensorMap& a2, const CUtensorMap& w2, const TcParams& p,
                        cudaStream_t st) {
    using S = Smem<BN, SPLIT, NSTG>;
    static_assert(S::STAGES * S::STAGE >= 4 * EW * STG_WARP, "epilogue staging lives in the idle pipeline stages");
    static bool attr_set = false;
    auto kern = tc_gemm_kernel<BN, EPI, TC, SPLIT, NSTG, EW>;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM);
    return launch_pdl(PDL_GEMM, kern, grid, dim3(64 + 128 * EW), (size_t)S::TOTAL, st, a, w, a2, w2, p);
}
template <int BN, int EPI, typename TC, int SPLIT, int NSTG>
cudaError_t launch_cfg2(const CUtensorMap& a, const CUtensorMap& w, const CUtensorMap& a2, const CUtensorMap& w2, const TcParams& p,
                        cudaStream_t st) {
    if constexpr (EPI != EPI_ARGMAX) {
        if (g_tc_epi_warps == 8) return launch_cfg3<BN, EPI, TC, SPLIT, NSTG, 2>(a, w, a2, w2, p, st);
    }
    return launch_cfg3<BN, EPI, TC, SPLIT, NSTG, 1>(a, w, a2, w2, p, st);
}

template <int BN, int EPI, typename TC, int SPLIT, int EW>
cudaError_t launch_persistent2(const CUtensorMap& a, const CUtensorMap& w, const CUtensorMap& a2, const CUtensorMap& w2, const TcParams& p,
                               long tiles, cudaStream_t st) {
    using S = SmemP<BN, SPLIT, EW>;
    static bool attr_set = false;
    static int sms = 0;
    auto kern = tc_gemm_persistent_kernel<BN, EPI, TC, SPLIT, EW>;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) return e;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        attr_set = true;
    }
    const unsigned grid = (unsigned)std::min<long>(tiles, sms);
    TcParams pp = p;
    pp.stages = g_tc_persistent_stages > 0 ? std::max(2, std::min(g_tc_persistent_stages, S::STAGES)) : S::STAGES;
    const size_t smem = (size_t)S::TOTAL - (size_t)(S::STAGES - pp.stages) * S::STAGE;
    return launch_pdl(PDL_GEMM, kern, dim3(grid), dim3(64 + 128 * EW), smem, st, a, w, a2, w2, pp);
}
template <int BN, int EPI, typename TC, int SPLIT>
cudaError_t launch_persistent(const CUtensorMap& a, const CUtensorMap& w, const CUtensorMap& a2, const CUtensorMap& w2, const TcParams& p,
                              long tiles, cudaStream_t st) {
    if (g_tc_epi_warps == 8) return launch_persistent2<BN, EPI, TC, SPLIT, 2>(a, w, a2, w2, p, tiles, st);
    return launch_persistent2<BN, EPI, TC, SPLIT, 1>(a, w, a2, w2, p, tiles, st);
}

template <int BN>
cudaError_t launch_gather(const CUtensorMap& w, const CUtensorMap& w2, const TcParams& p, cudaStream_t st) {
    using S = SmemG<BN>;
    static bool attr_set = false;
    static int sms = 0;
    auto kern = tc_conv_gather_kernel<BN>;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) return e;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        attr_set = true;
    }
    const long tiles = (long)((p.N + BN - 1) / BN) * ((p.M + BM - 1) / BM);
    const unsigned grid = (unsigned)std::min<long>(tiles, sms);
    return launch_pdl(PDL_GEMM, kern, dim3(grid), dim3(320), (size_t)S::TOTAL, st, w, w2, p);
}

template <int BN, int EPI, typename TC, int SPLIT>
cudaError_t launch_cfg(const CUtensorMap& a, const CUtensorMap& w, const CUtensorMap& a2, const CUtensorMap& w2, const TcParams& p,
                       cudaStream_t st) {
    // many tiles (encoder / teacher-forced sizes): shallow pipeline, several CTAs per SM; few tiles: deep pipeline
    const long tiles = (long)((p.N + BN - 1) / BN) * ((p.M + BM - 1) / BM);
    if constexpr (BN >= 64) {
        if (g_tc_persistent && tiles >= 296 && !p.a_block_k) return launch_persistent<BN, EPI, TC, SPLIT>(a, w, a2, w2, p, tiles, st);
    }
    if (SPLIT == 1 && BN >= 64 && tiles >= 592) return launch_cfg2<BN, EPI, TC, SPLIT, 2>(a, w, a2, w2, p, st);
    return launch_cfg2<BN, EPI, TC, SPLIT, 0>(a, w, a2, w2, p, st);
}

template <int BN, int SPLIT>
cudaError_t launch_epi(const GemmArgs& g, const CUtensorMap& a, const CUtensorMap& w, const CUtensorMap& a2, const CUtensorMap& w2,
                       const TcParams& p, cudaStream_t st) {
    switch (g.epi) {
        case EPI_STORE:
            if (g.dt_c == DT_F32) return launch_cfg<BN, EPI_STORE, float, SPLIT>(a, w, a2, w2, p, st);
            return launch_cfg<BN, EPI_STORE, bf16, SPLIT>(a, w, a2, w2, p, st);
        case EPI_GLU_RES: return launch_cfg<BN, EPI_GLU_RES, float, SPLIT>(a, w, a2, w2, p, st);
        case EPI_GEGLU: return launch_cfg<BN, EPI_GEGLU, bf16, SPLIT>(a, w, a2, w2, p, st);
        case EPI_BIAS_RES: return launch_cfg<BN, EPI_BIAS_RES, float, SPLIT>(a, w, a2, w2, p, st);
        case EPI_ARGMAX:
            if constexpr (BN == 32 && SPLIT == 1) return launch_cfg2<32, EPI_ARGMAX, float, 1, 0>(a, w, a2, w2, p, st);
            return cudaErrorInvalidValue;
    }
    return cudaErrorInvalidValue;
}

}  // namespace

// Stem convolution of the bf16 tier (tc_stem_kernel): img = ragged fp32 images, w_hi / w_lo = [64][64] bf16 (49 standardised taps + 15
// zero columns per output channel), raw1 [total_p1, 64] fp32, gn_part = GroupNorm block partials of raw1 (gn_block.cuh).
cudaError_t launch_stem_tc(const float* img, const void* w_hi, const void* w_lo, float* raw1, const int* img_off, const int* img_hw, int nimg,
                           long total_p1, int uniform_rpi, float* gn_part, cudaStream_t st) {
    if (total_p1 <= 0) return cudaSuccess;
    if (total_p1 % 32 != 0 || total_p1 > 0x7fffffffL) return cudaErrorInvalidValue;
    TcParams p{raw1, (int)total_p1, 64, 64, 64, nullptr, nullptr, 0, 0, 0, 0, 0, 0, 0, 0};
    p.stages = 0; p.a_block_k = 0; p.dbg = nullptr;
    p.gn_part = gn_part; p.gn_cpg = 2; p.gn_rpi = uniform_rpi > 0 ? uniform_rpi : 0; p.gn_img_off = img_off; p.gn_nimg = nimg; p.gn_level = 1;
    p.g_hi = p.g_lo = nullptr; p.g_img_off = p.g_img_hw = nullptr; p.g_nimg = p.g_lin = p.g_lout = p.g_pad = 0;
    StemParams sp{img, img_off, img_hw, nimg, uniform_rpi > 0 ? uniform_rpi : 0};
    CUtensorMap w, w2;
    cudaError_t e;
    if ((e = get_map(w_hi, 64, 64, 64, 64, &w)) != cudaSuccess) return e;
    if ((e = get_map(w_lo, 64, 64, 64, 64, &w2)) != cudaSuccess) return e;
    static bool attr_set = false;
    static int sms = 0;
    if (!attr_set) {
        if ((e = cudaFuncSetAttribute(tc_stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemStem::TOTAL)) != cudaSuccess) return e;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        attr_set = true;
    }
    const long tiles = (total_p1 + BM - 1) / BM;
    const unsigned grid = (unsigned)std::min<long>(tiles, sms);
    return launch_pdl(PDL_GEMM, tc_stem_kernel, dim3(grid), dim3(448), (size_t)SmemStem::TOTAL, st, w, w2, p, sp);
}

bool tc_gemm_supported(const GemmArgs& g) {
    if (g.dt_a != DT_BF16 || g.conv) return false;
    if (g.gather) return g.A2 && g.W2 && g.gather->cin % BK == 0 && g.N % 8 == 0 && g.ldw % 8 == 0 && g.ldc % 4 == 0 && !((uintptr_t)g.A & 15) && !((uintptr_t)g.A2 & 15);
    if (g.K % BK != 0 || g.N % 8 != 0 || g.lda % 8 != 0 || g.ldw % 8 != 0) return false;
    if (g.im2col.ksz > 0 && (g.im2col.C % BK != 0 || g.im2col.ksz > 7)) return false;
    if (((uintptr_t)g.A | (uintptr_t)g.W) & 15) return false;
    if (g.epi == EPI_ARGMAX) return g.A2 == nullptr && g.im2col.ksz == 0 && !g.a_block_k && g.ldc >= (g.N + 31) / 32;
    if (g.epi == EPI_STORE) { if (g.ldc % 8 != 0) return false; }
    else if (g.ldc % 4 != 0) return false;
    if ((g.epi == EPI_GLU_RES || g.epi == EPI_BIAS_RES) && (!g.res || g.ldres % 4 != 0)) return false;
    return true;
}

cudaError_t launch_gemm_tc(const GemmArgs& g, cudaStream_t st) {
    if (g.M <= 0) return cudaSuccess;
    // tile width: keep >= ~1 wave of CTAs when M is small (decode steps), 128 otherwise
    const long mt = (g.M + BM - 1) / BM;
    int bn = 128;
    if (mt * ((g.N + 127) / 128) < g_tc_min_ctas) bn = 64;
    if (mt * ((g.N + 63) / 64) < g_tc_min_ctas && g.N >= 64) bn = 32;
    const bool split = g.A2 != nullptr;
    if (split) bn = g.N <= 64 ? 64 : 128;
    // 128 x 256 tiles (one MMA reads A once for 256 columns: the shared-memory operand reads of a 128 x 128 x 16 MMA take as long as
    // the MMA itself) for the wide, deep convolutions; K <= 256 ones are bound by the epilogue, which has four warps here
    // (ncu: 9.4 us per 128 x 256 tile at K = 256 against 3.2 us of MMAs)
    if (split && g_tc_bn256 && g.epi == EPI_STORE && g.dt_c == DT_F32 && !g.bias && g.N % 256 == 0 && g.K >= 512 && mt * (g.N / 256) >= 148 && g_tc_persistent) bn = 256;
    if (g.a_block_k) {
        if (split || g.im2col.ksz > 0 || g.N % 64 != 0) return cudaErrorInvalidValue;
        bn = 64;
    }
    if (g.epi == EPI_ARGMAX) bn = 32;      // the partial layout is defined on 32-column tiles
    TcParams p{g.C, g.M, g.N, g.K, g.ldc, g.bias, g.res, g.ldres, (g_texocr_pdl_mid & 1) ? 2 : ((g_texocr_pdl >> 9) & 1), 0, 0, 0, 0, 0, 0};
    p.stages = 0; p.gn_part = nullptr; p.gn_cpg = 0; p.gn_rpi = 0; p.gn_img_off = nullptr; p.gn_nimg = 0; p.gn_level = 0;
    p.g_hi = p.g_lo = nullptr; p.g_img_off = p.g_img_hw = nullptr; p.g_nimg = p.g_lin = p.g_lout = p.g_pad = 0;
    p.a_block_k = g.a_block_k; p.dbg = g.dbg;
    if (g.gn_part) {
        if (g.epi != EPI_STORE || g.dt_c != DT_F32 || g.bias || g.N % 32 != 0 || g.M % 32 != 0 || g.N / 32 > 32 || (g.N / 32) & (g.N / 32 - 1))
            return cudaErrorInvalidValue;
        if (g.gn_rpi > 0 ? (g.gn_rpi % 32 != 0 || g.M % g.gn_rpi != 0) : (!g.gn_img_off || g.gn_nimg <= 0)) return cudaErrorInvalidValue;
        p.gn_part = g.gn_part; p.gn_cpg = g.N / 32; p.gn_rpi = g.gn_rpi;
        p.gn_img_off = g.gn_img_off; p.gn_nimg = g.gn_nimg; p.gn_level = g.gn_level;
    }
    CUtensorMap a, w, a2, w2;
    cudaError_t e;
    if (g.gather) {         // ragged implicit-GEMM convolution: A gathered by producer warps, W by TMA
        const ConvGather& cg = *g.gather;
        if (!split || g.epi != EPI_STORE || g.dt_c != DT_F32 || g.bias || cg.cin % BK != 0 || g.K != cg.ksz * cg.ksz * cg.cin)
            return cudaErrorInvalidValue;
        bn = g.N <= 64 ? 64 : 128;
        p.cv_cpk = cg.cin / BK; p.cv_ksz = cg.ksz; p.cv_stride = cg.stride;
        p.g_hi = (const bf16*)g.A; p.g_lo = (const bf16*)g.A2; p.g_img_off = cg.img_off; p.g_img_hw = cg.img_hw; p.g_nimg = cg.nimg;
        p.g_lin = cg.lin; p.g_lout = cg.lout; p.g_pad = cg.pad;
        if ((e = get_map(g.W, g.N, g.K, g.ldw, bn, &w)) != cudaSuccess) return e;
        if ((e = get_map(g.W2, g.N, g.K, g.ldw, bn, &w2)) != cudaSuccess) return e;
        return bn == 64 ? launch_gather<64>(w, w2, p, st) : launch_gather<128>(w, w2, p, st);
    }
    const bool conv = g.im2col.ksz > 0;
    if (conv) {
        const GemmArgs::Im2col& c = g.im2col;
        if (c.C % BK != 0 || g.K != c.ksz * c.ksz * c.C || (long)c.N * c.Wo * c.Ho != g.M) return cudaErrorInvalidValue;
        p.cv_cpk = c.C / BK; p.cv_ksz = c.ksz; p.cv_stride = c.stride; p.cv_lower = -c.pad_lo; p.cv_Wo = c.Wo; p.cv_Ho = c.Ho;
        if ((e = tma_map_im2col_bf16(g.A, c, &a)) != cudaSuccess) return e;
    } else if ((e = get_map(g.A, g.M, g.a_block_k ? (g.N / 64) * g.a_block_k : g.K, g.lda, BM, &a)) != cudaSuccess) return e;
    if ((e = get_map(g.W, g.N, g.K, g.ldw, bn, &w)) != cudaSuccess) return e;
    a2 = a; w2 = w;
    if (split) {
        if (conv) { if ((e = tma_map_im2col_bf16(g.A2, g.im2col, &a2)) != cudaSuccess) return e; }
        else if ((e = get_map(g.A2, g.M, g.K, g.lda, BM, &a2)) != cudaSuccess) return e;
        if ((e = get_map(g.W2, g.N, g.K, g.ldw, bn, &w2)) != cudaSuccess) return e;
        if (bn == 64) return launch_epi<64, 3>(g, a, w, a2, w2, p, st);
        if (bn == 256) return launch_persistent2<256, EPI_STORE, float, 3, 1>(a, w, a2, w2, p, mt * (g.N / 256), st);
        return launch_epi<128, 3>(g, a, w, a2, w2, p, st);
    }
    if (bn == 128) return launch_epi<128, 1>(g, a, w, a2, w2, p, st);
    if (bn == 64) return launch_epi<64, 1>(g, a, w, a2, w2, p, st);
    return launch_epi<32, 1>(g, a, w, a2, w2, p, st);
}
