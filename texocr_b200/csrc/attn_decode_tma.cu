// Decode-step attention for the bf16 tier: a persistent, TMA-fed streaming kernel (the HBM-bound kernel of the generate loop).
//
// Work unit = (sequence b, head pair hp): 2 heads x nk keys x (K 128 B + V 128 B).  Units are dealt round-robin to
// persistent CTAs (3 per SM).  Each CTA has one producer thread that issues 2-D TMA loads (box = 16 keys x 128 columns,
// K and V tiles of the head pair) into an 8-stage shared-memory ring, running ahead across unit boundaries so the memory
// pipe never drains, and two consumer warps (one per head) that do the online-softmax attention from shared memory:
// lane = (key group of 4, 8-dim slice) -> 16-byte conflict-free smem reads, 3 shuffles per 4 keys.
// Self-attention also folds in this step's own key/value (from the QKV GEMM output) and appends it to the cache.
//
// Reference semantics: model/attention.py:148-173 with q length 1 (energy * 0.125, softmax, . v); no masks are active in
// generate (model/decoder.py:95: mask all True; causal over a prefix == attend to everything cached).
#include "common.cuh"
#include "tc_gemm.h"

namespace {

constexpr int CH = 16;            // keys per stage
constexpr int NS = 8;             // ring stages
constexpr int TILE = CH * 256;    // bytes of one K (or V) tile: CH keys x 2 heads x 64 x bf16
constexpr int STAGE = 2 * TILE;
constexpr float SCALE = 0.125f;

struct Args {
    const bf16* q; int ldq;
    const bf16* knew; const bf16* vnew; int ldnew;     // self only
    bf16* cache; int tcap;                              // self: [B*tcap][1024] rows of this layer (append target)
    const int* k_off;                                   // cross: token offsets [B+1]
    const int* step;                                    // self: keys already cached
    bf16* o; int ldo;
    int batch, col0;
};

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
TX_DEVINL void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
TX_DEVINL void unpack8(const uint4& r, float* o) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
}

// WPH = consumer warps per head: the 16 keys of a stage are dealt to them in groups of 4, each keeps its own online-softmax
// state and the partial states are merged once per unit (more independent dependency chains per SM: the per-warp chain
// LDS -> FMA -> 3 shuffles -> max -> exp -> FMA is latency-bound).
template <bool SELF, int WPH>
__global__ void __launch_bounds__(32 * (1 + 2 * WPH)) attn_decode_tma_kernel(const __grid_constant__ CUtensorMap tm,
                                                                           const __grid_constant__ CUtensorMap tm4, const Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* ring = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + NS * STAGE);
    uint64_t* empty = full + NS;
    float* scratch = reinterpret_cast<float*>(empty + NS);          // [2 parity][2 heads][WPH][8 dg][10]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm4) : "memory");
        for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 2 * WPH); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_wait();

    const int units = a.batch * 4;
    const int t = SELF ? *a.step : 0;

    if (warp == 0) {
        // ------------------------------------------------------------ producer
        if (lane == 0) {
            int it = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int b = u >> 2, hp = u & 3;
                int row0, nk;
                if (SELF) { row0 = b * a.tcap; nk = t; }
                else { row0 = a.k_off[b]; nk = a.k_off[b + 1] - row0; }
                const int nchunk = (nk + CH - 1) / CH;
                for (int c = 0; c < nchunk; ++c, ++it) {
                    const int s = it % NS, ph = (it / NS) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* kt = ring + s * STAGE;
                    const int left = nk - c * CH;
                    const int kc = a.col0 + hp * 128, r = row0 + c * CH;
                    if (left >= CH) {
                        mbar_expect_tx(&full[s], STAGE);
                        tma_load_2d(&tm, &full[s], kt, kc, r);
                        tma_load_2d(&tm, &full[s], kt + TILE, kc + 512, r);
                    } else {               // tail: 4-row boxes, at most 3 rows fetched beyond the sequence
                        const int n4 = (left + 3) >> 2;
                        mbar_expect_tx(&full[s], n4 * 2 * 1024);
                        for (int j = 0; j < n4; ++j) {
                            tma_load_2d(&tm4, &full[s], kt + j * 1024, kc, r + 4 * j);
                            tma_load_2d(&tm4, &full[s], kt + TILE + j * 1024, kc + 512, r + 4 * j);
                        }
                    }
                }
            }
        }
        return;
    }
    // ---------------------------------------------------------------- consumers
    const int w = warp - 1, hd = w / WPH, part = w % WPH;
    const int kg = lane >> 3, dg = lane & 7;
    constexpr int ITER = 4 / WPH;          // key groups of 4 per warp per stage
    int it = 0, parity = 0;
    uint4 q_raw = make_uint4(0, 0, 0, 0), kn_raw = make_uint4(0, 0, 0, 0), vn_raw = make_uint4(0, 0, 0, 0);
    auto load_header = [&](int u) {
        const int b = u >> 2, h = (u & 3) * 2 + hd;
        q_raw = *reinterpret_cast<const uint4*>(a.q + (size_t)b * a.ldq + h * 64 + dg * 8);
        if (SELF && part == 0) {
            kn_raw = *reinterpret_cast<const uint4*>(a.knew + (size_t)b * a.ldnew + h * 64 + dg * 8);
            vn_raw = *reinterpret_cast<const uint4*>(a.vnew + (size_t)b * a.ldnew + h * 64 + dg * 8);
        }
    };
    if ((int)blockIdx.x < units) load_header(blockIdx.x);
    for (int u = blockIdx.x; u < units; u += gridDim.x, parity ^= 1) {
        const int b = u >> 2, h = (u & 3) * 2 + hd;
        float q8[8];
        unpack8(q_raw, q8);
        const uint4 kn_keep = kn_raw, vn_keep = vn_raw;
        const int un = u + gridDim.x;
        if (un < units) load_header(un);          // prefetch the next unit's q / new k,v while this one streams
        int nk;
        if (SELF) nk = t; else nk = a.k_off[b + 1] - a.k_off[b];
        const int nchunk = (nk + CH - 1) / CH;
        float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        for (int c = 0; c < nchunk; ++c, ++it) {
            const int s = it % NS, ph = (it / NS) & 1;
            mbar_wait(&full[s], ph);
            const uint8_t* kt = ring + s * STAGE + hd * 128 + dg * 16;
            const uint8_t* vt = kt + TILE;
            float sc[ITER];
            bool ok[ITER];
#pragma unroll
            for (int i = 0; i < ITER; ++i) {
                const int kl = kg + 4 * (i * WPH + part);
                ok[i] = c * CH + kl < nk;
                float k8[8];
                unpack8(*reinterpret_cast<const uint4*>(kt + kl * 256), k8);
                float d = 0.f;
#pragma unroll
                for (int e = 0; e < 8; ++e) d = fmaf(q8[e], k8[e], d);
                d += __shfl_xor_sync(0xffffffffu, d, 1);
                d += __shfl_xor_sync(0xffffffffu, d, 2);
                d += __shfl_xor_sync(0xffffffffu, d, 4);
                sc[i] = ok[i] ? d * SCALE : -INFINITY;       // rows past the sequence may hold anything (even NaN): never used
            }
            float cm = sc[0];
#pragma unroll
            for (int i = 1; i < ITER; ++i) cm = fmaxf(cm, sc[i]);
            cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 8));
            cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 16));
            const float mn = fmaxf(m, cm);
            if (mn != -INFINITY) {                            // this warp's share of a tail stage may be empty
                const float corr = __expf(m - mn);
                l *= corr;
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] *= corr;
#pragma unroll
                for (int i = 0; i < ITER; ++i) {
                    if (ok[i]) {
                        const float p = __expf(sc[i] - mn);
                        l += p;
                        float v8[8];
                        unpack8(*reinterpret_cast<const uint4*>(vt + (kg + 4 * (i * WPH + part)) * 256), v8);
#pragma unroll
                        for (int e = 0; e < 8; ++e) acc[e] = fmaf(p, v8[e], acc[e]);
                    }
                }
                m = mn;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        // fold the 4 key groups of this warp
        l += __shfl_xor_sync(0xffffffffu, l, 8);
        l += __shfl_xor_sync(0xffffffffu, l, 16);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
            acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
        }
        if (WPH > 1) {
            float* slot = scratch + ((((parity * 2 + hd) * WPH + part) * 8 + dg) * 10);
            if (part != 0 && kg == 0) {
                slot[0] = m; slot[1] = l;
#pragma unroll
                for (int e = 0; e < 8; ++e) slot[2 + e] = acc[e];
            }
            asm volatile("bar.sync %0, %1;" ::"r"(1 + hd), "r"(32 * WPH) : "memory");
            if (part != 0) continue;
#pragma unroll
            for (int p2 = 1; p2 < WPH; ++p2) {
                const float* o2 = scratch + ((((parity * 2 + hd) * WPH + p2) * 8 + dg) * 10);
                const float m2 = o2[0];
                const float mn = fmaxf(m, m2);
                if (mn != -INFINITY) {
                    const float c1 = __expf(m - mn), c2 = __expf(m2 - mn);
                    l = l * c1 + o2[1] * c2;
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[e] = acc[e] * c1 + o2[2 + e] * c2;
                    m = mn;
                }
            }
        }
        if (SELF) {
            // this step's own key / value (from the QKV GEMM output); also appended to the cache for the next steps
            float kn8[8], vn8[8];
            unpack8(kn_keep, kn8);
            unpack8(vn_keep, vn8);
            float d = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) d = fmaf(q8[e], kn8[e], d);
            d += __shfl_xor_sync(0xffffffffu, d, 1);
            d += __shfl_xor_sync(0xffffffffu, d, 2);
            d += __shfl_xor_sync(0xffffffffu, d, 4);
            d *= SCALE;
            const float mn = fmaxf(m, d);
            const float corr = __expf(m - mn), p = __expf(d - mn);
            l = l * corr + p;
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = fmaf(p, vn8[e], acc[e] * corr);
            if (kg == 0) {      // row t of sequence b: K at column h*64, V at 512 + h*64
                bf16* row = a.cache + ((size_t)b * a.tcap + t) * 1024 + h * 64 + dg * 8;
                *reinterpret_cast<uint4*>(row) = kn_keep;
                *reinterpret_cast<uint4*>(row + 512) = vn_keep;
            }
        }
        if (kg == 0) {
            const float inv = 1.0f / l;
            bf16* o = a.o + (size_t)b * a.ldo + h * 64 + dg * 8;
            st4(o, make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv));
            st4(o + 4, make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv));
        }
    }
}

int g_smem_set = 0;

}  // namespace

bool attn_decode_tma_supported(const AttnDecodeArgs& a) {
    if (a.dt != DT_BF16 || a.ldkv % 8 != 0 || a.ldq % 8 != 0 || a.ldo % 8 != 0) return false;
    if (a.knew && (a.ldkv != 1024 || a.ldnew % 8 != 0)) return false;
    return true;
}

// map_rows: number of rows of the K/V matrix the tensor map covers (self: B*tcap of this layer; cross: total tokens);
// map_base: its first row (column 0); col0: column of head 0's K inside a row (cross: layer*1024).
cudaError_t launch_attn_decode_tma(const AttnDecodeArgs& a, const void* map_base, long map_rows, int map_cols, int col0, int tcap,
                                   int max_ctas, cudaStream_t st) {
    if (a.batch <= 0) return cudaSuccess;
    CUtensorMap tm, tm4;
    cudaError_t e = tma_map_2d_bf16(map_base, map_rows, map_cols, a.ldkv, CH, 128, 0, &tm);
    if (e != cudaSuccess) return e;
    if ((e = tma_map_2d_bf16(map_base, map_rows, map_cols, a.ldkv, 4, 128, 0, &tm4)) != cudaSuccess) return e;
    constexpr int WPH = 2;
    const size_t smem = (size_t)NS * STAGE + 128 + 2 * NS * 8 + 2 * 2 * WPH * 8 * 10 * 4 + 64;
    if (!g_smem_set) {
        if ((e = cudaFuncSetAttribute(attn_decode_tma_kernel<true, WPH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(attn_decode_tma_kernel<false, WPH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        g_smem_set = 1;
    }
    Args k{};
    k.q = (const bf16*)a.q; k.ldq = a.ldq; k.knew = (const bf16*)a.knew; k.vnew = (const bf16*)a.vnew; k.ldnew = a.ldnew;
    k.cache = (bf16*)a.kcache; k.tcap = tcap; k.k_off = a.k_off; k.step = a.step; k.o = (bf16*)a.o; k.ldo = a.ldo;
    k.batch = a.batch; k.col0 = col0;
    const int units = a.batch * 4;
    const int grid = units < max_ctas ? units : max_ctas;
    const dim3 block(32 * (1 + 2 * WPH));
    if (a.knew) return launch_pdl(attn_decode_tma_kernel<true, WPH>, dim3(grid), block, smem, st, tm, tm4, k);
    return launch_pdl(attn_decode_tma_kernel<false, WPH>, dim3(grid), block, smem, st, tm, tm4, k);
}
