// Encoder: ragged batch geometry, ResNetV2 backbone (fp32 FFMA tier / bf16x3 tcgen05 tier) [model/resnet.py:141-149,251-254],
// patch projection + cls / positional assembly [model/encoder.py:128-143], ViT blocks with the shared double LayerNorm
// [model/attention.py:237-259], and the once-per-image cross-attention memory preparation.
#include "engine_internal.h"

// ------------------------------------------------------------------------------------------------ encoder
int plan_geometry(texocr_handle* h, const int32_t* hw, int B, EncGeom& g, cudaStream_t st) {
    if (B <= 0) return fail(h, TEXOCR_ERR_ARG, "batch must be positive");
    g.B = B;
    g.img_off.assign(B + 1, 0); g.tok_off.assign(B + 1, 0);
    g.uni_h = hw[0]; g.uni_w = hw[1];
    for (int b = 0; b < B; ++b) {
        const int H = hw[2 * b], W = hw[2 * b + 1];
        if (H != g.uni_h || W != g.uni_w) g.uni_h = g.uni_w = 0;
        if (H <= 0 || W <= 0 || H % 16 || W % 16 || H > 160 || W > 1008)
            return fail(h, TEXOCR_ERR_ARG, "image %d is %dx%d: height and width must be multiples of 16 with H <= 160 and "
                        "W <= 1008 (10x63 position grid, model/encoder.py:137-143)", b, H, W);
        const long next = (long)g.img_off[b] + (long)H * W;
        if (next > 0x7fffffffL / 64) return fail(h, TEXOCR_ERR_ARG, "batch too large: more than 2^31 stem activations");
        g.img_off[b + 1] = (int)next;
        for (int l = 0; l <= 4; ++l)
            if ((((long)(H >> l) * (W >> l)) & 31) != 0) g.rows32[l] = false;
        const int n = (H / 16) * (W / 16) + 1;
        g.tok_off[b + 1] = g.tok_off[b] + n;
        g.max_tok = std::max(g.max_tok, n);
    }
    for (int l = 0; l <= 4; ++l) g.P[l] = g.img_off[B] >> (2 * l);
    g.ntok = g.tok_off[B];
    std::vector<int> v;
    v.insert(v.end(), g.img_off.begin(), g.img_off.end());
    v.insert(v.end(), hw, hw + 2 * B);
    v.insert(v.end(), g.tok_off.begin(), g.tok_off.end());
    int r = upload_ints(h, v, st);
    if (r) return r;
    g.d_img_off = h->geom.as<int>();
    g.d_img_hw = g.d_img_off + (B + 1);
    g.d_tok_off = g.d_img_hw + 2 * B;
    return 0;
}

// GroupNorm work is split into per-image pixel chunks.  The grid always offers GN_MAX_CHUNKS chunks per image and each
// image uses min(GN_MAX_CHUNKS, ceil(pixels/64)) of them, a function of its OWN size only: the summation order of
// an image's statistics -- hence every bit of its result -- does not depend on what else is in the batch.
static int nchunk_for(long, int) { return 32; }

static int run_backbone(texocr_handle* h, const float* d_img, const EncGeom& g, cudaStream_t st, const float** feat_out) {
    const int B = g.B;
    ENSURE(h->raw1, (size_t)g.P[1] * 64 * 4);
    ENSURE(h->act2, (size_t)g.P[2] * 64 * 4);
    ENSURE(h->actA, (size_t)g.P[2] * 256 * 4);
    ENSURE(h->actB, (size_t)g.P[2] * 256 * 4);
    ENSURE(h->rawMid, (size_t)g.P[2] * 128 * 4);
    ENSURE(h->actMid, (size_t)g.P[2] * 128 * 4);
    ENSURE(h->rawMid2, (size_t)g.P[2] * 64 * 4);
    ENSURE(h->actMid2, (size_t)g.P[2] * 64 * 4);
    ENSURE(h->raw3, (size_t)g.P[2] * 256 * 4);
    ENSURE(h->rawDs, (size_t)g.P[2] * 256 * 4);
    ENSURE(h->gn_partial, (size_t)B * 32 * 32 * 2 * 8);
    for (int i = 0; i < 4; ++i) ENSURE(h->gn_stats[i], (size_t)B * 32 * 2 * 4);
    float* stats[4] = {h->gn_stats[0].as<float>(), h->gn_stats[1].as<float>(), h->gn_stats[2].as<float>(), h->gn_stats[3].as<float>()};
    double* partial = h->gn_partial.as<double>();

    // stem: conv 7x7/s2 -> GN+ReLU -> maxpool 3x3/s2  [model/resnet.py:218-222]
    LAUNCH(KC_STEM, 1, (double)g.P[0] * 4 + (double)g.P[1] * 64 * 4, 2.0 * 49 * 64 * g.P[1],
           launch_stem_conv(d_img, h->stem_w, h->raw1.as<float>(), g.d_img_off, g.d_img_hw, B, (int)g.P[1], st));
    int nc = nchunk_for(g.P[1], B);
    LAUNCH(KC_GN_STATS, 2, (double)g.P[1] * 64 * 4, 0.0,
           launch_gn_stats(h->raw1.as<float>(), 64, 1, g.d_img_off, B, nc, partial, stats[0], st));
    LAUNCH(KC_GN_APPLY, 1, (double)g.P[1] * 64 * 4 + (double)g.P[2] * 64 * 4, 0.0,
           launch_gn_apply_maxpool(h->raw1.as<float>(), stats[0], h->stem_g, h->stem_b, h->act2.as<float>(), nullptr, nullptr,
                                   g.d_img_off, g.d_img_hw, B, (int)g.P[2], st));

    const float* x = h->act2.as<float>();
    int Cx = 64, Lx = 2;
    float* pingpong[2] = {h->actA.as<float>(), h->actB.as<float>()};
    int pp = 0;
    size_t ci = 0;
    const int depths[3] = {2, 4, 6};
    auto conv = [&](const ConvW& cw, const float* in, int lin, int lout, float* out) -> int {
        const long M = g.P[lout];
        GemmArgs ga = mk_gemm(in, cw.cin, cw.w, cw.k * cw.k * cw.cin, out, cw.cout, (int)M, cw.cout, cw.k * cw.k * cw.cin,
                              EPI_STORE, DT_F32, DT_F32, nullptr, nullptr, 0);
        ConvGather cg{g.d_img_off, g.d_img_hw, B, lin, lout, cw.k, cw.stride, (cw.k == 3 && cw.stride == 1) ? 1 : 0, cw.cin};
        if (!(cw.k == 1 && cw.stride == 1)) ga.conv = &cg;       // 1x1/s1 is a plain GEMM over the pixel rows
        LAUNCH(KC_CONV, 1, gemm_bytes(ga, 4), gemm_flops(ga), launch_gemm_simt(ga, st));
        return 0;
    };
    auto gstats = [&](const float* raw, int C, int level, float* stt) -> int {
        LAUNCH(KC_GN_STATS, 2, (double)g.P[level] * C * 4, 0.0,
               launch_gn_stats(raw, C, level, g.d_img_off, B, nchunk_for(g.P[level], B), partial, stt, st));
        return 0;
    };
    int r;
    for (int s = 0; s < 3; ++s)
        for (int b = 0; b < depths[s]; ++b) {
            const ConvW* ds = nullptr;
            if (b == 0) ds = &h->convs[ci++];
            const ConvW& c1 = h->convs[ci++];
            const ConvW& c2 = h->convs[ci++];
            const ConvW& c3 = h->convs[ci++];
            const int Lout = Lx + (c2.stride == 2 ? 1 : 0);
            if (ds) {
                if ((r = conv(*ds, x, Lx, Lout, h->rawDs.as<float>()))) return r;
                if ((r = gstats(h->rawDs.as<float>(), ds->cout, Lout, stats[3]))) return r;
            }
            if ((r = conv(c1, x, Lx, Lx, h->rawMid.as<float>()))) return r;
            if ((r = gstats(h->rawMid.as<float>(), c1.cout, Lx, stats[0]))) return r;
            {
                GnApplyArgs a{};
                a.raw = h->rawMid.as<float>(); a.stats = stats[0]; a.gamma = c1.gamma; a.beta = c1.beta;
                a.out = h->actMid.as<float>(); a.C = c1.cout; a.level = Lx; a.relu = 1;
                LAUNCH(KC_GN_APPLY, 1, (double)g.P[Lx] * a.C * 8, 0.0, launch_gn_apply(a, g.d_img_off, B, nchunk_for(g.P[Lx], B), st));
            }
            if ((r = conv(c2, h->actMid.as<float>(), Lx, Lout, h->rawMid2.as<float>()))) return r;
            if ((r = gstats(h->rawMid2.as<float>(), c2.cout, Lout, stats[1]))) return r;
            {
                GnApplyArgs a{};
                a.raw = h->rawMid2.as<float>(); a.stats = stats[1]; a.gamma = c2.gamma; a.beta = c2.beta;
                a.out = h->actMid2.as<float>(); a.C = c2.cout; a.level = Lout; a.relu = 1;
                LAUNCH(KC_GN_APPLY, 1, (double)g.P[Lout] * a.C * 8, 0.0, launch_gn_apply(a, g.d_img_off, B, nchunk_for(g.P[Lout], B), st));
            }
            if ((r = conv(c3, h->actMid2.as<float>(), Lout, Lout, h->raw3.as<float>()))) return r;
            if ((r = gstats(h->raw3.as<float>(), c3.cout, Lout, stats[2]))) return r;
            {   // out = ReLU(GN3(y) + res)   [model/resnet.py:147-148]
                GnApplyArgs a{};
                a.raw = h->raw3.as<float>(); a.stats = stats[2]; a.gamma = c3.gamma; a.beta = c3.beta;
                if (ds) { a.raw2 = h->rawDs.as<float>(); a.stats2 = stats[3]; a.gamma2 = ds->gamma; a.beta2 = ds->beta; }
                else a.res = x;
                a.out = pingpong[pp]; a.C = c3.cout; a.level = Lout; a.relu = 1;
                LAUNCH(KC_GN_APPLY, 1, (double)g.P[Lout] * a.C * 12, 0.0, launch_gn_apply(a, g.d_img_off, B, nchunk_for(g.P[Lout], B), st));
            }
            x = pingpong[pp]; pp ^= 1; Cx = c3.cout; Lx = Lout;
        }
    (void)Cx;
    h->last_backbone_pixels = (int)g.P[4];
    *feat_out = x;
    return 0;
}

// bf16 tier: the same backbone with every convolution on the tensor cores in bf16x3 mode (tc_gemm.cu, SPLIT=3).
// Activations live as split-bf16 pairs (hi | lo halves of one buffer); 1x1/s1 convs are plain GEMMs over pixel rows,
// 3x3 and stride-2 convs go through an explicit im2col of the pair.  Raw conv outputs and all statistics stay fp32.
static int run_backbone_tc(texocr_handle* h, const float* d_img, const EncGeom& g, cudaStream_t st, const void** feat_hi,
                           const void** feat_lo) {
    const int B = g.B;
    ENSURE(h->raw1, (size_t)g.P[1] * 64 * 4);
    ENSURE(h->act2, (size_t)g.P[2] * 64 * 4);
    ENSURE(h->actA, (size_t)g.P[2] * 256 * 4);
    ENSURE(h->actB, (size_t)g.P[2] * 256 * 4);
    ENSURE(h->rawMid, (size_t)g.P[2] * 128 * 4);
    ENSURE(h->actMid, (size_t)g.P[2] * 128 * 4);
    ENSURE(h->rawMid2, (size_t)g.P[2] * 64 * 4);
    ENSURE(h->actMid2, (size_t)g.P[2] * 64 * 4);
    ENSURE(h->raw3, (size_t)g.P[2] * 256 * 4);
    ENSURE(h->rawDs, (size_t)g.P[2] * 256 * 4);
    ENSURE(h->col, (size_t)g.P[2] * 576 * 4);
    ENSURE(h->gn_partial, (size_t)B * 32 * 32 * 2 * 8);
    ENSURE(h->gn_part, (size_t)((g.P[1] >> 5) + B + 1) * 64 * 4);
    for (int i = 0; i < 4; ++i) ENSURE(h->gn_stats[i], (size_t)B * 32 * 2 * 4);
    float* stats[4] = {h->gn_stats[0].as<float>(), h->gn_stats[1].as<float>(), h->gn_stats[2].as<float>(), h->gn_stats[3].as<float>()};
    double* partial = h->gn_partial.as<double>();
    struct Pair { char* hi; char* lo; };
    auto pair_of = [](DevBuf& b, size_t elems) { Pair p; p.hi = (char*)b.p; p.lo = (char*)b.p + elems * 2; return p; };

    if (h->use_stem_tc && h->stem_w_hi) {      // implicit GEMM on the tensor cores, GroupNorm partials from its epilogue
        LAUNCH(KC_STEM, 1, (double)g.P[0] * 4 + (double)g.P[1] * 64 * 4, 2.0 * 49 * 64 * g.P[1],
               launch_stem_tc(d_img, h->stem_w_hi, h->stem_w_lo, h->raw1.as<float>(), g.d_img_off, g.d_img_hw, B, g.P[1],
                              g.uni_h > 0 ? (g.uni_h >> 1) * (g.uni_w >> 1) : 0, h->gn_part.as<float>(), st));
        LAUNCH(KC_GN_STATS, 1, (double)((g.P[1] >> 5) + B) * 256, 0.0,
               launch_gn_finalize_blocks(h->gn_part.as<float>(), 64, 1, g.d_img_off, B, stats[0], st));
    } else {
        LAUNCH(KC_STEM, 1, (double)g.P[0] * 4 + (double)g.P[1] * 64 * 4, 2.0 * 49 * 64 * g.P[1],
               launch_stem_conv(d_img, h->stem_w, h->raw1.as<float>(), g.d_img_off, g.d_img_hw, B, (int)g.P[1], st));
        LAUNCH(KC_GN_STATS, 2, (double)g.P[1] * 64 * 4, 0.0,
               launch_gn_stats(h->raw1.as<float>(), 64, 1, g.d_img_off, B, nchunk_for(g.P[1], B), partial, stats[0], st));
    }
    Pair x = pair_of(h->act2, (size_t)g.P[2] * 64);
    LAUNCH(KC_GN_APPLY, 1, (double)g.P[1] * 64 * 4 + (double)g.P[2] * 64 * 4, 0.0,
           launch_gn_apply_maxpool(h->raw1.as<float>(), stats[0], h->stem_g, h->stem_b, nullptr, x.hi, x.lo, g.d_img_off,
                                   g.d_img_hw, B, (int)g.P[2], st));
    int Lx = 2;
    DevBuf* pingpong[2] = {&h->actA, &h->actB};
    int pp = 0;
    size_t ci = 0;
    const int depths[3] = {2, 4, 6};
    // conv: split GEMM, through im2col unless 1x1/s1
    // GroupNorm statistics of a convolution's fp32 output: block partial sums (gn_block.cuh) left by the GEMM epilogue when
    // every image of the batch has a multiple of 32 pixel rows at the output level, by the stand-alone kernel otherwise
    // (identical bits either way), then one small kernel that adds an image's blocks in order.
    auto fused_ok = [&](int level) { return h->gn_fused && g.rows32[level]; };
    auto conv = [&](const ConvW& cw, Pair in, int lin, int lout, float* out, float* stt) -> int {
        const long M = g.P[lout];
        const int K = cw.k * cw.k * cw.cin;
        const void *a_hi = in.hi, *a_lo = in.lo;
        const int pad_lo = (cw.k == 3 && cw.stride == 1) ? 1 : 0;
        // same-size batch: the GEMM fetches its A tiles straight from the NHWC activation with TMA im2col loads
        const bool implicit = h->use_im2col_tma && g.uni_h > 0 && cw.cin % 64 == 0 && !(cw.k == 1 && cw.stride == 1);
        // ragged batch: the GEMM's producer warps gather the A tiles from the NHWC activation (tc_conv_gather_kernel)
        const bool gather = !implicit && h->use_conv_gather && cw.cin % 64 == 0 && !(cw.k == 1 && cw.stride == 1);
        ConvGather cg{g.d_img_off, g.d_img_hw, B, lin, lout, cw.k, cw.stride, pad_lo, cw.cin};
        if (!(cw.k == 1 && cw.stride == 1) && !implicit && !gather) {
            Pair c = pair_of(h->col, (size_t)M * K);
            LAUNCH(KC_GN_APPLY, 1, (double)M * K * 8, 0.0, launch_im2col_split(in.hi, in.lo, c.hi, c.lo, cg, M, st));
            a_hi = c.hi; a_lo = c.lo;
        }
        GemmArgs ga = mk_gemm(a_hi, K, cw.w_hi, K, out, cw.cout, (int)M, cw.cout, K, EPI_STORE, DT_BF16, DT_F32, nullptr, nullptr, 0);
        ga.A2 = a_lo; ga.W2 = cw.w_lo;
        if (implicit) {
            ga.lda = cw.cin;
            ga.im2col = {cw.k, cw.stride, pad_lo, cw.cin, g.uni_w >> lin, g.uni_h >> lin, B, g.uni_w >> lout, g.uni_h >> lout};
        }
        if (gather) { ga.lda = cw.cin; ga.gather = &cg; }
        const bool fused = fused_ok(lout);
        if (fused) {
            ga.gn_part = h->gn_part.as<float>();
            if (g.uni_h > 0) ga.gn_rpi = (g.uni_h >> lout) * (g.uni_w >> lout);
            else { ga.gn_rpi = 0; ga.gn_img_off = g.d_img_off; ga.gn_nimg = B; ga.gn_level = lout; }
        }
        if (!tc_gemm_supported(ga)) return fail(h, TEXOCR_ERR_ARG, "backbone conv %s not supported by the tcgen05 GEMM", cw.name.c_str());
        LAUNCH(KC_CONV, 1, gemm_bytes(ga, 4), gemm_flops(ga), launch_gemm_tc(ga, st));
        if (!fused)
            LAUNCH(KC_GN_STATS, 1, (double)M * cw.cout * 4, 0.0,
                   launch_gn_stats_blocks(out, cw.cout, lout, g.d_img_off, B, M, h->gn_part.as<float>(), st));
        LAUNCH(KC_GN_STATS, 1, (double)((M >> 5) + B) * 256, 0.0,
               launch_gn_finalize_blocks(h->gn_part.as<float>(), cw.cout, lout, g.d_img_off, B, stt, st));
        return 0;
    };
    auto gapply = [&](GnApplyArgs& a, Pair out, int level, double bytes_per) -> int {
        a.out = nullptr; a.out_hi = out.hi; a.out_lo = out.lo; a.level = level;
        LAUNCH(KC_GN_APPLY, 1, (double)g.P[level] * a.C * bytes_per, 0.0, launch_gn_apply(a, g.d_img_off, B, nchunk_for(g.P[level], B), st));
        return 0;
    };
    int r;
    for (int s = 0; s < 3; ++s)
        for (int b = 0; b < depths[s]; ++b) {
            const ConvW* ds = nullptr;
            if (b == 0) ds = &h->convs[ci++];
            const ConvW& c1 = h->convs[ci++];
            const ConvW& c2 = h->convs[ci++];
            const ConvW& c3 = h->convs[ci++];
            const int Lout = Lx + (c2.stride == 2 ? 1 : 0);
            if (ds) {
                if ((r = conv(*ds, x, Lx, Lout, h->rawDs.as<float>(), stats[3]))) return r;
            }
            if ((r = conv(c1, x, Lx, Lx, h->rawMid.as<float>(), stats[0]))) return r;
            Pair m1 = pair_of(h->actMid, (size_t)g.P[Lx] * c1.cout);
            {
                GnApplyArgs a{};
                a.raw = h->rawMid.as<float>(); a.stats = stats[0]; a.gamma = c1.gamma; a.beta = c1.beta; a.C = c1.cout; a.relu = 1;
                if ((r = gapply(a, m1, Lx, 8))) return r;
            }
            if ((r = conv(c2, m1, Lx, Lout, h->rawMid2.as<float>(), stats[1]))) return r;
            Pair m2 = pair_of(h->actMid2, (size_t)g.P[Lout] * c2.cout);
            {
                GnApplyArgs a{};
                a.raw = h->rawMid2.as<float>(); a.stats = stats[1]; a.gamma = c2.gamma; a.beta = c2.beta; a.C = c2.cout; a.relu = 1;
                if ((r = gapply(a, m2, Lout, 8))) return r;
            }
            if ((r = conv(c3, m2, Lout, Lout, h->raw3.as<float>(), stats[2]))) return r;
            Pair out = pair_of(*pingpong[pp], (size_t)g.P[Lout] * c3.cout);
            {
                GnApplyArgs a{};
                a.raw = h->raw3.as<float>(); a.stats = stats[2]; a.gamma = c3.gamma; a.beta = c3.beta; a.C = c3.cout; a.relu = 1;
                if (ds) { a.raw2 = h->rawDs.as<float>(); a.stats2 = stats[3]; a.gamma2 = ds->gamma; a.beta2 = ds->beta; }
                else { a.res_hi = x.hi; a.res_lo = x.lo; }
                if ((r = gapply(a, out, Lout, 12))) return r;
            }
            x = out; pp ^= 1; Lx = Lout;
        }
    h->last_backbone_pixels = 0;     // the fp32 tap is only kept by the fp32 tier
    *feat_hi = x.hi; *feat_lo = x.lo;
    return 0;
}

int sub_attn_out(texocr_handle* h, const RowCtx& rc, const AttnW& w, cudaStream_t st) {
    // y = o.Wo^T + bo -> GLU -> + residual   [model/attention.py:96-99,180 ; 254]
    GemmArgs ga = mk_gemm(rowa(h, h->o, rc, 512), 512, w.wo, 512, rowf(h->s, rc, 256), 256, rc.rows, 512, 512, EPI_GLU_RES, h->dt, DT_F32, w.bo, rowf(h->x, rc, 256), 256);
    LAUNCH(rc.kc_gemm == KC_DEC_GEMM ? KC_DEC_GEMM_WO : rc.kc_gemm, 1, gemm_bytes(ga, h->esz), gemm_flops(ga), run_gemm(h, ga, st));
    return 0;
}
int sub_mlp(texocr_handle* h, const RowCtx& rc, const MlpW& w, cudaStream_t st) {
    GemmArgs g1 = mk_gemm(rowa(h, h->xn, rc, 256), 256, w.w1, 256, rowa(h, h->hid, rc, 1024), 1024, rc.rows, 2048, 256, EPI_GEGLU, h->dt, h->dt, w.b1, nullptr, 0);
    LAUNCH(rc.kc_gemm == KC_DEC_GEMM ? KC_DEC_GEMM_W1 : rc.kc_gemm, 1, gemm_bytes(g1, h->esz), gemm_flops(g1), run_gemm(h, g1, st));
    GemmArgs g2 = mk_gemm(rowa(h, h->hid, rc, 1024), 1024, w.w2, 1024, rowf(h->s, rc, 256), 256, rc.rows, 256, 1024, EPI_BIAS_RES, h->dt, DT_F32, w.b2, rowf(h->x, rc, 256), 256);
    LAUNCH(rc.kc_gemm == KC_DEC_GEMM ? KC_DEC_GEMM_W2 : rc.kc_gemm, 1, gemm_bytes(g2, h->esz), gemm_flops(g2), run_gemm(h, g2, st));
    return 0;
}
// x = LN(s); xn = LN(x)  (shared LayerNorm twice, model/attention.py:242-259), or the stack's final norm.
int sub_norm(texocr_handle* h, const RowCtx& rc, bool last, const float* fin_g, const float* fin_b, float* fin_out_f,
                    void* fin_out_a, cudaStream_t st) {
    if ((h->dbg_skip & 4) && rc.kc_row == KC_DEC_ROW) return 0;
    Ln2Args a{};
    a.in = rowf(h->s, rc, 256); a.rows = rc.rows; a.dt_a = h->dt;
    if (!last) { a.g1 = rc.ln_g; a.b1 = rc.ln_b; a.g2 = rc.ln_g; a.b2 = rc.ln_b; a.o1f = rowf(h->x, rc, 256); a.o2a = rowa(h, h->xn, rc, 256); }
    else { a.g1 = fin_g; a.b1 = fin_b; a.o1f = fin_out_f; a.o1a = fin_out_a; }
    LAUNCH(rc.kc_row, 1, (double)rc.rows * 256 * (4 + 4 + h->esz), 0.0, launch_ln2(a, st));
    return 0;
}

int ensure_rows(texocr_handle* h, long rows) {
    ENSURE(h->x, (size_t)rows * 256 * 4);
    ENSURE(h->s, (size_t)rows * 256 * 4);
    ENSURE(h->xn, (size_t)rows * 256 * h->esz);
    ENSURE(h->qkv, (size_t)rows * 1536 * h->esz);
    ENSURE(h->o, (size_t)rows * 512 * h->esz);
    ENSURE(h->hid, (size_t)rows * 1024 * h->esz);
    return 0;
}

// images (device) -> h->enc_out (fp32) and h->enc_a (GEMM operand type)
int run_encoder(texocr_handle* h, const float* d_img, const EncGeom& g, cudaStream_t st) {
    const texocr_config& c = h->cfg;
    int r;
    ENSURE(h->proj_out, (size_t)g.P[4] * 256 * 4);
    if (c.encoder_kind == TEXOCR_ENC_HYBRID && h->dt == DT_BF16 && h->use_tcgen05) {
        const void *fh = nullptr, *fl = nullptr;
        if ((r = run_backbone_tc(h, d_img, g, st, &fh, &fl))) return r;
        GemmArgs ga = mk_gemm(fh, 1024, h->proj_w, 1024, h->proj_out.p, 256, (int)g.P[4], 256, 1024, EPI_STORE, DT_BF16, DT_F32, h->proj_b, nullptr, 0);
        ga.A2 = fl; ga.W2 = h->proj_w_lo;
        LAUNCH(KC_ENC_GEMM, 1, gemm_bytes(ga, 4), gemm_flops(ga), launch_gemm_tc(ga, st));
    } else if (c.encoder_kind == TEXOCR_ENC_HYBRID) {
        const float* feat = nullptr;
        if ((r = run_backbone(h, d_img, g, st, &feat))) return r;
        const void* a_ptr = feat;
        if (h->dt != DT_F32) {
            ENSURE(h->backbone_a, (size_t)g.P[4] * 1024 * h->esz);
            LAUNCH(KC_MISC, 1, (double)g.P[4] * 1024 * 6, 0.0, launch_cast_f32_to(feat, h->backbone_a.p, g.P[4] * 1024, h->dt, st));
            a_ptr = h->backbone_a.p;
        }
        GemmArgs ga = mk_gemm(a_ptr, 1024, h->proj_w, 1024, h->proj_out.p, 256, (int)g.P[4], 256, 1024, EPI_STORE, h->dt, DT_F32, h->proj_b, nullptr, 0);
        LAUNCH(KC_ENC_GEMM, 1, gemm_bytes(ga, h->esz), gemm_flops(ga), run_gemm(h, ga, st));
    } else {
        ENSURE(h->patch_cols, (size_t)g.P[4] * 256 * 4);
        LAUNCH(KC_MISC, 1, (double)g.P[0] * 8, 0.0, launch_im2col_patch(d_img, h->patch_cols.as<float>(), g.d_img_off, g.d_img_hw, g.B, (int)g.P[4], st));
        const void* a_ptr = h->patch_cols.p;
        if (h->dt != DT_F32) {
            ENSURE(h->backbone_a, (size_t)g.P[4] * 256 * h->esz);
            LAUNCH(KC_MISC, 1, (double)g.P[4] * 256 * 6, 0.0, launch_cast_f32_to(h->patch_cols.as<float>(), h->backbone_a.p, g.P[4] * 256, h->dt, st));
            a_ptr = h->backbone_a.p;
        }
        GemmArgs ga = mk_gemm(a_ptr, 256, h->proj_w, 256, h->proj_out.p, 256, (int)g.P[4], 256, 256, EPI_STORE, h->dt, DT_F32, h->proj_b, nullptr, 0);
        LAUNCH(KC_ENC_GEMM, 1, gemm_bytes(ga, h->esz), gemm_flops(ga), run_gemm(h, ga, st));
    }
    const int R = g.ntok;
    if ((r = ensure_rows(h, R))) return r;
    ENSURE(h->enc_out, (size_t)R * 256 * 4);
    ENSURE(h->enc_a, (size_t)R * 256 * h->esz);
    LAUNCH(KC_ENC_ROW, 1, (double)R * 256 * 12, 0.0,
           launch_assemble_tokens(h->proj_out.as<float>(), h->cls, h->pos, h->x.as<float>(), g.d_img_off, g.d_img_hw, g.d_tok_off, g.B, R, st));
    RowCtx rc{R, KC_ENC_GEMM, KC_ENC_ROW, h->enc_ln_g, h->enc_ln_b};
    {
        Ln2Args a{};
        a.in = h->x.as<float>(); a.rows = R; a.g2 = h->enc_ln_g; a.b2 = h->enc_ln_b; a.o2a = h->xn.p; a.dt_a = h->dt;
        LAUNCH(KC_ENC_ROW, 1, (double)R * 256 * (4 + h->esz), 0.0, launch_ln2(a, st));
    }
    for (int l = 0; l < c.enc_layers; ++l) {
        GemmArgs gq = mk_gemm(h->xn.p, 256, h->enc_attn[l].wqkv, 256, h->qkv.p, 1536, R, 1536, 256, EPI_STORE, h->dt, h->dt, nullptr, nullptr, 0);
        LAUNCH(KC_ENC_GEMM, 1, gemm_bytes(gq, h->esz), gemm_flops(gq), run_gemm(h, gq, st));
        AttnVarlenArgs av{};
        const char* base = (const char*)h->qkv.p;
        av.q = base; av.k = base + 512 * h->esz; av.v = base + 1024 * h->esz; av.ldq = av.ldk = av.ldv = 1536;
        av.o = h->o.p; av.ldo = 512; av.q_off = g.d_tok_off; av.k_off = g.d_tok_off; av.batch = g.B; av.max_q = g.max_tok;
        av.causal = 0; av.dt = h->dt;
        double aflops = 0.0;
        for (int b = 0; b < g.B; ++b) { const double n = g.tok_off[b + 1] - g.tok_off[b]; aflops += 4.0 * n * n * 512; }
        if (h->use_tcgen05 && attn_enc_mma_supported(av)) LAUNCH(KC_ENC_ATTN, 1, (double)R * 2048 * h->esz, aflops, launch_attn_enc_mma(av, st));
        else LAUNCH(KC_ENC_ATTN, 1, (double)R * 2048 * h->esz, aflops, launch_attn_varlen(av, st));
        if ((r = sub_attn_out(h, rc, h->enc_attn[l], st))) return r;
        if ((r = sub_norm(h, rc, false, nullptr, nullptr, nullptr, nullptr, st))) return r;
        if ((r = sub_mlp(h, rc, h->enc_mlp[l], st))) return r;
        const bool last = (l == c.enc_layers - 1);
        if ((r = sub_norm(h, rc, last, h->enc_norm_g, h->enc_norm_b, h->enc_out.as<float>(), h->dt == DT_F32 ? nullptr : h->enc_a.p, st))) return r;
    }
    return 0;
}

// memory (device fp32 or already-typed copy) -> h->crosskv [ntok, L*1024]   (K/V of every cross-attention layer, once)
// y = out-projection(C) + bo -> GLU -> + residual for the absorbed attention, C = [rows, 8 x 256] softmax-weighted latent averages:
// either one folded GEMM (K = 2048), or the per-head value projection (block-diagonal, K = 256) into `o` followed by the ordinary
// Wo GEMM -- 4x fewer FLOPs / weight bytes and a shorter dependent chain (12 instead of 32 k-blocks)
int sub_abs_out(texocr_handle* h, const RowCtx& rc, const AttnW& w, const void* ca, cudaStream_t st) {
    if (!h->absorb_two_stage) {
        GemmArgs go = mk_gemm(ca, 2048, w.wvo, 2048, rowf(h->s, rc, 256), 256, rc.rows, 512, 2048, EPI_GLU_RES, h->dt, DT_F32, w.bo, rowf(h->x, rc, 256), 256);
        LAUNCH(rc.kc_gemm == KC_DEC_GEMM ? KC_DEC_GEMM_WO : rc.kc_gemm, 1, gemm_bytes(go, h->esz), gemm_flops(go), run_gemm(h, go, st));
        return 0;
    }
    GemmArgs gv = mk_gemm(ca, 2048, w.wv, 256, rowa(h, h->o, rc, 512), 512, rc.rows, 512, 256, EPI_STORE, h->dt, h->dt, nullptr, nullptr, 0);
    gv.a_block_k = 256;
    // algorithmic work of the block-diagonal GEMM: every row contracts 8 heads x (256 -> 64)
    LAUNCH(rc.kc_gemm == KC_DEC_GEMM ? KC_DEC_GEMM_VPROJ : rc.kc_gemm, 1, (double)rc.rows * 2048 * h->esz + 512.0 * 256 * h->esz + (double)rc.rows * 512 * h->esz,
           2.0 * rc.rows * 512.0 * 256, run_gemm(h, gv, st));
    return sub_attn_out(h, rc, w, st);
}

// generate loop, bf16 tier: absorbed cross-attention (needs the TMA attention path and the per-branch kernel graphs)
static bool use_absorb(const texocr_handle* h) {
    return h->dt == DT_BF16 && h->use_tcgen05 && (h->use_tma_attn == 1 || h->use_tma_attn == 3);
}

int run_crosskv(texocr_handle* h, const float* enc_f32, const void* enc_typed, int ntok, cudaStream_t st, bool for_generate) {
    const int L = h->cfg.dec_layers;
    h->self_abs_active = for_generate && h->self_absorb && use_absorb(h);
    if (for_generate && h->cross_absorb && use_absorb(h)) {      // no K/V projection at all: the decode loop streams the bf16 encoder memory itself
        if (!enc_typed) {
            ENSURE(h->enc_a, (size_t)ntok * 256 * h->esz);
            LAUNCH(KC_MISC, 1, (double)ntok * 256 * 6, 0.0, launch_cast_f32_to(enc_f32, h->enc_a.p, (int64_t)ntok * 256, h->dt, st));
            enc_typed = h->enc_a.p;
        }
        h->dec_enc = enc_typed;
        h->crosskv_rows = ntok;
        return 0;
    }
    h->dec_enc = nullptr;
    ENSURE(h->crosskv, (size_t)ntok * L * 1024 * h->esz);
    h->crosskv_rows = ntok;
    const void* a_ptr = enc_f32;
    if (h->dt != DT_F32) {
        if (!enc_typed) {
            ENSURE(h->enc_a, (size_t)ntok * 256 * h->esz);
            LAUNCH(KC_MISC, 1, (double)ntok * 256 * 6, 0.0, launch_cast_f32_to(enc_f32, h->enc_a.p, (int64_t)ntok * 256, h->dt, st));
            enc_typed = h->enc_a.p;
        }
        a_ptr = enc_typed;
    }
    GemmArgs ga = mk_gemm(a_ptr, 256, h->w_crosskv, 256, h->crosskv.p, L * 1024, ntok, L * 1024, 256, EPI_STORE, h->dt, h->dt, nullptr, nullptr, 0);
    LAUNCH(KC_CROSSKV_GEMM, 1, gemm_bytes(ga, h->esz), gemm_flops(ga), run_gemm(h, ga, st));
    return 0;
}
