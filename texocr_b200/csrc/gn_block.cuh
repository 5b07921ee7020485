// GroupNorm(32) statistics of the bf16 tier as per-block partial sums with ONE summation order (model/resnet.py:32-35 is
// F.group_norm over an image's pixels x the group's channels).
//
// An image's pixel rows are cut into blocks of 32 rows counted from the image's first row; a block x a 32-channel chunk is
// reduced by one warp: lane (rr = lane >> 3, ch = lane & 7) owns channels ch*4 .. ch*4+3 of rows rr, rr+4, .. rr+28 and adds
// them in that order (sum: plain adds, sum of squares: fmaf), then the lane's four channels are folded into its group(s)
// and the lanes are combined with xor butterflies (rows: 8, 16; channels of a group wider than 4: 1, 2, 4).  Both producers
// of these partials -- the epilogue of the tcgen05 convolution GEMM (tc_gemm.cu, which passes the staged output tile through
// exactly this lane mapping on its way to global memory) and the stand-alone kernel for ragged batches (conv_gn.cu) -- call
// the two functions below, so an image's statistics are the same bits whichever path produced them and whatever else is in
// the batch.  partial[slot][32 groups][2] floats, slot = (first pixel row of the block >> 5) + image index (unique per
// (image, block) for any mixture of image sizes); gn_finalize_blocks_kernel adds an image's blocks in order in double.
#pragma once
#include "common.cuh"

TX_DEVINL void gn_block_acc(float* s, float* q, const float4 f) {
    s[0] += f.x; s[1] += f.y; s[2] += f.z; s[3] += f.w;
    q[0] = fmaf(f.x, f.x, q[0]); q[1] = fmaf(f.y, f.y, q[1]); q[2] = fmaf(f.z, f.z, q[2]); q[3] = fmaf(f.w, f.w, q[3]);
}

// s / q: the lane's column sums; col0: channel of the chunk's first column; cpg: channels per group (2, 4, 8, 16 or 32);
// slot_part = partial + slot * 64.  Must be called by all 32 lanes.
TX_DEVINL void gn_block_finish(const float* s, const float* q, int lane, int cpg, int col0, float* slot_part) {
    float a0 = s[0] + s[1], a1 = s[2] + s[3], b0 = q[0] + q[1], b1 = q[2] + q[3];
    if (cpg >= 4) { a0 += a1; b0 += b1; }
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o); b0 += __shfl_xor_sync(0xffffffffu, b0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o); b1 += __shfl_xor_sync(0xffffffffu, b1, o);
    }
    for (int o = 1; o * 4 < cpg; o <<= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o); b0 += __shfl_xor_sync(0xffffffffu, b0, o);
    }
    if (lane < 8) {
        const int c = col0 + lane * 4;
        if (cpg == 2) *reinterpret_cast<float4*>(slot_part + c) = make_float4(a0, b0, a1, b1);      // groups c/2 and c/2 + 1
        else if ((c & (cpg - 1)) == 0) *reinterpret_cast<float2*>(slot_part + (c / cpg) * 2) = make_float2(a0, b0);
    }
}
