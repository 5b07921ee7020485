// tcgen05 / TMEM / TMA GEMM for bf16 operands (sm_100a).  See tc_gemm.cu.
#pragma once
#include <cuda.h>

#include "kernels.h"

// true when launch_gemm_tc can take this problem (shape / alignment / epilogue); the engine falls back to the
// FFMA kernel otherwise.
bool tc_gemm_supported(const GemmArgs& g);
cudaError_t launch_gemm_tc(const GemmArgs& g, cudaStream_t st);
// Shared TMA helper (driver entry point resolved through the runtime; no libcuda link): cached tensor map of a 2-D bf16
// matrix [rows, cols] (row stride ld elements), box = box_rows x box_cols, optional 128-byte swizzle, zero OOB fill.
cudaError_t tma_map_2d_bf16(const void* ptr, long rows, int cols, long ld, int box_rows, int box_cols, int swizzle128,
                            CUtensorMap* out);
// 3-D variant (d0 contiguous, byte strides for d1 / d2), 128-byte swizzle; not cached.
cudaError_t tma_map_3d_bf16(const void* ptr, int d0, long d1, long d2, long stride1_bytes, long stride2_bytes, int b0, int b1, int b2,
                            CUtensorMap* out);
// Stem convolution of the bf16 tier as a tcgen05 implicit GEMM (7x7 / stride 2, one input channel, K = 49 taps padded to 64, bf16x3), with the
// GroupNorm block partials of its output (gn_block.cuh).  w_hi / w_lo: [64][64] bf16.
cudaError_t launch_stem_tc(const float* img, const void* w_hi, const void* w_lo, float* raw1, const int* img_off, const int* img_hw, int nimg,
                           long total_p1, int uniform_rpi /* level-1 pixels per image of a same-size batch, else 0 */, float* gn_part, cudaStream_t st);
