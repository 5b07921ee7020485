// tcgen05 / TMEM / TMA GEMM for bf16 operands (sm_100a).  See tc_gemm.cu.
#pragma once
#include "kernels.h"

// true when launch_gemm_tc can take this problem (shape / alignment / epilogue); the engine falls back to the
// FFMA kernel otherwise.
bool tc_gemm_supported(const GemmArgs& g);
cudaError_t launch_gemm_tc(const GemmArgs& g, cudaStream_t st);
