// Placeholder until the tcgen05 kernel lands: reports "unsupported" so the engine uses the FFMA path.
#include "tc_gemm.h"

bool tc_gemm_supported(const GemmArgs&) { return false; }
cudaError_t launch_gemm_tc(const GemmArgs&, cudaStream_t) { return cudaErrorNotSupported; }
