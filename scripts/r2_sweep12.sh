#!/bin/bash
cd "$(dirname "$0")/.."
export CUDA_DEVICE_MAX_CONNECTIONS=32
out=gpurun_out/r2_sweep12.log
: > $out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15 >> $out
for opt in "gn_fused=1" "gn_fused=0" "gemm_epi_warps=4" "gemm_epi_warps=4,gn_fused=0"; do
  timeout 200 python scripts/encoder_profile.py 512 "$opt" >> $out 2>&1 || echo "FAILED $opt" >> $out
done
B="decode_branches=1"
for opt in "$B" "$B,gemm_epi_warps=4"; do
  timeout 300 python scripts/inflight_probe.py 512 256 6 6 "$opt" >> $out 2>&1 || echo "FAILED $opt" >> $out
done
cat $out
