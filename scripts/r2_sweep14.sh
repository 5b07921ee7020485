#!/bin/bash
cd "$(dirname "$0")/.."
export CUDA_DEVICE_MAX_CONNECTIONS=32
out=gpurun_out/r2_sweep14.log
: > $out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15 >> $out
for opt in "gemm_bn256=1" "gemm_bn256=0"; do
  timeout 200 python scripts/encoder_profile.py 512 "$opt" >> $out 2>&1 || echo "FAILED $opt" >> $out
done
B="decode_branches=1"
for opt in "$B" "$B,gemm_bn256=0" "$B,dbg_skip=3" "$B,dbg_skip=12" "$B,dbg_skip=15"; do
  timeout 300 python scripts/inflight_probe.py 512 256 6 6 "$opt" >> $out 2>&1 || echo "FAILED $opt" >> $out
done
timeout 300 python scripts/inflight_probe.py 512 256 5,8 6 "$B" >> $out 2>&1
cat $out
