#!/bin/bash
cd "$(dirname "$0")/.."
export CUDA_DEVICE_MAX_CONNECTIONS=32
out=gpurun_out/r2_sweep7.log
: > $out
timeout 600 python -m pytest tests/test_gpu_parity_full.py tests/test_gpu_parity.py -q -m gpu -x -k "groupnorm or encoder or ragged or im2col or bf16_tier" >> $out 2>&1
timeout 300 python scripts/encoder_profile.py >> $out 2>&1 || echo "FAILED encprof" >> $out
for opt in "decode_branches=1,gn_fused=0" "decode_branches=1,gn_fused=1"; do
  timeout 300 python scripts/inflight_probe.py 512 256 6 6 "$opt" >> $out 2>&1 || echo "FAILED $opt" >> $out
done
cat $out
