#!/bin/bash
cd "$(dirname "$0")/.."
export CUDA_DEVICE_MAX_CONNECTIONS=32
out=gpurun_out/r2_sweep21.log
: > $out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py -x -q -m gpu -k "stem or encoder or im2col or groupnorm or bf16_tier" 2>&1 | tail -12 >> $out
for opt in "stem_tc=1" "stem_tc=0"; do
  timeout 200 python scripts/encoder_profile.py 512 "$opt" >> $out 2>&1 || echo "FAILED $opt" >> $out
done
timeout 300 python scripts/inflight_probe.py 512 256 6 6 "decode_branches=1" >> $out 2>&1
cat $out
