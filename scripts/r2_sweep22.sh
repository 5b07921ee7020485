#!/bin/bash
cd "$(dirname "$0")/.."
export CUDA_DEVICE_MAX_CONNECTIONS=32
out=gpurun_out/r2_sweep22.log
: > $out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py -x -q -m gpu -k "stem or encoder or im2col or groupnorm or bf16_tier" 2>&1 | tail -8 >> $out
timeout 200 python scripts/encoder_profile.py 512 >> $out 2>&1
bash scripts/r2_sweep19.sh > /dev/null 2>&1; grep -A8 "conv_gather=1" gpurun_out/r2_sweep19.log >> $out
timeout 300 python scripts/inflight_probe.py 512 256 6 6 "decode_branches=1" >> $out 2>&1
cat $out
