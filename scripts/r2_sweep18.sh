#!/bin/bash
cd "$(dirname "$0")/.."
export CUDA_DEVICE_MAX_CONNECTIONS=32
out=gpurun_out/r2_sweep18.log
: > $out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -6 >> $out
timeout 200 python scripts/encoder_profile.py 512 >> $out 2>&1
timeout 300 python scripts/inflight_probe.py 512 256 6 6 "decode_branches=1" >> $out 2>&1
cat $out
