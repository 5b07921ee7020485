#!/bin/bash
cd "$(dirname "$0")/.."
export CUDA_DEVICE_MAX_CONNECTIONS=32
out=gpurun_out/r2_sweep20.log
: > $out
B="decode_branches=1"
timeout 300 python scripts/inflight_probe.py 512 256 6,10 4 "$B" >> $out 2>&1
for mc in 60 32 16; do
  timeout 300 python scripts/inflight_probe.py 512 256 6,10 4 "$B,gemm_min_ctas=$mc" >> $out 2>&1 || echo "FAILED $mc" >> $out
done
cat $out
