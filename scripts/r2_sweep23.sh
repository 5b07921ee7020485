#!/bin/bash
cd "$(dirname "$0")/.."
export CUDA_DEVICE_MAX_CONNECTIONS=32
out=gpurun_out/r2_sweep23.log
: > $out
for b in 3 4 6 8; do
  timeout 200 python scripts/inflight_probe.py 512 256 1 6 "decode_branches=$b" >> $out 2>&1
done
timeout 200 python scripts/inflight_probe.py 512 256 1 6 "decode_branches=6,pdl_mid=0" >> $out 2>&1
timeout 200 python scripts/inflight_probe.py 512 256 2 6 "decode_branches=3" >> $out 2>&1
cat $out
