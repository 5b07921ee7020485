#!/bin/bash
cd "$(dirname "$0")/.."
export CUDA_DEVICE_MAX_CONNECTIONS=32
out=gpurun_out/r2_sweep13.log
: > $out
B="decode_branches=1"
for opt in "$B" "$B,pdl_mid=1" "$B,pdl_mid=2" "$B,pdl_mid=4" "$B,pdl_mid=3" "$B,pdl_mid=7" "$B"; do
  timeout 300 python scripts/inflight_probe.py 512 256 6 6 "$opt" >> $out 2>&1 || echo "FAILED $opt" >> $out
done
cat $out
