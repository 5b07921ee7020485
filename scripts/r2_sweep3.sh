#!/bin/bash
cd "$(dirname "$0")/.."
export CUDA_DEVICE_MAX_CONNECTIONS=32
out=gpurun_out/r2_sweep3.log
: > $out
for opt in "decode_branches=1,attn_l2_policy=0" "decode_branches=1,attn_l2_policy=1" "decode_branches=1,attn_l2_policy=2" "decode_branches=1,attn_l2_policy=3"; do
  timeout 300 python scripts/inflight_probe.py 512 256 6 6 "$opt" >> $out 2>&1 || echo "FAILED $opt" >> $out
done
timeout 300 python scripts/inflight_probe.py 512 256 3,4 6 "decode_branches=1,attn_l2_policy=3" >> $out 2>&1
timeout 300 python scripts/inflight_probe.py 512 256 3,4 6 "decode_branches=1,attn_l2_policy=0" >> $out 2>&1
cat $out
