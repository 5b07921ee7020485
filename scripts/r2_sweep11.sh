#!/bin/bash
cd "$(dirname "$0")/.."
export CUDA_DEVICE_MAX_CONNECTIONS=32
out=gpurun_out/r2_sweep11.log
: > $out
B="decode_branches=1"
for opt in "$B" "$B,gemm_min_ctas=60" "$B,gemm_min_ctas=240" "$B,pdl=0" "$B,pdl=831"; do
  timeout 300 python scripts/inflight_probe.py 512 256 6 6 "$opt" >> $out 2>&1 || echo "FAILED $opt" >> $out
done
timeout 300 python scripts/inflight_probe.py 512 256 4,8,10 6 "$B" >> $out 2>&1
timeout 300 python scripts/inflight_probe.py 1024 256 3,4 3 "$B" >> $out 2>&1
timeout 300 python scripts/inflight_probe.py 512 256 6 6 "$B,dbg_skip=3" >> $out 2>&1
timeout 300 python scripts/inflight_probe.py 512 256 6 6 "$B,dbg_skip=12" >> $out 2>&1
timeout 300 python scripts/inflight_probe.py 512 256 6 6 "$B,dbg_skip=15" >> $out 2>&1
cat $out
