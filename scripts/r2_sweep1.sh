#!/bin/bash
# round-2 experiment: which launch-policy options move the 6-batches-in-flight throughput
cd "$(dirname "$0")/.."
export CUDA_DEVICE_MAX_CONNECTIONS=32
out=gpurun_out/r2_sweep1.log
: > $out
for opt in "decode_branches=1" "decode_branches=1,pdl=0" "decode_branches=1,pdl=831" "decode_branches=1,pdl=575" "decode_branches=1,pdl=319" \
           "decode_branches=1,pdl=15" "decode_branches=1,pdl=48" "decode_branches=1,gemm_min_ctas=60" "decode_branches=1,gemm_min_ctas=30" \
           "decode_branches=1,pdl=0,attn_ctas_per_sm=2" "decode_branches=1,attn_ctas_per_sm=2" "decode_branches=1,pdl=0,gemm_min_ctas=30"; do
  timeout 300 python scripts/inflight_probe.py 512 256 6 6 "$opt" >> $out 2>&1 || echo "FAILED $opt" >> $out
done
timeout 300 python scripts/inflight_probe.py 512 256 8,10 6 "decode_branches=1,pdl=0" >> $out 2>&1
cat $out
