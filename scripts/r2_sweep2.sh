#!/bin/bash
# round-2 experiment: rows per launch (super-batches) vs batches in flight
cd "$(dirname "$0")/.."
export CUDA_DEVICE_MAX_CONNECTIONS=32
out=gpurun_out/r2_sweep2.log
: > $out
timeout 400 python scripts/inflight_probe.py 1024 256 1,2,3,4 3 "decode_branches=1" >> $out 2>&1 || echo FAILED >> $out
timeout 400 python scripts/inflight_probe.py 2048 256 1,2,3 2 "decode_branches=1" >> $out 2>&1 || echo FAILED >> $out
timeout 400 python scripts/inflight_probe.py 4096 256 1,2 2 "decode_branches=1" >> $out 2>&1 || echo FAILED >> $out
timeout 400 python scripts/inflight_probe.py 2048 256 2 2 "decode_branches=2" >> $out 2>&1 || echo FAILED >> $out
cat $out
