#!/bin/bash
cd "$(dirname "$0")/.."
export CUDA_DEVICE_MAX_CONNECTIONS=32
out=gpurun_out/r2_sweep5.log
: > $out
timeout 300 python scripts/slot_probe.py 512 256 6 4 >> $out 2>&1 || echo FAILED >> $out
timeout 300 python scripts/slot_probe.py 512 256 1 4 >> $out 2>&1 || echo FAILED >> $out
timeout 300 python scripts/slot_probe.py 512 256 3 4 >> $out 2>&1 || echo FAILED >> $out
timeout 300 python scripts/slot_probe.py 512 256 6 4 pdl=0 >> $out 2>&1 || echo FAILED >> $out
timeout 300 python scripts/slot_probe.py 512 256 6 4 dbg_skip=3 >> $out 2>&1 || echo FAILED >> $out
timeout 300 python scripts/slot_probe.py 512 256 6 4 dbg_skip=12 >> $out 2>&1 || echo FAILED >> $out
cat $out
