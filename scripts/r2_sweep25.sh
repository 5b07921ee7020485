#!/bin/bash
cd "$(dirname "$0")/.."
out=gpurun_out/r2_sweep25.log
: > $out
timeout 600 python -m pytest tests/test_gpu_parity_full.py -x -q -m gpu -k "tiny or stem" 2>&1 | tail -15 >> $out
cat $out
