#!/bin/bash
cd "$(dirname "$0")/.."
export CUDA_DEVICE_MAX_CONNECTIONS=32
out=gpurun_out/r2_sweep6.log
: > $out
for v in 1 2 3; do
  echo "== TEXOCR_ABS_SKIP=$v" >> $out
  TEXOCR_B200_LIB=$PWD/texocr_b200/libtexocr_b200_k$v.so timeout 300 python scripts/slot_probe.py 512 256 6 4 >> $out 2>&1 || echo FAILED >> $out
done
cat $out
