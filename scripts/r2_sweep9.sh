#!/bin/bash
cd "$(dirname "$0")/.."
export CUDA_DEVICE_MAX_CONNECTIONS=32
out=gpurun_out/r2_sweep9.log
: > $out
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x 2>&1 | tail -2 >> $out
for opt in "gemm_ring3=0" "gemm_ring3=1" "gemm_ring3=1,attn_abs_minb=4" "gemm_ring3=1,gemm_min_ctas=240"; do
  timeout 300 python scripts/slot_probe.py 512 256 6 4 "$opt" >> $out 2>&1 || echo "FAILED $opt" >> $out
done
timeout 300 python scripts/slot_probe.py 512 256 8 4 "gemm_ring3=1" >> $out 2>&1 || echo "FAILED" >> $out
cat $out
