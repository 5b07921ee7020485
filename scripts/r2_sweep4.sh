#!/bin/bash
cd "$(dirname "$0")/.."
export CUDA_DEVICE_MAX_CONNECTIONS=32
out=gpurun_out/r2_sweep4.log
: > $out
python -m pytest tests/test_gpu_gemm.py -x -q -m gpu >> $out 2>&1
B="decode_branches=1"
for opt in "$B" "$B,gemm_min_ctas=60" "$B,gemm_min_ctas=1" \
           "$B,gemm_min_ctas=1,gemm_persist_min_tiles=8,gemm_tiles_per_cta=2" \
           "$B,gemm_min_ctas=1,gemm_persist_min_tiles=8,gemm_tiles_per_cta=4" \
           "$B,gemm_min_ctas=1,gemm_persist_min_tiles=8,gemm_tiles_per_cta=8" \
           "$B,gemm_min_ctas=60,gemm_persist_min_tiles=8,gemm_tiles_per_cta=4" \
           "$B,gemm_min_ctas=1,gemm_persist_min_tiles=8,gemm_tiles_per_cta=4,gemm_persistent_stages=4"; do
  timeout 300 python scripts/inflight_probe.py 512 256 6 6 "$opt" >> $out 2>&1 || echo "FAILED $opt" >> $out
done
export TEXOCR_B200_LIB=$PWD/texocr_b200/libtexocr_b200_s7.so
echo "== 7-stage attention ring" >> $out
timeout 300 python scripts/inflight_probe.py 512 256 6 6 "$B" >> $out 2>&1 || echo "FAILED s7" >> $out
timeout 300 python scripts/inflight_probe.py 512 256 6 6 "$B,gemm_min_ctas=1,gemm_persist_min_tiles=8,gemm_tiles_per_cta=4" >> $out 2>&1 || echo "FAILED s7" >> $out
cat $out
