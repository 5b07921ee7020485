#!/bin/bash
# ncu --set full of the ViT-block GEMMs of the encoder (persistent tcgen05 kernel, bf16 operands) and of two backbone convolutions
cd "$(dirname "$0")/.."
NCU="ncu --set full --clock-control none --kernel-name-base demangled"
timeout 600 $NCU -k 'regex:tc_gemm_persistent_kernel' -s 38 -c 8 -o gpurun_out/r2_enc_vit_gemms -f python scripts/encoder_only.py 512 1 > gpurun_out/r2_ncu_vit.log 2>&1
tail -3 gpurun_out/r2_ncu_vit.log
ls -la gpurun_out/r2_enc_vit_gemms.ncu-rep
