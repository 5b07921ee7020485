#!/bin/bash
# round-2 ncu evidence: launch list of one generate call + full captures of the conv GEMM, the decode GEMMs, the encoder
# attention, the absorbed decode attention and the GroupNorm apply (tensor-pipe and DRAM metrics).  One GPU, never under a timed region.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python scripts/profile_generate.py --batch 512 --max-len 48 --warm 0 --no-graph --branches 1"
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 600 $NCU -k 'regex:tc_gemm_persistent_kernel<\(int\)128, \(int\)0, float, \(int\)3>' -s 8 -c 2 -o gpurun_out/r2_conv_gemm -f $CMD > gpurun_out/r2_ncu_conv.log 2>&1
timeout 600 $NCU -k 'regex:tc_gemm_kernel<\(int\)64, \(int\)0, __nv_bfloat16, \(int\)1, \(int\)0>' -s 400 -c 2 -o gpurun_out/r2_dec_q_gemm -f $CMD > gpurun_out/r2_ncu_q.log 2>&1
timeout 600 $NCU -k 'regex:tc_gemm_kernel<\(int\)32, \(int\)1, float, \(int\)1, \(int\)0>' -s 200 -c 1 -o gpurun_out/r2_dec_wo_gemm -f $CMD > gpurun_out/r2_ncu_wo.log 2>&1
timeout 600 $NCU -k 'regex:tc_gemm_kernel<\(int\)32, \(int\)4, float' -s 30 -c 1 -o gpurun_out/r2_dec_logits_gemm -f $CMD > gpurun_out/r2_ncu_logits.log 2>&1
ls -la gpurun_out/r2_*.ncu-rep
tail -3 gpurun_out/r2_ncu_q.log gpurun_out/r2_ncu_conv.log
