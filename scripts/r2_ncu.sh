#!/bin/bash
# round-2 ncu evidence: launch list of one generate call + full captures of the conv GEMM, the decode Q' GEMM, the encoder
# attention and the absorbed decode attention (tensor-pipe and DRAM metrics).  One GPU, never under a timed region.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python scripts/profile_generate.py --batch 512 --max-len 48 --warm 0 --no-graph --branches 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_launches.log 2>&1
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 600 $NCU -k regex:"tc_gemm_persistent_kernel<128, 0, float, 3>" -s 8 -c 2 -o gpurun_out/r2_conv_gemm -f $CMD > gpurun_out/r2_ncu_conv.log 2>&1
timeout 600 $NCU -k regex:"tc_gemm_kernel<64, 0, __nv_bfloat16, 1, 0, 1>" -s 400 -c 2 -o gpurun_out/r2_dec_q_gemm -f $CMD > gpurun_out/r2_ncu_q.log 2>&1
timeout 600 $NCU -k regex:"tc_gemm_kernel<32, 1, float, 1, 0, 1>" -s 200 -c 1 -o gpurun_out/r2_dec_wo_gemm -f $CMD > gpurun_out/r2_ncu_wo.log 2>&1
timeout 600 $NCU -k regex:attn_enc_mma_kernel -s 1 -c 1 -o gpurun_out/r2_attn_enc -f $CMD > gpurun_out/r2_ncu_enc.log 2>&1
timeout 600 $NCU -k regex:attn_abs_kernel -s 320 -c 2 -o gpurun_out/r2_attn_abs -f $CMD > gpurun_out/r2_ncu_abs.log 2>&1
timeout 600 $NCU -k regex:gn_apply_kernel -s 10 -c 1 -o gpurun_out/r2_gn_apply -f $CMD > gpurun_out/r2_ncu_gn.log 2>&1
ls -la gpurun_out/r2_*.ncu-rep
tail -3 gpurun_out/r2_ncu_q.log
