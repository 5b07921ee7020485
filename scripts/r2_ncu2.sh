#!/bin/bash
# ncu evidence for the warp-per-sequence attention kernel (attn_seq_kernel) + refreshed launch list
cd "$(dirname "$0")/.."
CMD="python scripts/profile_generate.py --batch 512 --max-len 200 --warm 0 --no-graph --branches 1"
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 900 $NCU -k 'regex:attn_seq_kernel' -s 1496 -c 2 -o gpurun_out/r2_attn_seq -f $CMD > gpurun_out/r2_ncu_seq.log 2>&1
CMD2="python scripts/profile_generate.py --batch 512 --max-len 48 --warm 0 --no-graph --branches 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_v2.csv $CMD2 > gpurun_out/r2_launches_v2.log 2>&1
ls -la gpurun_out/r2_attn_seq.ncu-rep gpurun_out/r2_launches_v2.csv; tail -2 gpurun_out/r2_ncu_seq.log
