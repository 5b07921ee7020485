#!/bin/bash
# end-of-round ncu --set full captures of the kernels that changed in the second session
cd "$(dirname "$0")/.."
NCU="ncu --set full --clock-control none --kernel-name-base demangled"
CMD="python scripts/profile_generate.py --batch 512 --max-len 48 --warm 0 --no-graph --branches 1"
# decode Q' GEMM with eight epilogue warps (late in the call: decode steps), conv GEMM with GroupNorm partials (3x3, K = 576), 16-byte GroupNorm apply
timeout 600 $NCU -k 'regex:tc_gemm_kernel<\(int\)64, \(int\)0, __nv_bfloat16, \(int\)1, \(int\)0, \(int\)2>' -s 400 -c 2 -o gpurun_out/r2b_dec_q_gemm -f $CMD > gpurun_out/r2b_ncu_q.log 2>&1
timeout 600 $NCU -k 'regex:tc_gemm_persistent_kernel<\(int\)(128|64), \(int\)0, float, \(int\)3' -s 2 -c 3 -o gpurun_out/r2b_conv_gemm -f $CMD > gpurun_out/r2b_ncu_conv.log 2>&1
timeout 600 $NCU -k 'regex:gn_apply8_kernel' -s 3 -c 3 -o gpurun_out/r2b_gn_apply8 -f $CMD > gpurun_out/r2b_ncu_gn.log 2>&1
# ragged batch: the gather convolution
cat > /tmp/ragged_enc.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import texocr_b200
from texocr_b200 import spec, synth
cfg = spec.default_config(max_length=256); cfg["device"] = "cuda:0"
d = spec.dims_from_config(cfg)
m = texocr_b200.create_model(cfg, precision="bf16"); m.load_state_dict(synth.seeded_state_dict(d, seed=0))
widths = synth.synth_widths(256, seed=77)
rag = [synth.synth_images(1, 64, w, seed=500 + i)[0].cuda() for i, w in enumerate(widths)]
m.encoder(rag); torch.cuda.synchronize()
PY
timeout 600 $NCU -k 'regex:tc_conv_gather_kernel' -s 1 -c 3 -o gpurun_out/r2b_conv_gather -f python /tmp/ragged_enc.py > gpurun_out/r2b_ncu_gather.log 2>&1
# the reports stay on the box (64 MiB limit on what comes back): only their summaries travel
for n in dec_q_gemm conv_gemm gn_apply8 conv_gather; do
  echo "## $n" >> gpurun_out/r2b_ncu_summary.md
  python scripts/ncu_summary.py gpurun_out/r2b_$n.ncu-rep >> gpurun_out/r2b_ncu_summary.md 2>&1
  rm -f gpurun_out/r2b_$n.ncu-rep
done
cat gpurun_out/r2b_ncu_summary.md | cut -c1-400
