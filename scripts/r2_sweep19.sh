#!/bin/bash
cd "$(dirname "$0")/.."
export CUDA_DEVICE_MAX_CONNECTIONS=32
out=gpurun_out/r2_sweep19.log
: > $out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py -x -q -m gpu -k "im2col or groupnorm or encoder" 2>&1 | tail -12 >> $out
timeout 300 python - >> $out 2>&1 <<'PY'
import torch, sys, os
sys.path.insert(0, os.getcwd())
import texocr_b200
from texocr_b200 import spec, synth
cfg = spec.default_config(max_length=256); cfg["device"] = "cuda:0"
d = spec.dims_from_config(cfg)
m = texocr_b200.create_model(cfg, precision="bf16"); m.load_state_dict(synth.seeded_state_dict(d, seed=0))
eng = m.engine()
widths = synth.synth_widths(256, seed=77)
rag = [synth.synth_images(1, 64, w, seed=500 + i)[0].cuda() for i, w in enumerate(widths)]
flops = sum(synth.encoder_flops(64, w) for w in widths)
for gather in (0, 1):
    eng.set_option("conv_gather", gather)
    for _ in range(2): m.encoder(rag)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): m.encoder(rag)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"config 2 ragged (256 images, widths 128..1008) conv_gather={gather}: {ms:.2f} ms -> {256 / ms * 1e3:.0f} img/s, {flops / ms / 1e9:.1f} TFLOP/s algorithmic")
    eng.profile_enable(True); m.encoder(rag); rows = eng.profile_read(); eng.profile_enable(False)
    for r in sorted(rows, key=lambda r: -r["ms"]):
        print(f"    {r['name']:14s} {r['launches']:4d} launches {r['ms']:7.3f} ms")
PY
cat $out
