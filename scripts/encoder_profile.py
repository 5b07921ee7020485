"""Per-kernel-class device time of one encoder pass (CUDA events around every launch; texocr_profile_*)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import texocr_b200
from texocr_b200 import spec, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cfg = spec.default_config(max_length=256); cfg["device"] = "cuda:0"
d = spec.dims_from_config(cfg)
m = texocr_b200.create_model(cfg, precision="bf16"); m.load_state_dict(synth.seeded_state_dict(d, seed=0))
eng = m.engine()
if len(sys.argv) > 2:
    for kv in sys.argv[2].split(','):
        k, v = kv.split('='); eng.set_option(k, int(v))
    print('options:', sys.argv[2])
img = synth.synth_images(B, 64, 384, seed=1234).cuda()
for _ in range(2): m.encoder(img)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): m.encoder(img)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"encoder B={B} 64x384: {ms:.2f} ms -> {B / ms * 1e3:.0f} img/s; algorithmic {synth.encoder_flops(64, 384) * B / ms / 1e9:.1f} TFLOP/s")
eng.profile_enable(True); m.encoder(img); rows = eng.profile_read(); eng.profile_enable(False)
tot = sum(r["ms"] for r in rows)
for r in sorted(rows, key=lambda r: -r["ms"]):
    extra = f"{r['flops'] / r['ms'] / 1e9:.0f} TFLOP/s" if r["flops"] else f"{r['bytes'] / r['ms'] / 1e6:.0f} GB/s"
    print(f"  {r['name']:14s} {r['launches']:4d} launches {r['ms']:7.3f} ms ({r['ms'] / tot * 100:4.1f}%)  {extra}")
print(f"  sum of kernels {tot:.2f} ms")
