"""Per-kernel-class device time of one eagerly launched generate call (texocr_profile_*): launches, total ms, us per launch,
achieved GB/s or TFLOP/s on the engine's algorithmic byte / flop accounting.  usage: class_times.py [B] [T] [opt=val,...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from texocr_b200 import synth
from texocr_b200.model import create_model
from texocr_b200.spec import default_config

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T = int(sys.argv[2]) if len(sys.argv) > 2 else 256
m = create_model(default_config(), precision="bf16")
m.load_state_dict(synth.seeded_state_dict(m.dims))
eng = m.engine()
for o in sys.argv[3:] or [""]:
    for kv in [x for x in o.split(",") if x]:
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    img = synth.synth_images(B, 64, 384, seed=21).cuda()
    m.generate(img, T)
    eng.profile_enable(True)
    m.generate(img, T)
    rows = eng.profile_read()
    eng.profile_enable(False)
    print(f"== {o or 'defaults'}  (B={B}, T={T})")
    for r in sorted(rows, key=lambda r: -r["ms"]):
        us = r["ms"] * 1e3 / max(1, r["launches"])
        print(f"{r['name']:16s} {r['launches']:6d} launches {r['ms']:8.2f} ms {us:8.2f} us/launch {r['bytes'] / max(r['ms'], 1e-9) / 1e6:8.1f} GB/s {r['flops'] / max(r['ms'], 1e-9) / 1e9:8.2f} TFLOP/s")
