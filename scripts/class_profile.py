"""Per-kernel-class device time of an eagerly launched generate (CUDA events around every launch), for option sets."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from texocr_b200 import synth
from texocr_b200.model import create_model
from texocr_b200.spec import default_config

B = int(sys.argv[1]); T = int(sys.argv[2]); opts = sys.argv[3:]
m = create_model(default_config(), precision="bf16")
m.load_state_dict(synth.seeded_state_dict(m.dims))
eng = m.engine()
img = synth.synth_images(B, 64, 384, seed=21).cuda()
for o in opts:
    for kv in o.split(","):
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    m.generate(img, T)
    eng.profile_enable(True)
    m.generate(img, T)
    rows = eng.profile_read()
    eng.profile_enable(False)
    print("==", o)
    for r in sorted(rows, key=lambda r: -r["ms"]):
        if r["name"].startswith("dec"):
            print("  %-16s %7d launches %9.3f ms  %7.2f us/launch" % (r["name"], r["launches"], r["ms"], 1e3 * r["ms"] / r["launches"]))
