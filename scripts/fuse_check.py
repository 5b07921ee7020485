"""Decode-step variants at the headline size: fused LayerNorm GEMMs on/off (tokens must be identical)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from texocr_b200 import synth
from texocr_b200.model import create_model
from texocr_b200.spec import default_config

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T = int(sys.argv[2]) if len(sys.argv) > 2 else 256
opts = sys.argv[3:] or ["fuse_ln=0", "fuse_ln=1"]
m = create_model(default_config(), precision="bf16")
m.load_state_dict(synth.seeded_state_dict(m.dims))
eng = m.engine()
img = synth.synth_images(B, 64, 384, seed=21).cuda()
ref = None
for o in opts:
    for kv in o.split(","):
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    try:
        for _ in range(2):
            out = m.generate(img, T)
    except RuntimeError as ex:
        print("%-40s FAILED: %s" % (o, ex), flush=True)
        continue
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 4
    for _ in range(n):
        out = m.generate(img, T)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    same = "-" if ref is None else str(bool(torch.equal(out, ref)))
    if ref is None:
        ref = out
    print("%-40s %.2f ms per generate -> %.1f eq/s   identical to first: %s" % (o, dt * 1e3, B / dt, same), flush=True)
