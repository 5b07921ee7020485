"""Phase timers of attn_abs_kernel (engine option attn_trace): per CTA, averaged over one generate call:
wait for the predecessor kernel, first stage arrival, stage loop, epilogue.  usage: abs_phases.py [B] [T] [opt=val,...]"""
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
for kv in (sys.argv[3].split(",") if len(sys.argv) > 3 else []):
    k, v = kv.split("=")
    eng.set_option(k, int(v))
img = synth.synth_images(B, 64, 384, seed=21).cuda()
eng.set_option("attn_trace", 1)
m.generate(img, T)
m.generate(img, T)
torch.cuda.synchronize()
raw = eng.debug_read("attn_trace", 16 * 3 * 2048 * 2 + 32).view(torch.int64).cpu()
d = raw[16 * 3 * 2048:].double()
for name, o in (("self", 0), ("cross", 8)):
    n = max(1.0, float(d[o + 4]))
    print(f"{name:5s}: CTAs {int(n)}  wait-for-predecessor {float(d[o]) / n / 1e3:6.2f} us  first stage {float(d[o + 1]) / n / 1e3:6.2f} us  stage loop {float(d[o + 2]) / n / 1e3:6.2f} us  epilogue {float(d[o + 3]) / n / 1e3:6.2f} us")
