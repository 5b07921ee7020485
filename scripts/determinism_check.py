import sys, torch
sys.path.insert(0, "/root/repo")
import texocr_b200
from texocr_b200 import spec, synth
cfg = spec.default_config(max_length=256); cfg["device"]="cuda:0"
d = spec.dims_from_config(cfg)
sd = synth.seeded_state_dict(d, seed=0)
img = synth.synth_images(512, 64, 384, seed=1234)
def diff(a, b):
    a, b = a.cpu(), b.cpu()
    rows = (a != b).any(1).nonzero().flatten()
    return len(rows), [(int(r), int((a[r] != b[r]).nonzero()[0])) for r in rows[:6]]
m32 = texocr_b200.create_model(cfg, precision="fp32"); m32.load_state_dict(sd)
if "--no32" not in sys.argv:
    m32.generate(img.cuda(), 32); m32.generate(img[:8].cuda(), 32)
m16 = texocr_b200.create_model(cfg, precision="bf16"); m16.load_state_dict(sd)
eng = m16.engine()
for opt in sys.argv[1:]:
    if "=" in opt:
        k, v = opt.split("="); eng.set_option(k, int(v)); print("option", k, v)
outs = []
for i in range(10):
    outs.append(m16.generate(img.cuda(), 32).clone())
    if i % 3 == 1: m16.generate(img[:8].cuda(), 32)
from collections import Counter
sig = [hash(o.cpu().numpy().tobytes()) for o in outs]
print("distinct results:", len(set(sig)), Counter(sig).most_common())
base = outs[max(range(10), key=lambda i: sig.count(sig[i]))]
for i, o in enumerate(outs):
    n, first = diff(base, o)
    if n: print(" call", i, "differs in", n, "rows", first)
