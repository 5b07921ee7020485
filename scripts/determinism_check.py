import sys, torch
sys.path.insert(0, "/root/repo")
import texocr_b200
from texocr_b200 import spec, synth
cfg = spec.default_config(max_length=256); cfg["device"]="cuda:0"
d = spec.dims_from_config(cfg)
sd = synth.seeded_state_dict(d, seed=0)
B = 512
ncalls, maxlen = 30, 32
opts = []
for o in sys.argv[1:]:
    if o.startswith("B="): B = int(o[2:])
    elif o.startswith("calls="): ncalls = int(o[6:])
    elif "=" in o: opts.append(o.split("="))
img = synth.synth_images(B, 64, 384, seed=1234).cuda()
m16 = texocr_b200.create_model(cfg, precision="bf16"); m16.load_state_dict(sd)
eng = m16.engine()
for k, v in opts: eng.set_option(k, int(v))
outs = [m16.generate(img, maxlen).clone() for _ in range(ncalls)]
from collections import Counter
sig = [hash(o.cpu().numpy().tobytes()) for o in outs]
cnt = Counter(sig)
base = outs[sig.index(cnt.most_common(1)[0][0])]
rows = Counter()
for o in outs:
    for r in (o != base).any(1).nonzero().flatten().tolist(): rows[(r, int((o[r] != base[r]).nonzero()[0]))] += 1
print(f"B={B} opts={opts}: distinct {len(cnt)} of {ncalls}; odd calls {ncalls - cnt.most_common(1)[0][1]}; (row,step)x{dict(rows)}")
