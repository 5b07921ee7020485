import sys, torch
sys.path.insert(0, "/root/repo")
import texocr_b200
from texocr_b200 import spec, synth
cfg = spec.default_config(max_length=256); cfg["device"]="cuda:0"
d = spec.dims_from_config(cfg)
m = texocr_b200.create_model(cfg, precision="bf16"); m.load_state_dict(synth.seeded_state_dict(d, seed=0))
eng = m.engine()
img = synth.synth_images(512, 64, 384, seed=1234).cuda()
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for nb, cps, stag in ((8, 4, 30), (12, 4, 20), (16, 4, 15), (16, 3, 15), (16, 6, 15), (8, 4, 0), (8, 6, 30)):
    eng.set_option("decode_branches", nb); eng.set_option("attn_ctas_per_sm", cps); eng.set_option("stagger_us", stag)
    print(f"branches {nb} attn_ctas/sm {cps} stagger {stag}: generate(256) {timeit(lambda: m.generate(img, 256)):.1f} ms")
