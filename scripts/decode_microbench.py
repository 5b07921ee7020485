"""Timing experiments for the decode loop (device-resident inputs, CUDA events, B=512, 64x384, max_len 256)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import texocr_b200  # noqa: E402
from texocr_b200 import spec, synth  # noqa: E402

cfg = spec.default_config(max_length=256)
cfg["device"] = "cuda:0"
d = spec.dims_from_config(cfg)
m = texocr_b200.create_model(cfg, precision="bf16")
m.load_state_dict(synth.seeded_state_dict(d, seed=0))
eng = m.engine()
img = synth.synth_images(512, 64, 384, seed=1234).cuda()


def timeit(fn, n=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


enc_ms = timeit(lambda: m.encoder(img))
print(f"encoder only: {enc_ms:.2f} ms")
def decode_us(label):
    t256 = timeit(lambda: m.generate(img, 256))
    t64 = timeit(lambda: m.generate(img, 64))
    print(f"{label}: generate(256) {t256:.1f} ms, generate(64) {t64:.1f} ms -> decode {(t256 - t64) / 192 * 1000:.0f} us/step (t in 64..256)")


for cps in (3, 4, 6):
    eng.set_option("attn_ctas_per_sm", cps)
    for nb, stag in ((1, 0), (4, 0), (4, 60), (8, 30)):
        eng.set_option("decode_branches", nb)
        eng.set_option("stagger_us", stag)
        decode_us(f"NS=4 attn_ctas/sm={cps} branches={nb} stagger={stag}us")
