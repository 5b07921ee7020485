"""TMA im2col convolutions (implicit GEMM) vs the explicit im2col path: same bits expected; encoder timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from texocr_b200 import synth
from texocr_b200.model import create_model
from texocr_b200.spec import default_config

m = create_model(default_config(), precision="bf16")
m.load_state_dict(synth.seeded_state_dict(m.dims))
eng = m.engine()
for (B, H, W) in ((2, 64, 384), (3, 48, 208), (1, 160, 1008), (16, 64, 384)):
    img = synth.synth_images(B, H, W, seed=5).cuda()
    outs = []
    for flag in (0, 1):
        eng.set_option("im2col_tma", flag)
        outs.append(m.encoder(img).float())
        torch.cuda.synchronize()
    d = (outs[0] - outs[1]).abs().max().item()
    print("B=%d %dx%d: max |explicit - implicit| = %.3e, finite %s, equal %s" % (B, H, W, d, bool(torch.isfinite(outs[1]).all()), bool(torch.equal(outs[0], outs[1]))), flush=True)
img = synth.synth_images(512, 64, 384, seed=1).cuda()
for flag in (0, 1):
    eng.set_option("im2col_tma", flag)
    for _ in range(2):
        m.encoder(img)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        m.encoder(img)
    torch.cuda.synchronize()
    print("im2col_tma=%d: encoder 512 images %.2f ms" % (flag, (time.perf_counter() - t0) / 5 * 1e3), flush=True)
