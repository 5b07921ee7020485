"""Profiling target: encoder passes only (bf16 tier, B x 64 x 384).  usage: encoder_only.py [B] [passes]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import texocr_b200
from texocr_b200 import spec, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = spec.default_config(max_length=256); cfg["device"] = "cuda:0"
d = spec.dims_from_config(cfg)
m = texocr_b200.create_model(cfg, precision="bf16"); m.load_state_dict(synth.seeded_state_dict(d, seed=0))
img = synth.synth_images(B, 64, 384, seed=1234).cuda()
for _ in range(n):
    m.encoder(img)
torch.cuda.synchronize()
