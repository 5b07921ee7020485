"""Profiling target: one warm generate() of BASELINE configs[2] (or a scaled-down variant) for ncu.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python scripts/profile_generate.py --batch 512 --max-len 256
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import texocr_b200  # noqa: E402
from texocr_b200 import spec, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--max-len", type=int, default=256)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--warm", type=int, default=1)
ap.add_argument("--no-graph", action="store_true")
ap.add_argument("--branches", type=int, default=0, help="decode branches (0 = engine default)")
args = ap.parse_args()
cfg = spec.default_config(max_length=256)
cfg["device"] = "cuda:0"
d = spec.dims_from_config(cfg)
m = texocr_b200.create_model(cfg, precision=args.precision)
m.load_state_dict(synth.seeded_state_dict(d, seed=0))
img = synth.synth_images(args.batch, 64, 384, seed=1234).cuda()
if args.no_graph:
    m.engine().set_option("cuda_graph", 0)
if args.branches:
    m.engine().set_option("decode_branches", args.branches)
for _ in range(args.warm):
    m.generate(img, args.max_len)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("profiled_generate")
tok = m.generate(img, args.max_len)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("tokens", tuple(tok.shape), "launches", m.engine().kernel_launches())
