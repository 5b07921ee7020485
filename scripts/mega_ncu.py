"""One short generate for profiling the cluster-persistent decode kernel under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from texocr_b200 import synth
from texocr_b200.model import create_model
from texocr_b200.spec import default_config

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T = int(sys.argv[2]) if len(sys.argv) > 2 else 32
m = create_model(default_config(), precision="bf16")
m.load_state_dict(synth.seeded_state_dict(m.dims))
img = synth.synth_images(B, 64, 384, seed=21).cuda()
m.engine().set_option("mega_steps", T)
out = m.generate(img, T)
torch.cuda.synchronize()
print(out.shape)
