"""Cluster-persistent decode kernel vs the per-branch kernel graphs: token agreement, teacher-forced agreement, timing."""
import sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from texocr_b200 import synth
from texocr_b200.model import create_model
from texocr_b200.spec import default_config

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    m = create_model(default_config(), precision="bf16")
    m.load_state_dict(synth.seeded_state_dict(m.dims))
    eng = m.engine()
    img = synth.synth_images(B, 64, 384, seed=21).cuda()
    outs = []
    for mega in (1, 0):
        eng.set_option("decode_mega", mega)
        o = m.generate(img, T)
        torch.cuda.synchronize()
        outs.append(o)
        print("mega", mega, "shape", tuple(o.shape), "first row", o[0, :12].tolist(), flush=True)
    same = (outs[0] == outs[1]).float().mean().item()
    first_diff = ((outs[0] != outs[1]).float().argmax(1))[(outs[0] != outs[1]).any(1)]
    print("agreement mega vs branches: %.4f; rows differing %d; first-diff steps %s" % (same, int((outs[0] != outs[1]).any(1).sum()), first_diff[:10].tolist()))
    enc = m.encoder(img)
    ids = torch.cat((torch.full((B, 1), m.dims.bos, device="cuda"), outs[0][:, :-1]), 1)
    logits = m.decoder.net(ids, enc=enc)
    print("teacher-forced agreement (mega): %.4f" % (logits.argmax(-1) == outs[0]).float().mean().item())
    ids = torch.cat((torch.full((B, 1), m.dims.bos, device="cuda"), outs[1][:, :-1]), 1)
    logits = m.decoder.net(ids, enc=enc)
    print("teacher-forced agreement (branches): %.4f" % (logits.argmax(-1) == outs[1]).float().mean().item())
    eng.set_option("decode_mega", 1)
    a = m.generate(img, T)
    print("mega deterministic:", bool(torch.equal(a, outs[0])))
    for mega in (1, 0):
        eng.set_option("decode_mega", mega)
        for _ in range(2):
            m.generate(img, T)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 3
        for _ in range(n):
            m.generate(img, T)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        print("mega %d: %.2f ms per generate (B=%d, T=%d) -> %.1f eq/s" % (mega, dt * 1e3, B, T, B / dt), flush=True)

if __name__ == "__main__":
    main()
