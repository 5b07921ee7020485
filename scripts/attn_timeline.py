"""Timeline of the decode-attention launches under the real multi-branch graph replay (globaltimer stamps written by
attn_decode_tma_kernel when the engine option attn_trace is on): per launch slot k (2*layer + {self, cross}) the time from
CTA entry to griddepcontrol release, the streaming time, and the gap to the branch's next attention launch (= the GEMM /
LayerNorm chain between them).  usage: attn_timeline.py [B] [T] [opt=val,...]"""
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
if len(sys.argv) > 3:
    for kv in sys.argv[3].split(","):
        k, v = kv.split("=")
        eng.set_option(k, int(v))
img = synth.synth_images(B, 64, 384, seed=21).cuda()
eng.set_option("attn_trace", 1)
for _ in range(3):
    m.generate(img, T)
torch.cuda.synchronize()
raw = eng.debug_read("attn_trace", 16 * 3 * 2048 * 2).view(torch.int64).cpu().reshape(16, 3, 256, 8).double()
nb = int((raw[:, 2, 10, 0] > 0).sum())
entry, ready, end = raw[:nb, 0], raw[:nb, 1], raw[:nb, 2]          # [branch][step][k] ns
lo, hi = T // 4, T - 2
print(f"B={B} T={T} branches={nb}; steps {lo}..{hi}; all times in us, mean over branches and steps")
step_time = (end[:, hi, 7] - end[:, lo, 7]) / (hi - lo) / 1e3
print("step time per branch (us):", [round(float(x), 1) for x in step_time])
print(" k  kind   wait(entry->ready)  stream(ready->end)  chain gap to next attention (end -> next entry | next ready)")
tot_s = tot_g = 0.0
for k in range(8):
    w = (ready[:, lo:hi, k] - entry[:, lo:hi, k]).mean() / 1e3
    s = (end[:, lo:hi, k] - ready[:, lo:hi, k]).mean() / 1e3
    if k < 7:
        g_e = (entry[:, lo:hi, k + 1] - end[:, lo:hi, k]).mean() / 1e3
        g_r = (ready[:, lo:hi, k + 1] - end[:, lo:hi, k]).mean() / 1e3
    else:
        g_e = (entry[:, lo + 1:hi + 1, 0] - end[:, lo:hi, 7]).mean() / 1e3
        g_r = (ready[:, lo + 1:hi + 1, 0] - end[:, lo:hi, 7]).mean() / 1e3
    tot_s += float(s); tot_g += float(g_r)
    print(f"{k:2d}  {'self ' if k % 2 == 0 else 'cross'}  {float(w):8.2f}            {float(s):8.2f}            {float(g_e):8.2f} | {float(g_r):8.2f}")
print(f"sum stream {tot_s:.1f} us, sum gaps {tot_g:.1f} us per step")
# phase picture of one step: when does each branch stream? (offsets from the earliest stamp of that step)
st = (lo + hi) // 2
t0 = float(ready[:, st, :].min())
print(f"step {st}: [ready, end] of every attention launch, us from the first one")
for b in range(nb):
    print(f"  branch {b}: " + "  ".join(f"{(float(ready[b, st, k]) - t0) / 1e3:6.1f}-{(float(end[b, st, k]) - t0) / 1e3:6.1f}" for k in range(8)))
# how many branches stream at the same time, on average (time-weighted over the window)
ev = []
for b in range(nb):
    for s_ in range(lo, hi):
        for k in range(8):
            ev.append((float(ready[b, s_, k]), 1)); ev.append((float(end[b, s_, k]), -1))
ev.sort()
cur, last, acc = 0, ev[0][0], [0.0] * (nb + 2)
for t, d in ev:
    acc[min(cur, nb + 1)] += t - last
    last = t
    cur += d
tot = sum(acc)
print("fraction of time with n branches streaming:", {n: round(a / tot, 3) for n, a in enumerate(acc) if a > 0})
