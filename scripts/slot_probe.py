"""Where the SM residency of a decode goes with several batches in flight: every engine runs with attn_trace=1, which makes
the attention and decode GEMM kernels add their per-CTA %globaltimer spans (waiting for the predecessor grid / working)
into a small buffer.  usage: slot_probe.py [B] [T] [in flight] [calls] [opt=val,...]"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from texocr_b200 import synth
from texocr_b200.model import create_model
from texocr_b200.spec import default_config

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n = int(sys.argv[3]) if len(sys.argv) > 3 else 6
calls = int(sys.argv[4]) if len(sys.argv) > 4 else 4
opts = sys.argv[5] if len(sys.argv) > 5 else ""
DECODE_ONLY = os.environ.get("SLOT_PROBE_DECODE_ONLY", "0") == "1"      # time decoder.generate over a precomputed memory (any B)
img = synth.synth_images(min(B, 512), 64, 384, seed=21).cuda()
models, sd = [], None
for i in range(n):
    m = create_model(default_config(), precision="bf16")
    if sd is None:
        sd = synth.seeded_state_dict(m.dims)
    m.load_state_dict(sd)
    e = m.engine()
    e.set_option("decode_branches", 1)
    for kv in [x for x in opts.split(",") if x]:
        k, v = kv.split("=")
        e.set_option(k, int(v))
    e.set_option("attn_trace", 1)
    models.append(m)
bar = threading.Barrier(n + 1)
enc = start = None
if DECODE_ONLY:
    e512 = models[0].encoder(img)
    enc = torch.cat([e512] * max(1, B // e512.shape[0]), 0).contiguous()
    start = torch.full((enc.shape[0], 1), models[0].dims.bos, dtype=torch.long, device="cuda")

def run(m):
    if DECODE_ONLY:
        m.decoder.generate(start_tokens=start, eos_tok=None, max_len=T, enc=enc)
    else:
        m.generate(img, T)

def work(i):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(2):
            run(models[i])
        st.synchronize()
        bar.wait()
        for _ in range(calls):
            run(models[i])
        st.synchronize()
    bar.wait()

th = [threading.Thread(target=work, args=(i,)) for i in range(n)]
for t in th:
    t.start()
bar.wait()
t0 = time.perf_counter()
bar.wait()
dt = time.perf_counter() - t0
for t in th:
    t.join()
ms_batch = dt * 1e3 / (n * calls)
print(f"{opts} B={B} in flight={n}: {ms_batch:.2f} ms per batch -> {B * n * calls / dt:.1f} eq/s")
tot = {}
for m in models:
    raw = m.engine().debug_read("attn_trace", 16 * 3 * 2048 * 2 + 128).view(torch.int64).cpu()
    d = raw[16 * 3 * 2048:].double()
    for name, o in (("attn_self", 0), ("attn_cross", 8)):
        a = tot.setdefault(name, [0.0] * 8)
        for k in range(7):
            a[k] += float(d[o + k])
    a = tot.setdefault("dec_gemm", [0.0] * 8)
    for k in range(7):
        a[k] += float(d[16 + k])
    gx = tot.setdefault("gemm_x", [0.0, 0.0])
    gx[0] += float(d[16 + 7]); gx[1] += float(d[16 + 8])
# the buffers hold the sums of each engine's LAST generate call = one batch per engine
slot_total = 0.0
for name in ("attn_self", "attn_cross"):
    a = tot[name]
    c1, c = max(1.0, a[4]), max(1.0, a[6])
    print(f"{name:10s}: CTAs/batch {c / n:9.0f}  wait {a[0] / c1 / 1e3:6.2f} us  first stage {a[1] / c1 / 1e3:5.2f}  loop {a[2] / c1 / 1e3:6.2f}  epilogue {a[3] / c1 / 1e3:5.2f}"
          f"  | residency {a[5] / c / 1e3:6.2f} us per CTA, {a[5] / n / 1e6:8.2f} ms-CTA per batch")
    slot_total += a[5] / n / 1e6
a = tot["dec_gemm"]
c = max(1.0, a[2])
print(f"dec_gemm  : CTAs/batch {c / n:9.0f}  wait {a[0] / c / 1e3:6.2f} us  work {a[1] / c / 1e3:6.2f} us  | residency {(a[0] + a[1]) / c / 1e3:6.2f} us per CTA, "
      f"{(a[0] + a[1]) / n / 1e6:8.2f} ms-CTA per batch (work only {a[1] / n / 1e6:8.2f})")
print(f"            per CTA from entry: first k-block landed {a[3] / c / 1e3:5.2f} us, last k-block landed {a[4] / c / 1e3:5.2f} us, accumulator complete {a[6] / c / 1e3:5.2f} us, "
      f"first chunk in registers {tot['gemm_x'][0] / c / 1e3:5.2f} us, first chunk staged {tot['gemm_x'][1] / c / 1e3:5.2f} us, epilogue warp done {a[5] / c / 1e3:5.2f} us")
slot_total += (a[0] + a[1]) / n / 1e6
print(f"sum of residency: {slot_total:.1f} ms-CTA per batch; wall {ms_batch:.2f} ms per batch -> {slot_total / ms_batch:.0f} CTAs resident on average ({slot_total / ms_batch / 148:.2f} per SM)")
