"""How much throughput is left on the table by running one batch at a time: K generate calls of B equations each,
issued from `n` host threads, one engine handle and one stream per thread (ctypes drops the GIL during the call).
usage: inflight_probe.py [B] [T] [n1,n2,...] [calls per thread] [opt=val,...]"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from texocr_b200 import synth
from texocr_b200.model import create_model
from texocr_b200.spec import default_config

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ns = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "1,2,3").split(",")]
calls = int(sys.argv[4]) if len(sys.argv) > 4 else 6
models = []
img = synth.synth_images(B, 64, 384, seed=21).cuda()
sd = None
for i in range(max(ns)):
    m = create_model(default_config(), precision="bf16")
    if sd is None:
        sd = synth.seeded_state_dict(m.dims)
    m.load_state_dict(sd)
    if len(sys.argv) > 5:
        for kv in sys.argv[5].split(","):
            k, v = kv.split("=")
            m.engine().set_option(k, int(v))
    models.append(m)
ref = models[0].generate(img, T).clone()
for n in ns:
    outs = [None] * n
    bar = threading.Barrier(n + 1)

    def work(i):
        st = torch.cuda.Stream(priority=int(os.environ.get('WORKER_PRIO', '0')))
        with torch.cuda.stream(st):
            for _ in range(2):
                models[i].generate(img, T)
            st.synchronize()
            bar.wait()
            for _ in range(calls):
                outs[i] = models[i].generate(img, T)
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
    same = all(torch.equal(o, ref) for o in outs)
    print((sys.argv[5] if len(sys.argv) > 5 else "") + " B=%d in flight=%d: %.2f ms per batch -> %.1f eq/s   tokens identical: %s" % (B, n, dt * 1e3 / (n * calls), B * n * calls / dt, same), flush=True)
