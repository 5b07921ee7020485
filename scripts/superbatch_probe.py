"""Decode-only throughput against rows per launch: N engine handles in flight, each decoding B rows with model.decoder.generate
over a precomputed encoder memory (no encoder work in the timed region).  usage: superbatch_probe.py B1,B2,.. n1,n2,.. [T] [calls]"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from texocr_b200 import synth
from texocr_b200.model import create_model
from texocr_b200.spec import default_config

Bs = [int(x) for x in sys.argv[1].split(",")]
ns = [int(x) for x in sys.argv[2].split(",")]
T = int(sys.argv[3]) if len(sys.argv) > 3 else 256
calls = int(sys.argv[4]) if len(sys.argv) > 4 else 3
models, sd = [], None
for i in range(max(ns)):
    m = create_model(default_config(), precision="bf16")
    if sd is None:
        sd = synth.seeded_state_dict(m.dims)
    m.load_state_dict(sd)
    m.engine().set_option("decode_branches", 1)
    models.append(m)
img = synth.synth_images(512, 64, 384, seed=21).cuda()
enc512 = models[0].encoder(img)
for B in Bs:
    enc = torch.cat([enc512] * (B // 512), 0).contiguous()
    start = torch.full((B, 1), models[0].dims.bos, dtype=torch.long, device="cuda")
    for n in ns:
        bar = threading.Barrier(n + 1)

        def work(i):
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                models[i].decoder.generate(start_tokens=start, eos_tok=None, max_len=T, enc=enc)
                st.synchronize()
                bar.wait()
                for _ in range(calls):
                    models[i].decoder.generate(start_tokens=start, eos_tok=None, max_len=T, enc=enc)
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
        print(f"decode only: B={B} rows per launch, {n} in flight: {dt * 1e3 / (n * calls) * 512 / B:.2f} ms per 512 rows -> {B * n * calls / dt:.0f} rows/s", flush=True)
