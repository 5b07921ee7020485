"""Warm single-launch latency of the decode GEMM shapes (M = 86 rows) through texocr_debug_gemm, CUDA events around 200 launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from texocr_b200 import synth
from texocr_b200.model import create_model
from texocr_b200.spec import default_config

m = create_model(default_config(), precision="bf16")
m.load_state_dict(synth.seeded_state_dict(m.dims))
eng = m.engine()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 86
shapes = [("qkv    K=256  N=1536 store", 256, 1536, 0), ("q_c    K=256  N=512  store", 256, 512, 0), ("out    K=512  N=512  glu+res", 512, 512, 1),
          ("ff1    K=256  N=2048 geglu", 256, 2048, 2), ("ff2    K=1024 N=256  bias+res", 1024, 256, 3), ("logits K=256  N=1000 store f32", 256, 1000, 0)]
for name, K, N, epi in shapes:
    A = torch.randn(M, K, device="cuda").bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    nout = N // 2 if epi in (1, 2) else N
    cdt = torch.bfloat16 if (epi == 2 or (epi == 0 and "f32" not in name)) else torch.float32
    C = torch.empty(M, nout, device="cuda", dtype=cdt)
    bias = torch.randn(N, device="cuda") if epi else None
    res = torch.randn(M, nout, device="cuda") if epi in (1, 3) else None
    for _ in range(10):
        eng.debug_gemm(A, W, C, epi=epi, bias=bias, res=res)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    e0.record()
    for _ in range(n):
        eng.debug_gemm(A, W, C, epi=epi, bias=bias, res=res)
    e1.record()
    torch.cuda.synchronize()
    print("%-34s %6.2f us per launch (back to back, PDL)" % (name, 1e3 * e0.elapsed_time(e1) / n))
