"""Launch the TMA decode-attention kernels concurrently on several streams, together with GEMM traffic, and check every
output against a reference."""
import sys, torch
sys.path.insert(0, "/root/repo")
import texocr_b200
from texocr_b200 import spec, synth
cfg = spec.default_config(max_length=256); cfg["device"]="cuda:0"
d = spec.dims_from_config(cfg)
m = texocr_b200.create_model(cfg, precision="bf16"); m.load_state_dict(synth.seeded_state_dict(d, seed=0))
eng = m.engine()
NS_, B = 8, 64
with_gemm = "gemm" in sys.argv
g = torch.Generator(device="cuda").manual_seed(1)
streams = [torch.cuda.Stream() for _ in range(NS_)]
ntok = 512 * 97
kv = torch.randn(8 * ntok, 128, device="cuda", generator=g).to(torch.bfloat16)
offs = [torch.arange(i * 64 * 97, (i * 64 + 65) * 97, 97, dtype=torch.int32, device="cuda") for i in range(NS_)]
qs = [torch.randn(B, 512, device="cuda", generator=g).to(torch.bfloat16) for _ in range(NS_)]
A = [(torch.randn(B, 512, device="cuda", generator=g) * 0.5).to(torch.bfloat16) for _ in range(NS_)]
W = (torch.randn(512, 512, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
res = [torch.randn(B, 256, device="cuda", generator=g) for _ in range(NS_)]
Cs = [torch.empty(B, 256, device="cuda") for _ in range(NS_)]
def run(i):
    return eng.debug_attn_decode(False, qs[i], None, None, kv, 0, 0, offs[i], None, B, 97, True)
def gemm(i):
    eng.debug_gemm(A[i], W, Cs[i], epi=1, res=res[i], use_tc=True)
refs = [run(i).clone() for i in range(NS_)]
torch.cuda.synchronize()
bad = 0
for it in range(300):
    outs = []
    for i, s in enumerate(streams):
        with torch.cuda.stream(s):
            if with_gemm: gemm(i)
            outs.append(run(i))
            if with_gemm: gemm(i)
    torch.cuda.synchronize()
    for i in range(NS_):
        if not torch.equal(outs[i], refs[i]):
            bad += 1
            if bad < 6:
                w = (outs[i].view(torch.int32) != refs[i].view(torch.int32))
                print("iter", it, "stream", i, "mismatch words", int(w.sum()), "rows", w.any(1).nonzero().flatten()[:6].tolist(), "cols", w.any(0).nonzero().flatten()[:8].tolist())
print("cross-attention concurrent launches:", 300 * NS_, "with_gemm", with_gemm, "mismatches:", bad)
