"""GPU tier: the tcgen05/TMEM/TMA GEMM (csrc/tc_gemm.cu) and the FFMA GEMM (csrc/gemm_simt.cu) against a plain
PyTorch reference of the same op on the same operands (fp32 math on the bf16-rounded inputs)."""
import pytest
import torch

from texocr_b200 import spec, synth

pytestmark = pytest.mark.gpu

EPI_STORE, EPI_GLU_RES, EPI_GEGLU, EPI_BIAS_RES, EPI_ARGMAX = 0, 1, 2, 3, 4


@pytest.fixture(scope="module")
def eng(sd):
    import texocr_b200
    cfg = spec.default_config()
    cfg["device"] = "cuda:0"
    m = texocr_b200.create_model(cfg, precision="bf16")
    m.load_state_dict(sd)
    return m.engine()


def _ref(A, W, epi, bias, res):
    acc = A.float() @ W.float().t()
    if bias is not None:
        acc = acc + bias
    if epi == EPI_STORE:
        return acc
    if epi == EPI_BIAS_RES:
        return acc + res
    a, g = acc[:, 0::2], acc[:, 1::2]             # interleaved (value, gate) columns
    if epi == EPI_GLU_RES:
        return a * torch.sigmoid(g) + res
    return a * torch.nn.functional.gelu(g)


CASES = [
    (512, 1536, 256, EPI_STORE, torch.bfloat16, False),
    (512, 512, 256, EPI_STORE, torch.bfloat16, False),
    (8, 512, 512, EPI_GLU_RES, torch.float32, True),
    (512, 512, 512, EPI_GLU_RES, torch.float32, True),
    (512, 2048, 256, EPI_GEGLU, torch.bfloat16, True),
    (512, 256, 1024, EPI_BIAS_RES, torch.float32, True),      # decode MLP-out
    (86, 256, 1024, EPI_BIAS_RES, torch.float32, True),
    (86, 512, 512, EPI_GLU_RES, torch.float32, True),         # decode out-projection, M tail
    (512, 1000, 256, EPI_STORE, torch.float32, True),       # vocab projection: N tail inside a tile
    (300, 4096, 256, EPI_STORE, torch.bfloat16, False),     # M tail
    (49664, 1536, 256, EPI_STORE, torch.bfloat16, False),   # encoder-sized
    (49152, 256, 1024, EPI_STORE, torch.float32, True),     # patch projection
    (20000, 2048, 256, EPI_GEGLU, torch.bfloat16, True),
    (40000, 512, 512, EPI_GLU_RES, torch.float32, True),      # encoder-sized with a residual: persistent kernel, residual fetched a chunk ahead
    (39990, 256, 1024, EPI_BIAS_RES, torch.float32, True),    # ... with an M tail
]


@pytest.mark.parametrize("M,N,K,epi,out_dtype,use_bias", CASES)
@pytest.mark.parametrize("use_tc", [True, False])
def test_bf16_gemm_vs_torch(eng, M, N, K, epi, out_dtype, use_bias, use_tc):
    _gemm_case(eng, M, N, K, epi, out_dtype, use_bias, use_tc)


def _gemm_case(eng, M, N, K, epi, out_dtype, use_bias, use_tc):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K + epi)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.06).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g) * 0.1 if use_bias else None
    n_out = N // 2 if epi in (EPI_GLU_RES, EPI_GEGLU) else N
    res = torch.randn(M, n_out, device="cuda", generator=g) if epi in (EPI_GLU_RES, EPI_BIAS_RES) else None
    C = torch.full((M, n_out), float("nan"), device="cuda", dtype=out_dtype)
    eng.debug_gemm(A, W, C, epi=epi, bias=bias, res=res, use_tc=use_tc)
    torch.cuda.synchronize()
    ref = _ref(A, W, epi, bias, res)
    tol = 1e-2 if out_dtype == torch.bfloat16 else 2e-5
    err = (C.float() - ref).abs().max() / ref.abs().max()
    assert torch.isfinite(C.float()).all()
    assert err < tol, float(err)


@pytest.mark.parametrize("M,N,K", [(1536, 256, 64), (384, 512, 256), (96, 1024, 256), (12288, 128, 1152), (49152, 64, 576)])
def test_split_bf16x3_gemm_carries_fp32_operands(eng, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g)
    hi = lambda t: t.to(torch.bfloat16)
    lo = lambda t: (t - t.to(torch.bfloat16).float()).to(torch.bfloat16)
    C = torch.full((M, N), float("nan"), device="cuda")
    eng.debug_gemm(hi(A), hi(W), C, epi=EPI_STORE, use_tc=True, A2=lo(A), W2=lo(W))
    torch.cuda.synchronize()
    ref = (A.double() @ W.double().t()).float()
    err = (C - ref).abs().max() / ref.abs().max()
    assert err < 5e-5, float(err)      # single-pass bf16 would sit near 4e-3


def test_fp32_ffma_gemm(eng):
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.randn(777, 512, device="cuda", generator=g)
    W = torch.randn(512, 512, device="cuda", generator=g) * 0.05
    bias = torch.randn(512, device="cuda", generator=g)
    res = torch.randn(777, 256, device="cuda", generator=g)
    C = torch.empty(777, 256, device="cuda")
    eng.debug_gemm(A, W, C, epi=EPI_GLU_RES, bias=bias, res=res, use_tc=False)
    ref = _ref(A.double(), W.double(), EPI_GLU_RES, bias.double(), res.double()).float()
    assert (C - ref).abs().max() / ref.abs().max() < 1e-5


@pytest.mark.parametrize("M,N", [(512, 1000), (86, 1000), (300, 64), (129, 1024)])
def test_vocab_gemm_with_argmax_epilogue(eng, M, N):
    """EPI_ARGMAX (the decode step's vocabulary projection, model/decoder.py:60,103): every 32-column tile is reduced to
    (max, first index of the max) of acc + bias in the GEMM epilogue; the logits themselves are never stored.  Checked against
    torch on the same operands: per-tile maxima within fp32 GEMM tolerance, indices equal to torch's argmax of the reference
    wherever the top-2 gap inside the tile is not a rounding tie, and exact lowest-index tie-breaking on duplicated columns."""
    K = 256
    g = torch.Generator(device="cuda").manual_seed(M + N)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.06).to(torch.bfloat16)
    W[N // 2 + 1] = W[N // 2]                                   # two identical columns inside one tile: the first must win
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    bias[N // 2 + 1] = bias[N // 2]
    nt = (N + 31) // 32
    P = torch.full((M, nt, 2), float("nan"), device="cuda")
    eng.debug_gemm(A, W, P, epi=EPI_ARGMAX, bias=bias, use_tc=True, ldc=nt)
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t() + bias
    pad = torch.full((M, nt * 32 - N), float("-inf"), device="cuda")
    tiles = torch.cat((ref, pad), 1).reshape(M, nt, 32)
    top2 = tiles.topk(2, dim=-1).values
    mx, idx = P[..., 0], P[..., 1].contiguous().view(torch.int32)
    assert (mx - top2[..., 0]).abs().max() / ref.abs().max() < 2e-5
    ref_idx = tiles.argmax(-1).to(torch.int32) + 32 * torch.arange(nt, device="cuda", dtype=torch.int32)
    clear = (top2[..., 0] - top2[..., 1]) > 1e-4 * ref.abs().max()
    assert bool((idx == ref_idx)[clear].all())
    assert bool((idx >= 0).all()) and bool((idx < N).all())
    dup_tile = (N // 2) // 32
    assert not bool((idx[:, dup_tile] == N // 2 + 1).any())     # the duplicate never beats the first occurrence
    # whole-row argmax from the partials == torch argmax of the reference (up to rounding ties)
    row_best = mx.argmax(-1)
    row_idx = idx.gather(1, row_best[:, None])[:, 0]
    t2 = ref.topk(2, dim=-1).values
    ok = (t2[:, 0] - t2[:, 1]) > 1e-4 * ref.abs().max()
    assert bool((row_idx.long() == ref.argmax(-1))[ok].all())


@pytest.mark.parametrize("M,N,K,epi,out_dtype,use_bias", [CASES[0], CASES[3], CASES[4], CASES[5], CASES[8], CASES[10]])
def test_bf16_gemm_four_epilogue_warps(eng, M, N, K, epi, out_dtype, use_bias):
    """The default is two epilogue warpgroups per CTA (every other column chunk each); option gemm_epi_warps = 4 keeps one."""
    eng.set_option("gemm_epi_warps", 4)
    try:
        _gemm_case(eng, M, N, K, epi, out_dtype, use_bias, True)
    finally:
        eng.set_option("gemm_epi_warps", 8)
