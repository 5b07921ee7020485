"""GPU tier: the decode-step attention kernels (persistent TMA + mma.sync kernel, simple per-sequence kernel) against a
plain PyTorch fp32 reference of the same op: softmax(q.K^T * 0.125).V over the cached keys (+ this step's key)."""
import pytest
import torch

from texocr_b200 import spec, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(sd):
    import texocr_b200
    cfg = spec.default_config()
    cfg["device"] = "cuda:0"
    m = texocr_b200.create_model(cfg, precision="bf16")
    m.load_state_dict(sd)
    return m.engine()


def _ref(q, K, V):
    """q [8,64]; K, V [n, 8, 64] (fp32) -> [512]"""
    s = torch.einsum("hd,nhd->hn", q, K) * 0.125
    return torch.einsum("hn,nhd->hd", torch.softmax(s, dim=-1), V).reshape(-1)


@pytest.mark.parametrize("use_tma", [True, False])
@pytest.mark.parametrize("B", [37, 300])
@pytest.mark.parametrize("t", [0, 1, 3, 4, 5, 12, 15, 16, 17, 31, 33, 100, 255])
def test_self_attention_decode(eng, t, use_tma, B):
    tcap = 256
    g = torch.Generator(device="cuda").manual_seed(t + 1)
    cache = torch.randn(B * 8 * tcap, 128, device="cuda", generator=g).to(torch.bfloat16)     # [b][head][key][K 64 | V 64]
    cache.view(B, 8, tcap, 128)[:, :, t:] = float("nan")      # rows not yet written must never matter
    qkv = torch.randn(B, 1536, device="cuda", generator=g).to(torch.bfloat16)
    step = torch.tensor([t], dtype=torch.int32, device="cuda")
    out = eng.debug_attn_decode(True, qkv[:, :512], qkv[:, 512:1024], qkv[:, 1024:], cache, 0, tcap, None, step, B, tcap, use_tma)
    torch.cuda.synchronize()
    c = cache.view(B, 8, tcap, 128).float()
    assert torch.equal(cache.view(B, 8, tcap, 128)[:, :, t, :64], qkv[:, 512:1024].view(B, 8, 64))     # appended key
    assert torch.equal(cache.view(B, 8, tcap, 128)[:, :, t, 64:], qkv[:, 1024:].view(B, 8, 64))        # appended value
    for b in range(B):
        K = c[b, :, : t + 1, :64].permute(1, 0, 2)
        V = c[b, :, : t + 1, 64:].permute(1, 0, 2)
        ref = _ref(qkv[b, :512].float().reshape(8, 64), K, V)
        err = (out[b].float() - ref).abs().max() / ref.abs().max()
        assert torch.isfinite(out[b].float()).all() and err < 2e-2, (b, float(err))
    out2 = eng.debug_attn_decode(True, qkv[:, :512], qkv[:, 512:1024], qkv[:, 1024:], cache, 0, tcap, None, step, B, tcap, use_tma)
    assert torch.equal(out, out2)            # run-to-run reproducible


@pytest.mark.parametrize("use_tma", [True, False])
def test_cross_attention_decode_ragged(eng, use_tma):
    lens = [17, 97, 100, 253, 1, 16, 33, 631, 97, 97, 64] * 14
    B, L = len(lens), 4
    off = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
    ntok = int(off[-1])
    g = torch.Generator(device="cuda").manual_seed(7)
    kv_all = torch.randn(L, 8 * ntok, 128, device="cuda", generator=g).to(torch.bfloat16)    # [layer][head][token][K 64 | V 64]
    q = torch.randn(B, 512, device="cuda", generator=g).to(torch.bfloat16)
    for layer in (0, 3):
        kv = kv_all[layer]
        out = eng.debug_attn_decode(False, q, None, None, kv, 0, 0, off, None, B, max(lens), use_tma)
        torch.cuda.synchronize()
        kvf = kv.view(8, ntok, 128).float()
        for b in range(0, B, 3):
            rows = kvf[:, int(off[b]):int(off[b + 1])]
            ref = _ref(q[b].float().reshape(8, 64), rows[:, :, :64].permute(1, 0, 2), rows[:, :, 64:].permute(1, 0, 2))
            err = (out[b].float() - ref).abs().max() / ref.abs().max()
            assert err < 2e-2, (layer, b, lens[b], float(err))


def test_cross_attention_absorbed_ragged(eng):
    """attn_abs_kernel<cross>: per head softmax(q'_h . enc^T * 0.125) . enc over the sequence's own memory tokens, all 8 heads
    sharing the [S, 256] bf16 memory; ragged lengths incl. 1 token, non-multiples of 4 / 16 and the 631-token maximum."""
    lens = [17, 97, 100, 253, 1, 16, 33, 631, 97, 97, 64, 2, 3, 15] * 12
    B = len(lens)
    off = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
    ntok = int(off[-1])
    g = torch.Generator(device="cuda").manual_seed(9)
    enc = torch.randn(ntok, 256, device="cuda", generator=g).to(torch.bfloat16)
    q = (torch.randn(B, 2048, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    out = eng.debug_attn_abs(q, enc, k_off=off)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    for b in range(B):
        mem = enc[int(off[b]):int(off[b + 1])].float()
        s = (q[b].float().reshape(8, 256) @ mem.T) * 0.125
        ref = (torch.softmax(s, dim=-1) @ mem).reshape(-1)
        err = (out[b].float() - ref).abs().max() / ref.abs().max()
        assert err < 2e-2, (b, lens[b], float(err))
    assert torch.equal(out, eng.debug_attn_abs(q, enc, k_off=off))


@pytest.mark.parametrize("B", [37, 300])
@pytest.mark.parametrize("t", [0, 1, 3, 4, 5, 12, 15, 16, 17, 31, 32, 33, 100, 255])
def test_self_attention_absorbed(eng, t, B):
    """attn_abs_kernel<self>: keys = the t cached latent rows + this step's own row (appended to the cache by the kernel)."""
    tcap = 256
    g = torch.Generator(device="cuda").manual_seed(100 + t)
    cache = torch.randn(B * tcap, 256, device="cuda", generator=g).to(torch.bfloat16)
    cache.view(B, tcap, 256)[:, t:] = float("nan")            # rows not yet written must never matter
    znew = torch.randn(B, 256, device="cuda", generator=g).to(torch.bfloat16)
    q = (torch.randn(B, 2048, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    step = torch.tensor([t], dtype=torch.int32, device="cuda")
    out = eng.debug_attn_abs(q, cache, znew=znew, tcap=tcap, step=step)
    torch.cuda.synchronize()
    assert torch.equal(cache.view(B, tcap, 256)[:, t], znew)                      # appended
    if t + 1 < tcap:
        assert torch.isnan(cache.view(B, tcap, 256)[:, t + 1:].float()).all()     # nothing else touched
    z = cache.view(B, tcap, 256)[:, : t + 1].float()
    s = torch.einsum("bhc,bjc->bhj", q.float().view(B, 8, 256), z) * 0.125
    ref = torch.einsum("bhj,bjc->bhc", torch.softmax(s, dim=-1), z).reshape(B, 2048)
    err = (out.float() - ref).abs().amax(1) / ref.abs().amax(1)
    assert torch.isfinite(out.float()).all() and float(err.max()) < 2e-2, float(err.max())
    cache.view(B, tcap, 256)[:, t] = float("nan")
    assert torch.equal(out, eng.debug_attn_abs(q, cache, znew=znew, tcap=tcap, step=step))      # run-to-run reproducible


def test_absorbed_attention_several_units_per_cta(eng):
    """More sequences than resident CTAs: a CTA walks several units (ring and score-exchange buffers carry over between them)."""
    lens = [17, 97, 33, 16, 1, 40, 97] * 260                     # 1820 sequences
    B = len(lens)
    off = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(19)
    enc = torch.randn(int(off[-1]), 256, device="cuda", generator=g).to(torch.bfloat16)
    q = (torch.randn(B, 2048, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    out = eng.debug_attn_abs(q, enc, k_off=off)
    for _ in range(3):
        assert torch.equal(out, eng.debug_attn_abs(q, enc, k_off=off))
    for b in list(range(0, B, 37)) + [B - 1]:
        mem = enc[int(off[b]):int(off[b + 1])].float()
        ref = (torch.softmax((q[b].float().reshape(8, 256) @ mem.T) * 0.125, dim=-1) @ mem).reshape(-1)
        assert (out[b].float() - ref).abs().max() / ref.abs().max() < 2e-2, b
    # self-attention, 1500 sequences, odd stage counts
    tcap, Bs = 64, 1500
    for t in (5, 37, 48):
        cache = torch.randn(Bs * tcap, 256, device="cuda", generator=g).to(torch.bfloat16)
        znew = torch.randn(Bs, 256, device="cuda", generator=g).to(torch.bfloat16)
        qs = (torch.randn(Bs, 2048, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
        step = torch.tensor([t], dtype=torch.int32, device="cuda")
        o1 = eng.debug_attn_abs(qs, cache, znew=znew, tcap=tcap, step=step)
        z = cache.view(Bs, tcap, 256)[:, : t + 1].float()
        sc = torch.einsum("bhc,bjc->bhj", qs.float().view(Bs, 8, 256), z) * 0.125
        ref = torch.einsum("bhj,bjc->bhc", torch.softmax(sc, dim=-1), z).reshape(Bs, 2048)
        err = (o1.float() - ref).abs().amax(1) / ref.abs().amax(1)
        assert float(err.max()) < 2e-2, (t, float(err.max()))
        assert torch.equal(o1, eng.debug_attn_abs(qs, cache, znew=znew, tcap=tcap, step=step))
