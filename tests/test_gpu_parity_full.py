"""GPU tier (-m gpu), part 2: the MEASURED configurations against the CPU oracle at their real sizes.

* the bf16 generate loop (what bench.py times: absorbed attention, fused decode kernels) -- last-position logits of decode
  steps t = 1 / 17 / 100 / 255 against ``oracle.decoder_logits`` on the same prefix (2e-2, BASELINE.json north_star);
* BASELINE configs[3] at size: teacher-forced forward, B = 256, L = 257 labels, fp32 (1e-4) and bf16 (2e-2) against
  ``oracle.decoder_loss`` (model/decoder.py:124-145) -- 8 key tiles of the causal kernel, padded rows included;
* BASELINE configs[2] at size: fp32 greedy generate, B = 512 x 256 tokens, against the KV-cached oracle, tie-aware;
* bf16 encoder against the reference goldens "b" (48x208) and "c" (160x1008) and a 16-image mixed-width batch.

The oracle runs on the host cores of the GPU box (tens of seconds per test); /root/reference is never read here.
"""
import numpy as np
import pytest
import torch

from conftest import rel_max, tie_aware_rows
from texocr_b200 import spec, synth

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
BF16_TOL = 2e-2


@pytest.fixture(scope="module")
def O():
    from oracle import texocr_oracle
    return texocr_oracle


def _model(sd, precision):
    import texocr_b200
    cfg = spec.default_config(max_length=256)
    cfg["device"] = "cuda:0"
    m = texocr_b200.create_model(cfg, precision=precision)
    m.load_state_dict(sd)
    return m.eval()


@pytest.fixture(scope="module")
def m32(sd):
    return _model(sd, "fp32")


@pytest.fixture(scope="module")
def m16(sd):
    return _model(sd, "bf16")


def _img(golden, name):
    B, H, W, dense, seed = [int(v) for v in golden[f"enc_{name}_shape"]]
    return synth.synth_images(B, H, W, seed=seed, dense=bool(dense))


def test_bf16_generate_loop_logits_vs_oracle_at_depth(m16, sd, O, dims):
    """The loop bench.py measures, pinned to the oracle beyond its first step: run t + 1 greedy steps, read the logits the
    last step chose from, and compare with the oracle's last-position logits for the SAME prefix (BOS + the loop's own
    first t tokens) over the fp32 oracle encoder memory.  Also: the stored token is the lowest-index argmax of those logits."""
    B, V = 16, dims.vocab
    img = synth.synth_images(B, 64, 384, seed=4242)
    with torch.no_grad():
        enc_ref = O.encoder_forward(sd, img)
    eng = m16.engine()
    eng.set_option("keep_logits", 1)
    try:
        worst = 0.0
        for t in (1, 17, 100, 255):
            tok = m16.generate(img.cuda(), t + 1)
            assert tok.shape == (B, t + 1)
            lg = eng.debug_read("logits", B * V).reshape(B, V).cpu()
            ids = torch.cat((torch.full((B, 1), dims.bos, dtype=torch.long), tok[:, :t].cpu()), 1)
            with torch.no_grad():
                ref = O.decoder_logits(sd, ids, enc_ref)[:, -1]
            err = rel_max(lg.numpy(), ref.numpy())
            worst = max(worst, err)
            assert err < BF16_TOL, (t, err)
            assert torch.equal(lg.argmax(-1), tok[:, t].cpu()), t
        print(f"bf16 generate-loop logits vs oracle: worst rel err {worst:.2e}")
    finally:
        eng.set_option("keep_logits", 0)


def test_config4_teacher_forced_at_size(m32, m16, sd, O, dims):
    """BASELINE configs[3]: B = 256, 64x384, labels (256, 257) -- logits (256, 256, 1000) and the CE loss of both tiers
    against the oracle (model/ocr_model.py:34-44, model/decoder.py:124-145)."""
    B, L = 256, 257
    img = synth.synth_images(B, 64, 384, seed=1234)
    trg = synth.synth_labels(B, L, dims, seed=4321)
    with torch.no_grad():
        enc_ref = O.encoder_forward(sd, img)
        loss_ref, logits_ref = O.decoder_loss(sd, trg, enc_ref, mask=trg != dims.pad)
    loss_ref = float(loss_ref)
    logits_ref = logits_ref.numpy()
    for m, tol in ((m32, FP32_TOL), (m16, BF16_TOL)):
        enc = m.encoder(img.cuda())
        e_enc = rel_max(enc.cpu().numpy(), enc_ref.numpy())
        loss, logits = m.decoder(trg.cuda(), enc=enc, mask=m.make_trg_mask(trg.cuda()), return_out=True)
        assert logits.shape == (B, L - 1, dims.vocab)
        e_log = rel_max(logits.cpu().numpy(), logits_ref)
        e_loss = abs(float(loss) - loss_ref) / abs(loss_ref)
        fwd = float(m(img.cuda(), trg.cuda()))
        print(f"config 4 at size, {m.precision}: encoder {e_enc:.2e}, logits {e_log:.2e}, loss rel {e_loss:.2e}")
        assert e_enc < tol and e_log < tol and e_loss < tol, (m.precision, e_enc, e_log, e_loss)
        assert abs(fwd - loss_ref) / abs(loss_ref) < tol


def test_config3_fp32_greedy_tokens_at_size(m32, sd, O, dims):
    """BASELINE configs[2] on the fp32 parity tier: B = 512, 64x384, 256 greedy steps against the KV-cached oracle
    (token-identical to the reference loop in fp32).  Tie-aware (tau = 1e-4): a row must match exactly or
    first diverge where the oracle's own top-2 gap is below tau; at least 98 % of the rows match exactly."""
    B, T = 512, 256
    img = synth.synth_images(B, 64, 384, seed=1234)
    tok = m32.generate(img.cuda(), T).cpu().numpy()
    assert tok.shape == (B, T)
    gaps = []
    with torch.no_grad():
        enc_ref = O.encoder_forward(sd, img)
        ref = O.generate_greedy_cached(sd, enc_ref, T, dims.bos, dims.eos, gaps=gaps)
    assert ref.shape == (B, T)
    gaps = torch.stack(gaps, 1).numpy()
    exact, div, ok = tie_aware_rows(tok, ref.numpy(), gaps, tau=1e-4)
    print(f"config 3 fp32 tokens: {exact}/{B} rows exact; divergences (row, step, oracle top-2 gap): {div}")
    assert ok, div
    assert exact >= int(0.98 * B), (exact, div)


def test_bf16_encoder_more_shapes(golden, m16, sd, O):
    """bf16 tier encoder against the reference goldens of the other two shapes (non-64-multiple width; the 160x1008
    maximum) and a 16-image mixed-width batch (BASELINE configs[1] style) against the oracle, image by image."""
    for name in ("b", "c"):
        enc = m16.encoder(_img(golden, name).cuda()).cpu().numpy()
        if name == "c":
            enc = enc[:, ::6]
        err = rel_max(enc, golden[f"enc_{name}"])
        print(f"bf16 encoder vs reference golden {name}: {err:.2e}")
        assert err < BF16_TOL, (name, err)
    widths = synth.synth_widths(16, seed=9)
    imgs = [synth.synth_images(1, 64, int(w), seed=900 + i)[0] for i, w in enumerate(widths)]
    outs = m16.encoder([im.cuda() for im in imgs])
    worst = 0.0
    with torch.no_grad():
        for im, out in zip(imgs, outs):
            ref = O.encoder_forward(sd, im[None])[0]
            assert out.shape == ref.shape
            worst = max(worst, rel_max(out.cpu().numpy(), ref.numpy()))
    print(f"bf16 encoder, 16 mixed widths {sorted(set(widths))}: worst {worst:.2e}")
    assert worst < BF16_TOL, worst


def test_tensor_core_stem_vs_ffma_stem(golden, m16, sd, O):
    """bf16 tier: the stem convolution runs as a tcgen05 implicit GEMM (bf16x3: ~16 mantissa bits per operand, GroupNorm partials from
    its epilogue); option stem_tc = 0 selects the FFMA kernel of the fp32 tier + the stand-alone statistics pass.  Both within the
    bf16 tolerance of the reference golden / the oracle, and within 5e-3 of each other, for a same-size and a ragged batch; a ragged
    batch gives the rows of the per-image runs bit for bit."""
    eng = m16.engine()
    img = _img(golden, "a").cuda()
    imgs = [synth.synth_images(1, h, w, seed=40 + i)[0].cuda() for i, (h, w) in enumerate(((64, 384), (16, 16), (160, 1008), (48, 208)))]
    try:
        outs = {}
        for flag in (1, 0):
            eng.set_option("stem_tc", flag)
            outs[flag] = (m16.encoder(img).cpu(), eng.encode_packed(imgs)[0].cpu())
            assert rel_max(outs[flag][0].numpy(), golden["enc_a"]) < BF16_TOL, flag
        for a, b in zip(outs[1], outs[0]):
            assert torch.isfinite(a).all() and rel_max(a.numpy(), b.numpy()) < 5e-3
        eng.set_option("stem_tc", 1)
        off = 0
        for im in imgs:
            single = m16.encoder(im[None])[0].cpu()
            assert torch.equal(single, outs[1][1][off:off + single.shape[0]])
            ref = O.encoder_forward(sd, im[None].cpu())[0]
            assert rel_max(single.numpy(), ref.numpy()) < BF16_TOL
            off += single.shape[0]
    finally:
        eng.set_option("stem_tc", 1)


def test_many_tiny_images_in_one_ragged_batch(m16, m32):
    """Hundreds of 16x16 / 16x32 / 32x16 images in one batch: a 128-row tile of the persistent kernels spans up to 128 images here, so the
    image lookup of the producers / epilogues takes its binary-search path (more than 8 images ahead of the previous tile's) as well as
    the forward-step one.  Rows of the batch equal the per-image runs bit for bit, on both tiers."""
    shapes = [(16, 16), (16, 32), (32, 16)]
    imgs = [synth.synth_images(1, *shapes[i % 3], seed=3000 + i)[0].cuda() for i in range(331)]
    for m in (m16, m32):
        outs = m.encoder(imgs)
        assert len(outs) == len(imgs) and all(torch.isfinite(o).all() for o in outs)
        for i in (0, 1, 2, 57, 130, 329, 330):
            single = m.encoder(imgs[i][None])[0]
            assert torch.equal(single, outs[i]), i


def test_alternating_entry_points_share_no_stale_graph(m16, dims):
    """decoder.generate(enc=...) and model.generate(src) with the same batch, max_len and eos replay different captured
    graphs: the cross-attention launches bake the device pointer of the memory offsets, which the two entry points place
    differently (the graph cache key must hold it)."""
    img = synth.synth_images(24, 64, 384, seed=77).cuda()
    ref = m16.generate(img, 20)
    enc = m16.encoder(img)
    start = torch.full((24, 1), dims.bos, dtype=torch.long, device="cuda")
    for _ in range(2):
        a = m16.decoder.generate(start_tokens=start, eos_tok=dims.eos, max_len=20, enc=enc)
        b = m16.generate(img, 20)
        assert torch.equal(b, ref)
        n = min(a.shape[1], ref.shape[1])
        # decoder.generate starts from the fp32 memory handed over the API (re-rounded to bf16): same tokens up to near-ties
        assert (a[:, :n] == ref[:, :n]).float().mean().item() > 0.9
    # a ragged batch right after a rectangular one of the same size
    widths = [384] * 23 + [128]
    rag = [synth.synth_images(1, 64, w, seed=77)[0].cuda() for w in widths]
    r1 = m16.generate(rag, 20)
    assert torch.equal(m16.generate(img, 20), ref)
    assert torch.equal(m16.generate(rag, 20), r1)


def test_fused_vocab_argmax_equals_logits_path(m16):
    """The greedy bf16 loop reduces the vocabulary projection to per-tile (max, index) partials in the GEMM epilogue and never
    stores logits; with keep_logits the same step writes the logits row and scans it.  Same tokens, early-exit included."""
    img = synth.synth_images(130, 64, 384, seed=55).cuda()
    eng = m16.engine()
    fused = m16.generate(img, 48)
    eng.set_option("keep_logits", 1)
    try:
        plain = m16.generate(img, 48)
    finally:
        eng.set_option("keep_logits", 0)
    assert torch.equal(fused, plain)
    enc = m16.encoder(img)
    start = torch.full((130, 1), m16.dims.bos, dtype=torch.long, device="cuda")
    eos = int(fused[0, 12])
    a = m16.decoder.generate(start_tokens=start, eos_tok=eos, max_len=48, enc=enc)
    eng.set_option("keep_logits", 1)
    try:
        b = m16.decoder.generate(start_tokens=start, eos_tok=eos, max_len=48, enc=enc)
    finally:
        eng.set_option("keep_logits", 0)
    assert a.shape == b.shape and torch.equal(a, b)


def test_one_large_generate_call_equals_sub_batch_calls(m16):
    """model.generate with B >= 1024 decodes 512-row sub-batches concurrently on internal engine replicas: same tokens as
    separate 512-row calls, and the reference's width contract when one sub-batch reaches its all-EOS step before another."""
    img = synth.synth_images(1100, 32, 128, seed=61).cuda()
    ref = torch.cat([m16.generate(c, 24) for c in torch.split(img, 512)], 0)
    out = m16.generate(img, 24)
    assert out.shape == (1100, 24) and torch.equal(out, ref)
    # early exit: the whole batch stops at the step where its LAST row has produced the EOS
    old_eos = m16.dims.eos
    full = ref
    import dataclasses
    cand = None
    for t in full[0, :20].tolist():                      # a token every row emits early becomes the EOS
        hit = (full == t)
        if bool(hit.any(1).all()):
            steps = hit.float().argmax(1) + 1
            per_chunk = [int(s.max()) for s in torch.split(steps, 512)]
            if len(set(per_chunk)) > 1:
                cand = (t, max(per_chunk))
                break
    if cand is None:
        pytest.skip("no token reaches every row at different steps per sub-batch in this sample")
    try:
        m16.dims = dataclasses.replace(m16.dims, eos=cand[0])      # engine and replicas are rebuilt for the new EOS id
        m16.eos_token = cand[0]
        out = m16.generate(img, 24)
        assert out.shape == (1100, cand[1]) and torch.equal(out, full[:, :cand[1]])
    finally:
        m16.dims = dataclasses.replace(m16.dims, eos=old_eos)
        m16.eos_token = old_eos
