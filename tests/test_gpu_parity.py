"""GPU tier (-m gpu): the CUDA path, called through the public API -> ctypes -> C-ABI, against the CPU oracle
and the golden vectors of the unmodified reference.  fp32 tier: 1e-4 (max|a-b|/max|b|), tokens tie-aware exact;
bf16 tier: 2e-2 (BASELINE.json north_star)."""
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


def _model(sd, precision, kind="hybrid", max_length=256):
    import texocr_b200
    cfg = spec.default_config(max_length=max_length)
    cfg["device"] = "cuda:0"
    m = texocr_b200.create_model(cfg, encoder_kind=kind, precision=precision)
    m.load_state_dict(sd)
    return m.eval()


@pytest.fixture(scope="module")
def m32(sd):
    return _model(sd, "fp32")


@pytest.fixture(scope="module")
def m16(sd):
    return _model(sd, "bf16")


def _absorb(eng, on, self_on=None):
    """bf16 generate loop: absorbed (latent) attention for cross / self-attention on or off (off = projected K/V caches)."""
    eng.set_option("cross_absorb", int(on))
    eng.set_option("self_absorb", int(on if self_on is None else self_on))


def _img(golden, name):
    B, H, W, dense, seed = [int(v) for v in golden[f"enc_{name}_shape"]]
    return synth.synth_images(B, H, W, seed=seed, dense=bool(dense))


def test_native_library_is_loaded(m32):
    eng = m32.engine()
    assert eng.lib._name.endswith("libtexocr_b200.so")
    with open("/proc/self/maps") as f:
        assert "libtexocr_b200.so" in f.read()


def test_encoder_fp32_vs_reference_golden(golden, m32):
    for name in ("a", "b", "c"):
        enc = m32.encoder(_img(golden, name).cuda()).cpu().numpy()
        if name == "c":
            enc = enc[:, ::6]
        assert rel_max(enc, golden[f"enc_{name}"]) < FP32_TOL, name


def test_backbone_tap_fp32(golden, m32, sd, O):
    img = _img(golden, "a")
    m32.encoder(img.cuda())
    feat = m32.engine().debug_read("backbone", 2 * 96 * 1024).cpu().reshape(2, 4, 24, 1024).permute(0, 3, 1, 2)
    assert rel_max(feat[:, ::16].numpy(), golden["backbone_a_sub"]) < FP32_TOL


def test_encoder_ragged_batch_equals_per_image(m32, sd, O):
    """Each image sees only its own pixels / tokens: a mixed-size batch must equal per-image runs."""
    shapes = [(64, 384), (32, 128), (160, 1008), (48, 208), (64, 384), (16, 16)]
    imgs = [synth.synth_images(1, h, w, seed=100 + i)[0] for i, (h, w) in enumerate(shapes)]
    outs = m32.encoder([im.cuda() for im in imgs])
    with torch.no_grad():
        for im, out in zip(imgs, outs):
            ref = O.encoder_forward(sd, im[None])[0]
            assert out.shape == ref.shape
            assert rel_max(out.cpu().numpy(), ref.numpy()) < FP32_TOL, im.shape
            single = m32.encoder(im[None].cuda())[0]
            assert torch.equal(single, out)          # bitwise: batch composition must not change a row's result


def test_encoder_patch_variant(golden):
    d_p = spec.dims_from_config(spec.default_config(), encoder_kind="patch")
    sd_p = synth.seeded_state_dict(d_p, seed=0)
    for prec, tol in (("fp32", FP32_TOL), ("bf16", BF16_TOL)):
        m = _model(sd_p, prec, kind="patch")
        enc = m.encoder(synth.synth_images(2, 64, 384, seed=99).cuda()).cpu().numpy()
        assert rel_max(enc, golden["enc_patch"]) < tol, prec


def test_teacher_forced_logits_and_loss_fp32(golden, m32, dims):
    img = _img(golden, "a").cuda()
    trg = torch.from_numpy(golden["tf_trg"]).cuda()
    enc = m32.encoder(img)
    loss, logits = m32.decoder(trg, enc=enc, mask=m32.make_trg_mask(trg), return_out=True)
    assert rel_max(logits.cpu().numpy(), golden["tf_logits"]) < FP32_TOL
    assert abs(float(loss) - float(golden["tf_loss"])) < 1e-4
    assert abs(float(m32(img, trg)) - float(golden["fwd_loss"])) < 1e-4
    # fully masked query rows -> uniform attention over all keys
    trg2 = torch.from_numpy(golden["tf2_trg"]).cuda()
    _, logits2 = m32.decoder(trg2, enc=enc, mask=m32.make_trg_mask(trg2), return_out=True)
    assert rel_max(logits2.cpu().numpy()[:, :, ::8], golden["tf2_logits_sub"]) < FP32_TOL
    # decoder.net called directly, no mask
    l3 = m32.decoder.net(trg[:, :-1], enc=enc)
    assert l3.shape == (2, 32, 1000)


def test_bf16_tier_encoder_and_logits(golden, m16, dims):
    img = _img(golden, "a").cuda()
    enc = m16.encoder(img)
    assert rel_max(enc.cpu().numpy(), golden["enc_a"]) < BF16_TOL
    trg = torch.from_numpy(golden["tf_trg"]).cuda()
    _, logits = m16.decoder(trg, enc=enc, mask=m16.make_trg_mask(trg), return_out=True)
    assert rel_max(logits.cpu().numpy(), golden["tf_logits"]) < BF16_TOL


def test_greedy_tokens_config1_fp32(golden, m32, dims):
    """BASELINE config 1: B=8, 64x384, max_len 256 -- tokens equal the reference's (tie-aware, tau = 1e-4)."""
    img8 = synth.synth_images(8, 64, 384, seed=1234).cuda()
    tok = m32.generate(src=img8, max_len=256)
    assert tok.dtype == torch.int64 and tok.shape == (8, 256) and tok.is_cuda
    exact, div, ok = tie_aware_rows(tok.cpu().numpy(), golden["gen8_tokens"].astype(np.int64), golden["gen8_gaps"], tau=1e-4)
    assert ok and exact >= 7, (exact, div)
    # the CUDA-graph replay and the eager launch sequence are the same computation
    m32.engine().set_option("cuda_graph", 0)
    tok2 = m32.generate(img8, 256)
    m32.engine().set_option("cuda_graph", 1)
    assert torch.equal(tok, tok2)


def test_early_exit_and_decoder_generate(golden, m32, dims):
    eos = int(golden["early_eos"])
    img8 = synth.synth_images(8, 64, 384, seed=1234).cuda()
    enc = m32.encoder(img8)
    start = torch.full((8, 1), dims.bos, dtype=torch.long, device="cuda")
    out = m32.decoder.generate(start_tokens=start, eos_tok=eos, max_len=256, temp=0.3, enc=enc)
    assert out.shape == golden["early_tokens"].shape
    assert np.array_equal(out.cpu().numpy(), golden["early_tokens"])
    full = m32.decoder.generate(start_tokens=start, eos_tok=None, max_len=40, enc=enc)
    assert full.shape == (8, 40)
    assert np.array_equal(full.cpu().numpy(), golden["gen8_tokens"][:, :40])


def test_bf16_generate_runs_and_is_deterministic(m16):
    img = synth.synth_images(16, 64, 384, seed=7).cuda()
    a = m16.generate(img, 64)
    b = m16.generate(img, 64)
    assert a.shape == (16, 64) and torch.equal(a, b)
    assert int(a.min()) >= 0 and int(a.max()) < 1000


def test_full_size_batch_properties(m32, m16, sd, O):
    """BASELINE config 3 size (B=512, 64x384): rows are independent, so the first rows of the big batch must
    equal a small-batch run bit for bit, and a host-buffer call must equal the device-buffer call."""
    img = synth.synth_images(512, 64, 384, seed=1234)
    for m in (m32, m16):
        big = m.generate(img.cuda(), 32)
        small = m.generate(img[:8].cuda(), 32)
        assert big.shape == (512, 32)
        assert torch.equal(big[:8], small)
    host_out = torch.empty((512, 32), dtype=torch.int64).pin_memory()
    res = m16.engine().generate(img.pin_memory(), 32, out=host_out).clone()
    assert torch.equal(res, big.cpu())
    # run-to-run reproducibility of the concurrent, PDL-chained decode branches (bit-identical token ids)
    dev = img.cuda()
    for _ in range(8):
        assert torch.equal(m16.generate(dev, 32), big)


def test_decode_branches_do_not_change_tokens(m16, golden):
    """The batch is cut into concurrently running sub-batches (rows are independent): any split gives the same ids,
    and the early-exit contract holds across branches."""
    img = synth.synth_images(96, 64, 384, seed=11).cuda()
    eng = m16.engine()
    outs = []
    for nb in (1, 2, 4, 0):
        eng.set_option("decode_branches", nb)
        outs.append(m16.generate(img, 48))
    eng.set_option("decode_branches", 0)
    for o in outs[1:]:
        assert torch.equal(outs[0], o)
    # early exit across independent branches: same n_steps and tokens as a single branch
    enc = m16.encoder(img)
    start = torch.full((96, 1), m16.dims.bos, dtype=torch.long, device="cuda")
    eos = int(outs[0][0, 20])
    res = []
    for nb in (1, 4):
        eng.set_option("decode_branches", nb)
        res.append(m16.decoder.generate(start_tokens=start, eos_tok=eos, max_len=48, enc=enc))
    eng.set_option("decode_branches", 0)
    assert res[0].shape == res[1].shape and torch.equal(res[0], res[1])


def test_pipeline_of_batches_in_flight_matches_generate(m16):
    """GeneratePipeline (several batches decoded concurrently by replica handles on their own host threads / streams):
    every batch gets exactly the tokens model.generate gives it, for device and for host (pinned) buffers."""
    from texocr_b200.pipeline import GeneratePipeline
    batches = [synth.synth_images(40 + 8 * i, 32, 128 + 64 * (i % 2), seed=30 + i) for i in range(5)]
    ref = [m16.generate(b.cuda(), 24).cpu() for b in batches]
    with GeneratePipeline(m16, in_flight=3, branches=2) as pipe:
        pipe.warm_up(batches[0].cuda(), 24)
        got = [t.cpu() for t in pipe.generate_batches([b.cuda() for b in batches], 24)]
        outs = [torch.empty((b.shape[0], 24), dtype=torch.int64).pin_memory() for b in batches]
        got_h = list(pipe.generate_batches([b.pin_memory() for b in batches], 24, outs=outs))
        assert pipe.kernel_launches() > 0
    for r, g, gh, o in zip(ref, got, got_h, outs):
        assert torch.equal(r, g)
        assert not gh.is_cuda and torch.equal(r, gh) and gh.data_ptr() == o.data_ptr()


def test_tma_attention_matches_simple_kernel(m16):
    """The persistent TMA-fed decode attention and the simple per-sequence kernel compute the same attention."""
    img = synth.synth_images(40, 64, 384, seed=21).cuda()
    eng = m16.engine()
    enc = m16.encoder(img)
    trg = synth.synth_labels(40, 40, m16.dims, seed=5).cuda()
    outs = []
    _absorb(eng, 0)         # compare the two kernels on the same (projected K/V) formulation
    for tma in (1, 0):
        eng.set_option("tma_attention", tma)
        outs.append(m16.generate(img, 40))
    eng.set_option("tma_attention", 1)
    _absorb(eng, 1)
    same = (outs[0] == outs[1]).float().mean().item()
    assert same > 0.90, same          # same math up to bf16 rounding of the softmax weights: only near-ties may flip
    # teacher-forced logits of the generated prefix agree with the decode loop's choices (KV cache == full recompute)
    ids = torch.cat((torch.full((40, 1), m16.dims.bos, device="cuda"), outs[0][:, :-1]), 1)
    logits = m16.decoder.net(ids, enc=enc)
    agree = (logits.argmax(-1) == outs[0]).float().mean().item()
    assert agree > 0.97, agree


def test_absorbed_attention_matches_projected_kv(m16, m32):
    """bf16 generate loop: attention with the K / V projections absorbed into the query / output projections (the decode loop
    streams 256-wide latent rows -- encoder memory / cached LayerNorm'd inputs -- instead of per-head K/V) is the same function:
    first-step logits within the bf16 tolerance of the fp32 parity tier, the same tokens as the projected-K/V path up to
    near-ties, and agreement with the teacher-forced decoder on its own prefix."""
    widths = synth.synth_widths(48, seed=3)
    src = [synth.synth_images(1, 64, int(w), seed=700 + i)[0].cuda() for i, w in enumerate(widths)]      # ragged memory lengths
    eng = m16.engine()
    B, V = len(src), m16.dims.vocab

    def first_logits(model):
        model.engine().set_option("keep_logits", 1)        # the greedy bf16 loop keeps no logits unless asked to (fused vocab GEMM + argmax)
        try:
            model.generate(src, 1)
            return model.engine().debug_read("logits", B * V).reshape(B, V).cpu()
        finally:
            model.engine().set_option("keep_logits", 0)

    ref32 = first_logits(m32)
    modes = ((1, 1), (0, 0), (1, 0), (0, 1))          # (cross, self)
    try:
        lg, toks = {}, {}
        for mode in modes:
            _absorb(eng, mode[0], mode[1])
            lg[mode] = first_logits(m16)
            toks[mode] = m16.generate(src, 40)
            assert torch.equal(toks[mode], m16.generate(src, 40)), mode          # deterministic
    finally:
        _absorb(eng, 1)
    assert not torch.equal(lg[(0, 0)], lg[(1, 1)])                               # the formulations really differ in rounding
    enc = m16.encoder(src)                                  # ragged: list of (N_i, 256)
    bos = torch.full((B, 1), m16.dims.bos, device="cuda")
    for mode in modes:
        assert rel_max(lg[mode].numpy(), ref32.numpy()) < BF16_TOL, mode
        # greedy decoding of random-init weights is chaotic (one flipped near-tie changes the rest of the row), so the formulations
        # are compared through the teacher-forced decoder on each one's own prefix, and only loosely token by token
        ids = torch.cat((bos, toks[mode][:, :-1]), 1)
        agree = (m16.decoder.net(ids, enc=enc).argmax(-1) == toks[mode]).float().mean().item()
        assert agree > 0.97, (mode, agree)
        assert (toks[mode][:, :8] == toks[(0, 0)][:, :8]).float().mean().item() > 0.9, mode


def test_input_validation_raises(m32):
    with pytest.raises(RuntimeError, match="multiples of 16"):
        m32.generate(torch.zeros(1, 1, 60, 384, device="cuda"), 8)
    with pytest.raises(RuntimeError, match="1008"):
        m32.encoder(torch.zeros(1, 1, 64, 1024, device="cuda"))
    with pytest.raises(RuntimeError, match="max_length"):
        m32.generate(torch.zeros(1, 1, 64, 384, device="cuda"), 257)


def test_packed_blob_and_resized_positional_table(tmp_path, sd, dims, m16):
    """the bf16 tier built from a 'mixed' blob is bit-identical to the one built from the fp32 state dict;
    a checkpoint with a longer positional table (model/ocr_model.py:82-90) lifts the max_len limit accordingly."""
    import texocr_b200
    from texocr_b200 import checkpoint
    path = str(tmp_path / "w.bin")
    checkpoint.save_blob(sd, dims, path, dtype="mixed")
    back, d2 = checkpoint.load_blob(path)
    cfg = spec.default_config(max_length=d2.max_length)
    cfg["device"] = "cuda:0"
    mb = texocr_b200.create_model(cfg, precision="bf16")
    mb.load_state_dict(back)
    img = synth.synth_images(16, 64, 384, seed=77).cuda()
    assert torch.equal(mb.generate(img, 48), m16.generate(img, 48))
    sd2 = dict(sd)
    gen = torch.Generator().manual_seed(3)
    sd2[checkpoint.POS_KEY] = torch.cat((sd[checkpoint.POS_KEY], torch.randn(44, 256, generator=gen) * 0.02), 0)     # 300 rows
    checkpoint.load_state_dict_resizing(mb, sd2)
    out = mb.generate(img, 290)
    assert out.shape == (16, 290) and torch.equal(out[:, :48], m16.generate(img, 48))
    with pytest.raises(RuntimeError, match="max_length"):
        m16.generate(img, 290)


def test_preprocess_u8_bit_exact_with_torchvision_pipeline(m32):
    """texocr_preprocess_u8 == ToTensor -> Grayscale(1) -> Invert of the reference (golden from torchvision),
    bit for bit, for a ragged RGB / L batch, host and device inputs, with the zero padding to the 16-pixel grid."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_prep_v1.npz"))
    imgs = [torch.from_numpy(g[f"img{i}_u8"]) for i in range(4)]
    refs = [torch.from_numpy(g[f"img{i}_f32"]) for i in range(4)]
    eng = m32.engine()
    # second order: the W % 4 == 0 images first, 4-byte aligned in the packed buffer -> the 4-pixels-per-thread path
    order = [1, 2, 3, 0]
    for src, refs in ((imgs, refs), ([t.cuda() for t in imgs], refs), ([imgs[i].cuda() for i in order], [refs[i] for i in order])):
        exact = eng.preprocess_u8(src, pad_multiple=1)
        for o, r in zip(exact, refs):
            assert o.shape == r.shape and torch.equal(o.cpu(), r)
        padded = eng.preprocess_u8(src, pad_multiple=16)
        for o, r in zip(padded, refs):
            H, W = r.shape[1:]
            assert o.shape[1] % 16 == 0 and o.shape[2] % 16 == 0 and o.shape[1] - H < 16 and o.shape[2] - W < 16
            assert torch.equal(o[:, :H, :W].cpu(), r)
            assert float(o[:, H:, :].abs().sum()) == 0.0 and float(o[:, :, W:].abs().sum()) == 0.0
    # straight into the path: a ragged batch of preprocessed images decodes
    out = m32.generate(eng.preprocess_u8(imgs, 16), 8)
    assert out.shape == (4, 8)
    with pytest.raises(RuntimeError, match="channels"):
        eng.preprocess_u8([torch.zeros(16, 16, 4, dtype=torch.uint8)])


def test_wrapper_image_to_latex_end_to_end(tmp_path, sd, golden):
    """TeXOCRWrapper (model/ocr_model.py:69-110): tokenizer file + checkpoint file + uint8 image -> (tokens, LaTeX)."""
    import json, os
    from texocr_b200.wrapper import TeXOCRWrapper
    from texocr_b200.detok import Detokenizer
    here = os.path.dirname(os.path.abspath(__file__))
    tk = json.load(open(os.path.join(here, "golden", "golden_tokenizer_v1.json")))
    merges = {(a, b): t for a, b, t in tk["bp_merges"]}
    (tmp_path / "tok.txt").write_text(f"{tk['vocab_size']}\n{tk['special_tokens']}\n{merges}\n")
    torch.save(sd, tmp_path / "model.pth")
    cfg = spec.default_config(max_length=256)
    cfg.update(device="cuda:0", tokenizer_path=str(tmp_path / "tok.txt"), model_path=str(tmp_path / "model.pth"))
    w = TeXOCRWrapper(cfg, precision="fp32")
    # the float images of the golden greedy run, quantised back to the uint8 the transform would have produced them from
    img_f = synth.synth_images(8, 64, 384, seed=1234)
    u8 = torch.round((1.0 - img_f[:, 0]) * 255.0).clamp(0, 255).to(torch.uint8)
    back = w.model.engine().preprocess_u8([u for u in u8], 16)
    assert max(float((b.cpu() - f).abs().max()) for b, f in zip(back, img_f)) <= 0.5 / 255 + 1e-6
    tokens, texts = w.batch([u for u in u8], max_len=24)
    assert tokens.shape[0] == 8 and len(texts) == 8 and all(isinstance(t, str) for t in texts)
    d = Detokenizer.load(str(tmp_path / "tok.txt"))
    assert texts == d.decode_batch(tokens.cpu(), eos_token=997)
    one_tokens, one_text = w(u8[0].numpy(), max_len=24)
    assert one_text == texts[0] and one_tokens == [t for t in tokens[0].tolist()][: len(one_tokens)]
    s_tokens, s_text = w(u8[0], max_len=24, sample=True, seed=4)
    assert isinstance(s_text, str) and s_tokens != one_tokens


def test_tma_im2col_convolutions_bit_identical_to_explicit_im2col(m16):
    """bf16 tier: the 3x3 / strided backbone convolutions run as implicit GEMMs -- same-size batches fetch their A tiles with TMA
    im2col loads, ragged batches gather them with cp.async in the GEMM's producer warps (stride-1 SAME, stride-2 TF-SAME (0,1)
    padding, 1x1 stride-2 subsampling, M tails).  Same operand values in the same k order as the explicit im2col buffer -> identical
    bits from all three; a ragged batch containing the same images gives the same rows."""
    eng = m16.engine()
    try:
        for (B, H, W) in ((2, 64, 384), (3, 48, 208), (1, 160, 1008), (5, 16, 16), (40, 64, 384)):
            img = synth.synth_images(B, H, W, seed=5 + B).cuda()
            outs = []
            for tma, gather in ((0, 0), (0, 1), (1, 1)):
                eng.set_option("im2col_tma", tma)
                eng.set_option("conv_gather", gather)
                outs.append(m16.encoder(img))
            assert torch.isfinite(outs[1]).all(), (B, H, W)
            assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]), (B, H, W)
        eng.set_option("im2col_tma", 1)
        a = synth.synth_images(2, 64, 384, seed=3).cuda()
        b = synth.synth_images(1, 32, 128, seed=4).cuda()
        c = synth.synth_images(1, 160, 1008, seed=6).cuda()
        uni = m16.encoder(a)                                              # TMA im2col path
        n = uni.shape[1]
        for gather in (0, 1):                                             # ragged -> explicit im2col buffer / gathered A tiles
            eng.set_option("conv_gather", gather)
            rag = eng.encode_packed([a[0], b[0], c[0], a[1]])
            assert torch.equal(rag[0][:n], uni[0]) and torch.equal(rag[0][-n:], uni[1]), gather
    finally:
        eng.set_option("im2col_tma", 1)
        eng.set_option("conv_gather", 1)


def test_groupnorm_partials_from_gemm_epilogue_bit_identical_to_block_kernel(m16):
    """bf16 tier: the GroupNorm statistics come from per-32-row-block partial sums with one summation order (csrc/gn_block.cuh),
    produced either by the epilogue of the convolution GEMM (same-size batches, rows per image a multiple of 32 at that level)
    or by the stand-alone block kernel (everything else).  Same bits from both, for 4 and 8 epilogue warps, with and without the
    128 x 256 tiles of the wide convolutions (same k order per output element), and a ragged batch containing the same images
    gives the same rows."""
    eng = m16.engine()
    try:
        for (B, H, W) in ((3, 64, 384), (2, 160, 1008), (2, 48, 208), (4, 32, 128), (200, 64, 384)):
            img = synth.synth_images(B, H, W, seed=11 + B).cuda()
            outs = []
            for fused, warps, bn256 in ((0, 8, 0), (1, 8, 1), (1, 4, 1), (1, 8, 0), (0, 4, 1)):
                eng.set_option("gn_fused", fused)
                eng.set_option("gemm_epi_warps", warps)
                eng.set_option("gemm_bn256", bn256)
                outs.append(m16.encoder(img))
            assert torch.isfinite(outs[0]).all(), (B, H, W)
            for o in outs[1:]:
                assert torch.equal(outs[0], o), (B, H, W)
        eng.set_option("gn_fused", 1)
        eng.set_option("gemm_epi_warps", 8)
        eng.set_option("gemm_bn256", 1)
        a = synth.synth_images(2, 64, 384, seed=3).cuda()
        b = synth.synth_images(1, 48, 208, seed=4).cuda()
        uni = m16.encoder(a)                                              # fused partials
        rag = eng.encode_packed([a[0], b[0], a[1]])                       # ragged, 48x208 has no multiple of 32 rows -> block kernel
        n = uni.shape[1]
        assert torch.equal(rag[0][:n], uni[0]) and torch.equal(rag[0][-n:], uni[1])
        # ragged batch whose images all have a multiple of 32 rows at every level: partials from the epilogue, image looked up per row block
        c = synth.synth_images(1, 64, 128, seed=8).cuda()
        d = synth.synth_images(1, 32, 256, seed=9).cuda()
        rag2 = eng.encode_packed([c[0], a[0], d[0], a[1]])
        nc, nd = m16.encoder(c).shape[1], m16.encoder(d).shape[1]
        assert torch.equal(rag2[0][:nc], m16.encoder(c)[0]) and torch.equal(rag2[0][nc:nc + n], uni[0])
        assert torch.equal(rag2[0][nc + n:nc + n + nd], m16.encoder(d)[0]) and torch.equal(rag2[0][-n:], uni[1])
        eng.set_option("gn_fused", 0)
        assert torch.equal(eng.encode_packed([c[0], a[0], d[0], a[1]])[0], rag2[0])
    finally:
        eng.set_option("gn_fused", 1)
        eng.set_option("gemm_epi_warps", 8)
        eng.set_option("gemm_bn256", 1)
