"""CPU tier: pin the oracle (oracle/texocr_oracle.py) against outputs of the unmodified reference
(tests/golden/golden_v1.npz, written by tests/golden/make_golden.py in the build container)."""
import numpy as np
import torch

from conftest import rel_max, tie_aware_rows
from oracle import texocr_oracle as O
from texocr_b200 import spec, synth

FP32_TOL = 1e-4      # BASELINE.json north_star: 1e-4 relative for fp32


def _img(golden, name):
    B, H, W, dense, seed = [int(v) for v in golden[f"enc_{name}_shape"]]
    return synth.synth_images(B, H, W, seed=seed, dense=bool(dense))


def test_state_dict_matches_reference_inventory(dims, sd):
    assert len(sd) == 372
    uniq = {id(v): v for v in sd.values()}
    assert sum(v.numel() for v in uniq.values()) == 23_651_752
    assert sd["encoder.attn_layers.layers.5.0.weight"] is sd["encoder.attn_layers.layers.0.0.weight"]


def test_encoder_hybrid_three_shapes(golden, sd):
    with torch.no_grad():
        for name in ("a", "b", "c"):
            taps = {}
            enc = O.encoder_forward(sd, _img(golden, name), taps=taps).numpy()
            if name == "c":
                enc = enc[:, ::6]
            assert rel_max(enc, golden[f"enc_{name}"]) < FP32_TOL, name
            if name == "a":
                assert rel_max(taps["backbone"][:, ::16].numpy(), golden["backbone_a_sub"]) < FP32_TOL


def test_encoder_patch_variant(golden, dims):
    cfg = spec.default_config()
    d_p = spec.dims_from_config(cfg, encoder_kind="patch")
    sd_p = synth.seeded_state_dict(d_p, seed=0)
    with torch.no_grad():
        enc = O.encoder_forward(sd_p, synth.synth_images(2, 64, 384, seed=99), kind="patch").numpy()
    assert rel_max(enc, golden["enc_patch"]) < FP32_TOL


def test_teacher_forced_logits_and_loss(golden, sd, dims):
    with torch.no_grad():
        img = _img(golden, "a")
        enc = O.encoder_forward(sd, img)
        trg = torch.from_numpy(golden["tf_trg"])
        loss, logits = O.decoder_loss(sd, trg, enc, trg != dims.pad)
        assert rel_max(logits.numpy(), golden["tf_logits"]) < FP32_TOL
        assert abs(float(loss) - float(golden["tf_loss"])) < 1e-4
        loss2, _ = O.model_forward(sd, img, trg, pad=dims.pad)
        assert abs(float(loss2) - float(golden["fwd_loss"])) < 1e-4
        # fully-masked query rows soften to the uniform average over ALL keys
        trg2 = torch.from_numpy(golden["tf2_trg"])
        _, logits2 = O.decoder_loss(sd, trg2, enc, trg2 != dims.pad)
        assert rel_max(logits2[:, :, ::8].numpy(), golden["tf2_logits_sub"]) < FP32_TOL


def test_greedy_tokens_config1_cached_and_reference_loop(golden, sd, dims):
    img8 = synth.synth_images(8, 64, 384, seed=1234)
    ref = golden["gen8_tokens"].astype(np.int64)
    with torch.no_grad():
        enc = O.encoder_forward(sd, img8)
        cached = O.generate_greedy_cached(sd, enc, 256, dims.bos, dims.eos).numpy()
        exact, div, ok = tie_aware_rows(cached, ref, golden["gen8_gaps"], tau=1e-4)
        assert ok and exact >= 7, (exact, div)
        # the reference's own O(T^2) loop, bounded to 48 steps to keep the CPU suite short
        rec = O.generate_greedy_recompute(sd, enc, 48, dims.bos, dims.eos).numpy()
        assert np.array_equal(rec, cached[:, :48])


def test_early_exit_contract(golden, sd, dims):
    """Output length = first step at which every row holds an EOS (model/decoder.py:115-116)."""
    eos = int(golden["early_eos"])
    img8 = synth.synth_images(8, 64, 384, seed=1234)
    with torch.no_grad():
        enc = O.encoder_forward(sd, img8)
        a = O.generate_greedy_cached(sd, enc, 256, dims.bos, eos).numpy()
        b = O.generate_greedy_recompute(sd, enc, 256, dims.bos, eos).numpy()
    assert a.shape == golden["early_tokens"].shape == b.shape
    assert np.array_equal(a, golden["early_tokens"]) and np.array_equal(b, a)


def test_batch_acc_matches_reference_formula():
    pred = torch.tensor([[1, 2, 3, 4, 5, 6, 7, 8], [1, 2, 3, 4, 5, 6, 7, 8]])
    target = torch.tensor([[1, 2, 3, 4, 5, 6, 7, 8], [1, 2, 3, 4, 6, 999, 999, 999]])
    assert abs(O.batch_acc(pred, target, 999) - (1.0 + 4 / 8) / 2) < 1e-6
