"""checkpoint ingestion (reference .pth formats, positional-table resize) and the packed weight blob."""
import os

import numpy as np
import pytest
import torch

from texocr_b200 import checkpoint, spec, synth
from texocr_b200.model import OCRModel


def _cpu_model(max_length=256):
    cfg = spec.default_config(max_length=max_length, vocab_size=1000)
    cfg["device"] = "cpu"
    return OCRModel(cfg, seed=1)


def test_load_checkpoint_both_reference_formats(tmp_path, sd):
    train_ckpt = {"epoch": 7, "model_state_dict": sd, "optimizer_state_dict": {"state": {}, "param_groups": []}}   # utils.py:50-60
    torch.save(train_ckpt, tmp_path / "checkpoint_e7.pth")
    torch.save(sd, tmp_path / "model.pth")                                                                           # ocr_model.py:78
    for name, epoch in (("checkpoint_e7.pth", 7), ("model.pth", None)):
        m = _cpu_model()
        m2, ep = checkpoint.load_checkpoint(m, str(tmp_path / name))
        assert m2 is m and ep == epoch
        got = m.state_dict()
        assert set(got) == set(sd)
        assert all(torch.equal(got[k], sd[k]) for k in sd)
    torch.save({"foo": 1}, tmp_path / "bad.pth")
    with pytest.raises(ValueError, match="not a TeXOCR checkpoint"):
        checkpoint.load_checkpoint(_cpu_model(), str(tmp_path / "bad.pth"))


def test_positional_table_resize_like_the_wrapper(sd):
    """model/ocr_model.py:82-90: a checkpoint trained with another max_length replaces the positional table."""
    sd2 = dict(sd)
    sd2[checkpoint.POS_KEY] = torch.randn(350, 256)
    m = _cpu_model(max_length=256)
    with pytest.raises(RuntimeError, match="size mismatch"):
        m.load_state_dict(sd2)
    checkpoint.load_state_dict_resizing(m, sd2)
    assert m.dims.max_length == 350 and m.decoder.max_len == 350
    assert torch.equal(m.state_dict()[checkpoint.POS_KEY], sd2[checkpoint.POS_KEY])
    assert len(m.state_dict()) == len(sd)


@pytest.mark.parametrize("dtype", ["fp32", "mixed"])
def test_blob_round_trip(tmp_path, sd, dims, dtype):
    path = str(tmp_path / f"w_{dtype}.bin")
    size = checkpoint.save_blob(sd, dims, path, dtype=dtype)
    assert size == os.path.getsize(path)
    back, d2 = checkpoint.load_blob(path)
    assert d2 == dims and set(back) == set(sd)
    n_bf16 = 0
    for k, v in sd.items():
        if dtype == "fp32" or not (v.ndim == 2 and checkpoint._is_bf16_matrix(k)):
            assert torch.equal(back[k], v), k
        else:
            n_bf16 += 1
            assert torch.equal(back[k], v.to(torch.bfloat16).to(torch.float32)), k      # exactly torch's RNE rounding
    if dtype == "mixed":
        assert n_bf16 >= 60 and size < 0.75 * 4 * sum(p.numel() for p in spec.unique_params(spec.param_table(dims)).values() for p in [torch.empty(p.shape)])
    # aliases point at the same storage after load, like the reference's shared LayerNorm modules
    table = spec.param_table(dims)
    alias = next(p for p in table if p.alias_of is not None)
    assert back[alias.key].data_ptr() == back[alias.alias_of].data_ptr()
    m = _cpu_model()
    m.load_state_dict(back)
    with open(path, "r+b") as f:
        f.write(b"XXXXXXXX")
    with pytest.raises(ValueError, match="not a texocr_b200 weight blob"):
        checkpoint.load_blob(path)


def test_bf16_bits_round_to_nearest_even():
    x = np.array([1.0, 1.00390625, 1.01171875, -2.5, 3.1415927, 1e-40, np.inf], dtype=np.float32)
    ref = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().astype(np.uint16)
    assert np.array_equal(checkpoint._to_bf16_bits(x), ref)
