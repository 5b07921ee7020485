"""Checkpoint ingestion: the reference's ``.pth`` files and a packed weight blob.

* ``load_checkpoint(model, path)`` -- utils.py:63-71 without the optimizer: accepts the training checkpoint dict
  (``{'model_state_dict': ..., 'optimizer_state_dict': ..., 'epoch': ...}``) or the bare state dict the wrapper loads
  (model/ocr_model.py:78), applies the wrapper's positional-table resize (model/ocr_model.py:82-90).
* ``save_blob`` / ``load_blob`` -- one flat little-endian file of the 224 unique parameters (aliased keys are
  re-expanded on load), so trained weights can ship without torch pickles.  ``dtype="mixed"`` stores the transformer
  matrices the bf16 tier rounds to bf16 anyway (q/k/v/out/MLP/logits weights) as bf16 and everything else (convolutions,
  norms, biases, embeddings, cls / positional tables) as fp32: the bf16 tier's engine weights are bit-identical to the
  ones built from the fp32 state dict, at 60 % of the size.  ``dtype="fp32"`` is lossless for both tiers.

Blob layout: 8-byte magic ``TXOCRW01`` | uint64 header length | UTF-8 JSON header | 64-byte aligned tensor data.
Header: ``{"dims": {...ModelDims...}, "tensors": [{"name", "shape", "dtype": "f32"|"bf16", "offset", "nbytes"}]}``.
"""
import dataclasses
import json
import struct
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import spec

MAGIC = b"TXOCRW01"
POS_KEY = "decoder.net.pos_embedding.embedding.weight"


def extract_state_dict(obj) -> Tuple[Dict[str, torch.Tensor], Optional[int]]:
    """(state_dict, epoch) from what ``torch.load`` returned for a reference checkpoint."""
    if isinstance(obj, dict) and "model_state_dict" in obj:
        return obj["model_state_dict"], obj.get("epoch")
    if isinstance(obj, dict) and obj and all(isinstance(v, torch.Tensor) for v in obj.values()):
        return obj, None
    raise ValueError("not a TeXOCR checkpoint: expected a state dict or a dict with 'model_state_dict'")


def load_state_dict_resizing(model, state_dict: Dict[str, torch.Tensor], strict: bool = True):
    """``load_state_dict`` preceded by the wrapper's positional-table resize (model/ocr_model.py:82-90)."""
    if POS_KEY in state_dict:
        model.resize_pos_embedding(int(state_dict[POS_KEY].shape[0]))
    return model.load_state_dict(state_dict, strict=strict)


def load_checkpoint(model, load_path: str, strict: bool = True):
    """utils.py:63-71 for inference: returns (model, epoch or None).  ``weights_only=True`` like the reference."""
    obj = torch.load(load_path, map_location="cpu", weights_only=True)
    sd, epoch = extract_state_dict(obj)
    load_state_dict_resizing(model, sd, strict=strict)
    return model, epoch


# ------------------------------------------------------------------------------------------------ packed blob
def _is_bf16_matrix(key: str) -> bool:
    """2-D weights the bf16 tier converts to bf16 GEMM operands (texocr_b200/csrc/engine_weights.cu, finalize_weights): the
    q / k / v / output / MLP matrices of both transformer stacks and the vocabulary projection."""
    return key == "decoder.net.to_logits.weight" or (".attn_layers.layers." in key and key.endswith(".weight"))


def _to_bf16_bits(x: np.ndarray) -> np.ndarray:
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    rounded = (u + 0x7FFF + ((u >> 16) & 1)) >> 16            # round to nearest even
    nan = np.isnan(x)
    out = rounded.astype(np.uint16)
    out[nan] = 0x7FC0
    return out


def _from_bf16_bits(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << 16).view(np.float32)


def save_blob(state_dict: Dict[str, torch.Tensor], dims: spec.ModelDims, path: str, dtype: str = "mixed") -> int:
    """Writes the unique parameters of ``state_dict``; returns the file size in bytes."""
    if dtype not in ("mixed", "fp32"):
        raise ValueError("dtype must be 'mixed' or 'fp32'")
    table = spec.param_table(dims)
    uniq = spec.unique_params(table)
    entries, chunks, off = [], [], 0
    for key, p in uniq.items():
        if key not in state_dict:
            raise KeyError(f"state dict lacks '{key}'")
        arr = state_dict[key].detach().to(torch.float32).cpu().contiguous().numpy()
        if tuple(arr.shape) != tuple(p.shape):
            raise ValueError(f"'{key}': shape {tuple(arr.shape)}, expected {tuple(p.shape)}")
        as_bf16 = dtype == "mixed" and arr.ndim == 2 and _is_bf16_matrix(key)
        raw = (_to_bf16_bits(arr) if as_bf16 else arr.astype("<f4")).tobytes()
        pad = (-off) % 64
        if pad:
            chunks.append(b"\0" * pad)
            off += pad
        entries.append({"name": key, "shape": list(arr.shape), "dtype": "bf16" if as_bf16 else "f32", "offset": off, "nbytes": len(raw)})
        chunks.append(raw)
        off += len(raw)
    header = json.dumps({"dims": dataclasses.asdict(dims), "tensors": entries}).encode("utf-8")
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<Q", len(header)))
        f.write(header)
        f.write(b"\0" * ((-(16 + len(header))) % 64))
        for c in chunks:
            f.write(c)
        return f.tell()


def load_blob(path: str) -> Tuple[Dict[str, torch.Tensor], spec.ModelDims]:
    """(state dict with every reference key incl. the aliases, ModelDims) from a ``save_blob`` file."""
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError(f"{path}: not a texocr_b200 weight blob")
        (hlen,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(hlen).decode("utf-8"))
        f.seek((16 + hlen + 63) // 64 * 64)
        data = f.read()
    dims = spec.ModelDims(**header["dims"])
    out: Dict[str, torch.Tensor] = {}
    for e in header["tensors"]:
        raw = data[e["offset"]: e["offset"] + e["nbytes"]]
        n = int(np.prod(e["shape"])) if e["shape"] else 1
        if e["dtype"] == "bf16":
            arr = _from_bf16_bits(np.frombuffer(raw, dtype="<u2", count=n))
        else:
            arr = np.frombuffer(raw, dtype="<f4", count=n)
        out[e["name"]] = torch.from_numpy(arr.copy().reshape(e["shape"]))
    for p in spec.param_table(dims):
        if p.alias_of is not None:
            out[p.key] = out[p.alias_of]
    return out, dims
