"""Seeded synthetic weights, images and labels.

There is no checkpoint and no dataset for the reference (README.md:1-102), so every
run uses random-init weights of the reference architecture and synthetic images of
the reference's input contract (float32 (B,1,H,W) in [0,1], ink = 1 on a 0 background,
``data_wrangling/dataset.py:365-371``).  All draws come from numpy's PCG64 stream so
the same seed gives the same bytes on every machine.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch

from .spec import ModelDims, param_table


def seeded_state_dict(d: ModelDims, seed: int = 0, rerandomise: bool = True, seed2: int = 123) -> Dict[str, torch.Tensor]:
    """Random-init weights with the reference's default initialisers.

    conv / linear: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (torch's kaiming_uniform(a=sqrt(5)) default);
    embeddings N(0, 0.02) (model/decoder.py:38-39, model/attention.py:27-28); cls_token, pos_embed,
    LN/GN affine at their (0 / 1,0) defaults (model/encoder.py:106-107).  With ``rerandomise`` those
    degenerate defaults are replaced (cls/pos ~ N(0,0.02); gamma += N(0,0.1); beta ~ N(0,0.1)) so that
    a parity check can see a wrong pos-id gather, a missing LayerNorm or swapped affine.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    rng2 = np.random.Generator(np.random.PCG64(seed2))
    sd: Dict[str, torch.Tensor] = {}
    for p in param_table(d):
        if p.alias_of is not None:
            sd[p.key] = sd[p.alias_of]
            continue
        if p.init in ("conv", "linear_w", "linear_b"):
            bound = 1.0 / math.sqrt(p.fan_in)
            a = rng.uniform(-bound, bound, size=p.shape)
        elif p.init == "normal02":
            a = rng.standard_normal(size=p.shape) * 0.02
        elif p.init == "ones":
            a = np.ones(p.shape)
            if rerandomise:
                a = a + rng2.standard_normal(size=p.shape) * 0.1
        elif p.init == "zeros":
            a = np.zeros(p.shape)
            if rerandomise:
                std = 0.02 if p.key in ("encoder.cls_token", "encoder.pos_embed") else 0.1
                a = rng2.standard_normal(size=p.shape) * std
        else:
            raise ValueError(p.init)
        sd[p.key] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return sd


def synth_images(batch: int, height: int, width: int, seed: int = 1234, dense: bool = False) -> torch.Tensor:
    """(B,1,H,W) float32 in [0,1]: sparse 'strokes' (8 % ink) or a dense-noise variant."""
    rng = np.random.Generator(np.random.PCG64(seed))
    val = rng.random(size=(batch, 1, height, width), dtype=np.float32)
    if not dense:
        ink = rng.random(size=(batch, 1, height, width), dtype=np.float32) < 0.08
        val = val * ink.astype(np.float32)
    return torch.from_numpy(val)


def synth_widths(batch: int, seed: int = 77, lo: int = 128, hi: int = 1008) -> List[int]:
    """Mixed widths for BASELINE config 2: multiples of 16 in [128, 1008] (cap: SURVEY.md 0.7)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return [int(16 * v) for v in rng.integers(lo // 16, hi // 16 + 1, size=batch)]


def synth_labels(batch: int, length: int, d: ModelDims, seed: int = 4321, min_len: int = 16) -> torch.Tensor:
    """Teacher-forcing labels (B,L) int64: [BOS, tokens..., EOS, PAD...] (BatchCollator contract)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.full((batch, length), d.pad, dtype=np.int64)
    hi = max(min_len + 1, length - 1)
    for b in range(batch):
        n = int(rng.integers(min(min_len, length - 2), hi))
        n = max(0, min(n, length - 2))
        out[b, 0] = d.bos
        out[b, 1:1 + n] = rng.integers(0, min(d.eos, d.vocab), size=n)
        out[b, 1 + n] = d.eos
    return torch.from_numpy(out)


def encoder_tokens(height: int, width: int, patch: int = 16) -> int:
    return (height // patch) * (width // patch) + 1


def encoder_flops(height: int, width: int, kind: str = "hybrid") -> float:
    """Algorithmic FLOPs of one encoder pass over one image."""
    n = encoder_tokens(height, width)
    first = 120864.0 * height * width if kind == "hybrid" else 131072.0 * (n - 1)
    return first + 11534336.0 * n + 8192.0 * n * n


def decode_step_bytes(batch: int, t: int, s: int, w_step: int = 15222736, kv_row: int = 8192, mem_row: int = 8192) -> float:
    """Algorithmic bytes of decode step t (1-based) with a bf16 KV cache.  kv_row = self-attention bytes
    per cached key (4 layers x K,V x 512 x bf16); mem_row = cross-attention bytes per memory token: 8192 with projected K/V,
    2048 when the K / V projections are absorbed and the [S, 256] bf16 memory itself is streamed (4 layers x 256 x bf16)."""
    return float(w_step + batch * (kv_row * t + mem_row * s + kv_row))
