"""Host-side mirror of the reference's model surface (model/ocr_model.py) over the native engine.

Same names, argument meaning and error behaviour as the reference for the inference path:

    model = create_model(config)                      # model/ocr_model.py:113-130
    tokens = model.generate(src, max_len=256)         # model/ocr_model.py:46-66   (greedy; see below)
    loss = model(src, trg)                            # model/ocr_model.py:38-44
    enc = model.encoder(src)                          # model/encoder.py:128-152
    logits = model.decoder.net(ids, mask=m, enc=enc)  # model/decoder.py:41-67
    loss, logits = model.decoder(trg, enc=enc, mask=m, return_out=True)   # model/decoder.py:124-145
    model.decoder.generate(start_tokens=, eos_tok=, max_len=, temp=, enc=)  # model/decoder.py:77-122

The module tree only HOLDS parameters under the reference's state_dict names, so
``state_dict()/load_state_dict()/to()/eval()`` behave as a caller of the reference expects; all
arithmetic happens in libtexocr_b200.so.  Deliberate deviations: ``generate`` is greedy by
default (argmax; ``sample=True`` selects the reference's top-k / temperature draw with a reproducible Philox stream),
inputs must be multiples of 16 with H<=160, W<=1008, ``max_len <= config['max_length']``, and there is no CPU
execution path.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Sequence, Union

import torch
import torch.nn as nn

from . import spec
from ._lib import Engine


class _Holder(nn.Module):
    """A node of the parameter tree; children and parameters are attached by name."""


def _attach(root: nn.Module, key: str, param: nn.Parameter):
    parts = key.split(".")
    node = root
    for p in parts[:-1]:
        nxt = node._modules.get(p)
        if nxt is None:
            nxt = _Holder()
            node.add_module(p, nxt)
        node = nxt
    node.register_parameter(parts[-1], param)


def _init_tensor(p: spec.ParamSpec, gen: torch.Generator) -> torch.Tensor:
    if p.init in ("conv", "linear_w", "linear_b"):
        bound = 1.0 / math.sqrt(p.fan_in)
        return (torch.rand(p.shape, generator=gen) * 2 - 1) * bound
    if p.init == "normal02":
        return torch.randn(p.shape, generator=gen) * 0.02
    if p.init == "ones":
        return torch.ones(p.shape)
    return torch.zeros(p.shape)


class _Encoder(_Holder):
    def forward(self, src):
        return self._owner().encode(src)


class _Net(_Holder):
    """model.decoder.net -- Transformer.forward(ids, mask=, enc=) (model/decoder.py:41-67)."""
    max_len: int = 0

    def forward(self, x, mask=None, return_embeddings=False, return_attn=False, **kwargs):
        if return_embeddings or return_attn:
            raise NotImplementedError("return_embeddings / return_attn are not on the inference path")
        enc = kwargs.get("enc")
        if enc is None:
            raise AssertionError("Must provide enc if cross_attend is True.")       # model/attention.py:232-233
        if kwargs.get("enc_mask") is not None:
            raise NotImplementedError("enc_mask is never passed by the reference and is not implemented")
        return self._owner().decoder_logits(x, mask, enc)


class _Decoder(_Holder):
    """model.decoder -- AutoRegressiveDecoder (model/decoder.py:70-145)."""

    @property
    def max_len(self):
        return self.net.max_len

    @torch.no_grad()
    def generate(self, start_tokens, eos_tok, max_len, temp=1.0, **kwargs):
        enc = kwargs.get("enc")
        if enc is None:
            raise AssertionError("Must provide enc if cross_attend is True.")
        if "mask" in kwargs and kwargs["mask"] is not None and not bool(kwargs["mask"].all()):
            raise NotImplementedError("generate with a partial start mask is not implemented")
        squeeze = start_tokens.ndim == 1
        start_tokens = start_tokens[None, :] if squeeze else start_tokens           # model/decoder.py:88
        if start_tokens.shape[1] != 1:
            raise NotImplementedError("texocr_b200 generates from a single start column (the BOS column of "
                                      "model/ocr_model.py:57); longer prompts are not implemented")
        sample = bool(kwargs.get("sample", False))
        eng = self._owner().engine()
        if sample:
            eng.set_sampling(temp, 0.9, int(kwargs.get("seed", 0)))
        try:
            out = self._owner().decoder_generate(start_tokens, eos_tok, enc, max_len)
        finally:
            if sample:
                eng.set_sampling(0.0)
        return out.squeeze(0) if squeeze else out

    def forward(self, x, mask=None, return_out=False, **kwargs):
        x_in, x_out = x[:, :-1], x[:, 1:]                                           # model/decoder.py:133-134
        if mask is not None and mask.shape[1] == x.shape[1]:
            mask = mask[:, :-1]
        out = self.net(x_in, mask=mask, **kwargs)
        loss = self._owner().engine().cross_entropy(out, x_out.to(out.device))
        return (loss, out) if return_out else loss


class OCRModel(nn.Module):
    """Drop-in for the reference OCRModel (model/ocr_model.py:14-66) on one B200."""

    def __init__(self, config: dict, encoder_kind: str = "hybrid", precision: Optional[str] = None, seed: Optional[int] = None):
        super().__init__()
        self.dims = spec.dims_from_config(config, encoder_kind)
        self.bos_token = self.dims.bos
        self.eos_token = self.dims.eos
        self.trg_pad_idx = self.dims.pad
        self.device = torch.device(config.get("device", "cuda"))
        self.precision = precision or config.get("precision") or os.environ.get("TEXOCR_PRECISION", "fp32")
        gen = torch.Generator().manual_seed(torch.initial_seed() if seed is None else seed)
        self.encoder = _Encoder()
        self.decoder = _Decoder()
        self.decoder.add_module("net", _Net())
        self.decoder.net.max_len = self.dims.max_length
        made: Dict[str, nn.Parameter] = {}
        for p in spec.param_table(self.dims):
            if p.alias_of is not None:
                param = made[p.alias_of]
            else:
                param = nn.Parameter(_init_tensor(p, gen), requires_grad=False)
                made[p.key] = param
            _attach(self, p.key, param)
        for m in (self.encoder, self.decoder, self.decoder.net):
            object.__setattr__(m, "_owner", self._self_ref)
        self._engine: Optional[Engine] = None
        self._engine_key = None
        self._pipe = None
        self._pipe_key = None

    def _self_ref(self):
        return self

    # ------------------------------------------------------------------ engine lifecycle
    def _weights_key(self):
        return (self.precision, str(self.device), self.dims,
                tuple((p.data_ptr(), p._version) for p in self.parameters()))

    def engine(self) -> Engine:
        """The native handle for the current weights; rebuilt when parameters were replaced or modified."""
        if self.device.type != "cuda":
            raise RuntimeError("texocr_b200 runs on a CUDA (sm_100a) device only; config['device'] / .to() selected "
                               f"'{self.device}'. There is no CPU fallback.")
        key = self._weights_key()
        if self._engine is None or key != self._engine_key:
            if self._engine is not None:
                self._engine.close()
            idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
            self._engine = Engine(self.dims, self.state_dict(), self.precision, idx)
            self._engine_key = key
        return self._engine

    # ------------------------------------------------------------------ one large call = several sub-batches in flight
    sub_batch = 512          # rows per sub-batch of a large generate call (BASELINE configs[2]'s batch)
    max_in_flight = 6

    def _generate_sub_batches(self, src: torch.Tensor, max_len: int) -> torch.Tensor:
        """``generate`` for B >= 2 x sub_batch: the rows are independent, so the call is cut into sub-batches of ``sub_batch``
        rows that are decoded concurrently by internal engine replicas (texocr_b200/pipeline.py) -- a caller of the
        reference's loop (test.py:27-40) with a larger batch gets the pipelined rate from ONE call.  Same tokens as
        sub-batch-sized calls.  The reference's shape contract is kept: the result is as wide as the slowest sub-batch needs
        (the step at which every row of the WHOLE batch holds an EOS, model/decoder.py:115-118); a sub-batch that finished
        earlier is decoded again without its early exit so that its rows carry the tokens the reference would have there."""
        from .pipeline import GeneratePipeline
        key = self._weights_key()
        chunks = list(torch.split(src, self.sub_batch))
        want = min(self.max_in_flight, len(chunks))
        if self._pipe is None or self._pipe_key != key or len(self._pipe.engines) < want:
            if self._pipe is not None:
                self._pipe.close()
            self._pipe = GeneratePipeline(self, in_flight=want, branches=1)
            self._pipe_key = key
        outs = list(self._pipe.generate_batches(chunks, max_len))
        n = max(o.shape[1] for o in outs)
        short = [i for i, o in enumerate(outs) if o.shape[1] < n]
        if short:
            for e in self._pipe.engines:
                e.set_option("no_early_exit", 1)
            try:
                for i, o in zip(short, self._pipe.generate_batches([chunks[i] for i in short], n)):
                    outs[i] = o
            finally:
                for e in self._pipe.engines:
                    e.set_option("no_early_exit", 0)
        return torch.cat(outs, dim=0)

    def resize_pos_embedding(self, new_len: int):
        """TeXOCRWrapper.__init__ (model/ocr_model.py:82-90): a checkpoint whose decoder positional table has another
        length than config['max_length'] replaces the table; the engine is rebuilt for that length on next use."""
        import dataclasses
        if new_len == self.dims.max_length:
            return self
        key = "decoder.net.pos_embedding.embedding.weight"
        old = self.state_dict()[key]
        _attach(self, key, nn.Parameter(torch.zeros((new_len, old.shape[1]), dtype=old.dtype, device=old.device), requires_grad=False))
        self.dims = dataclasses.replace(self.dims, max_length=int(new_len))
        self.decoder.net.max_len = int(new_len)
        return self

    def set_precision(self, precision: str):
        self.precision = precision
        return self

    def to(self, *args, **kwargs):
        out = super().to(*args, **kwargs)
        try:
            self.device = next(self.parameters()).device
        except StopIteration:
            pass
        return out

    # ------------------------------------------------------------------ reference surface
    def make_trg_mask(self, trg: torch.Tensor) -> torch.Tensor:
        return (trg != self.trg_pad_idx).to(self.device)                           # model/ocr_model.py:34-36

    def encode(self, src):
        eng = self.engine()
        packed, counts = eng.encode_packed(src)
        if isinstance(src, torch.Tensor):
            return packed.reshape(src.shape[0], counts[0], 256)
        return list(torch.split(packed, counts))                                   # ragged batch: one (N_i,256) per image

    @staticmethod
    def _pack_memory(enc):
        if isinstance(enc, torch.Tensor):
            if enc.dim() != 3:
                raise ValueError("enc must be (B,S,256)")
            return enc.reshape(-1, enc.shape[-1]), [enc.shape[1]] * enc.shape[0]
        return torch.cat(list(enc)), [e.shape[0] for e in enc]

    def decoder_logits(self, ids, mask, enc):
        packed, lens = self._pack_memory(enc)
        return self.engine().decoder_logits(ids, mask, packed, lens)

    def decoder_generate(self, start_tokens, eos_tok, enc, max_len):
        packed, lens = self._pack_memory(enc)
        return self.engine().decoder_generate(start_tokens, eos_tok, packed, lens, max_len)

    def forward(self, src, trg, return_out: bool = False):
        trg_mask = self.make_trg_mask(trg)
        enc = self.encoder(src)
        return self.decoder(trg, enc=enc, mask=trg_mask)       # return_out is accepted and ignored, as in the reference

    @torch.no_grad()
    def generate(self, src, max_len: int, temp: float = 0.3, sample: bool = False, seed: int = 0):
        """(B, n_steps) int64 on the model's device.  Greedy by default (the temp -> 0 limit of model/decoder.py:103-108,
        the path BASELINE.json measures); ``sample=True`` draws like the reference (top-k 0.9 filter, softmax(/temp),
        one multinomial draw per step) from a Philox stream seeded with ``seed``.  ``src`` may also be a list of
        (1,H,W) images of different sizes (ragged batch) -- an extension over the reference's same-size batches."""
        eng = self.engine()
        if not sample:
            if isinstance(src, torch.Tensor) and src.dim() == 4 and src.shape[0] >= 2 * self.sub_batch:
                return self._generate_sub_batches(src, max_len)
            return eng.generate(src, max_len)
        eng.set_sampling(temp, 0.9, seed)
        try:
            return eng.generate(src, max_len)
        finally:
            eng.set_sampling(0.0)


def create_model(config: dict, encoder_kind: str = "hybrid", precision: Optional[str] = None) -> OCRModel:
    """create_model(config) of model/ocr_model.py:113-130.  The weights start at the reference's default
    initialisation; ``load_state_dict`` takes a reference checkpoint's ``model_state_dict`` unchanged."""
    model = OCRModel(config, encoder_kind=encoder_kind, precision=precision)
    dev = torch.device(config.get("device", "cuda"))
    if dev.type == "cuda" and torch.cuda.is_available():
        model.to(dev)
    return model
