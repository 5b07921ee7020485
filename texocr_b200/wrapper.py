"""``TeXOCRWrapper`` (model/ocr_model.py:69-110): image in, LaTeX string out, on the B200 path.

    wrapper = TeXOCRWrapper(config)          # config['tokenizer_path'], config['model_path'], config['device'], model keys
    tokens, latex = wrapper(img)             # img: PIL image or uint8 array (H, W[, 3]); max_len=350, temp=0.3

Differences from the reference, all stated in DESIGN.md: the image transform runs on the GPU without the reference's
train-time RandomAffine; images are padded to the encoder's 16-pixel grid; decoding is greedy unless ``sample=True``
(then the reference's top-k / temperature draw from a seeded Philox stream); ``batch()`` decodes many images at once.
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import checkpoint
from .detok import Detokenizer
from .model import create_model


def _as_u8(img) -> torch.Tensor:
    if hasattr(img, "mode") and hasattr(img, "size") and not isinstance(img, (np.ndarray, torch.Tensor)):      # PIL image
        if img.mode not in ("L", "RGB"):
            img = img.convert("RGB")
        img = np.asarray(img)
    t = torch.as_tensor(img)
    if t.dtype != torch.uint8:
        raise ValueError("images must be uint8 (PIL image, numpy array or tensor)")
    return t


class TeXOCRWrapper(object):
    def __init__(self, config: dict, precision: Optional[str] = None):
        self.tokenizer = Detokenizer.load(config["tokenizer_path"])
        config = dict(config)
        config["vocab_size"] = self.tokenizer.vocab_size                                   # model/ocr_model.py:76
        self.model = create_model(config, precision=precision)
        obj = torch.load(config["model_path"], map_location="cpu", weights_only=True)
        sd, _ = checkpoint.extract_state_dict(obj)
        checkpoint.load_state_dict_resizing(self.model, sd)                               # incl. the positional-table resize
        self.eos = self.model.eos_token

    def batch(self, imgs: Sequence, max_len: int = 350, temp: float = 0.3, sample: bool = False, seed: int = 0) -> Tuple[torch.Tensor, List[str]]:
        """Many images (any sizes) -> (token ids (B, n_steps) on the device, LaTeX strings cut at each row's EOS)."""
        src = self.model.engine().preprocess_u8([_as_u8(i) for i in imgs], 16)
        max_len = min(int(max_len), self.model.dims.max_length)
        tokens = self.model.generate(src, max_len=max_len, temp=temp, sample=sample, seed=seed)
        return tokens, self.tokenizer.decode_batch(tokens.cpu(), eos_token=self.eos)

    def __call__(self, img, max_len: int = 350, temp: float = 0.3, sample: bool = False, seed: int = 0) -> Tuple[List[int], str]:
        """One image -> (token ids without the EOS, LaTeX string), like the reference's wrapper."""
        tokens, text = self.batch([img], max_len, temp, sample, seed)
        row = tokens[0].tolist()
        out_tokens = row[: row.index(self.eos)] if self.eos in row else row
        return out_tokens, text[0]
