"""ctypes binding of libtexocr_b200.so (include/texocr.h) and the thin ``Engine`` wrapper.

PyTorch is used for device memory and streams only: every tensor crosses the C-ABI as a raw
pointer (``data_ptr()``), host or device.  There is no CPU implementation behind this module --
if the shared library is missing or no sm_100 GPU is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch

from .spec import ModelDims

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TEXOCR_B200_LIB") or os.path.join(_HERE, "libtexocr_b200.so")      # the override is for A/B builds of experiments
ABI_VERSION = 1

# every symbol include/texocr.h declares (tests check the library exports exactly these)
EXPORTS = (
    "texocr_create", "texocr_destroy", "texocr_last_error", "texocr_set_weight", "texocr_finalize_weights",
    "texocr_encode", "texocr_decoder_logits", "texocr_decoder_generate", "texocr_generate", "texocr_cross_entropy",
    "texocr_kernel_launches", "texocr_profile_enable", "texocr_profile_read", "texocr_set_option", "texocr_debug_read",
    "texocr_debug_gemm", "texocr_debug_attn_decode", "texocr_debug_attn_abs", "texocr_debug_fold_absorbed", "texocr_set_sampling", "texocr_debug_sample_step", "texocr_preprocess_u8",
)


class TexocrConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "abi_version", "vocab_size", "max_length", "enc_layers", "dec_layers", "bos_token", "eos_token", "pad_token",
        "encoder_kind", "precision")]


class ProfileRow(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("launches", C.c_int64), ("ms", C.c_double), ("bytes", C.c_double),
                ("flops", C.c_double)]


_lib = None


def load_library() -> C.CDLL:
    """Load the in-tree shared library; raise (never fall back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m texocr_b200.build` (nvcc, sm_100a). "
            "texocr_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.texocr_create.argtypes = [C.POINTER(TexocrConfig), C.c_int, C.POINTER(vp)]
    lib.texocr_destroy.argtypes = [vp]
    lib.texocr_destroy.restype = None
    lib.texocr_last_error.argtypes = [vp]
    lib.texocr_last_error.restype = C.c_char_p
    lib.texocr_set_weight.argtypes = [vp, C.c_char_p, vp, i32, C.POINTER(i64)]
    lib.texocr_finalize_weights.argtypes = [vp]
    lib.texocr_encode.argtypes = [vp, vp, C.POINTER(i32), i32, vp, vp]
    lib.texocr_decoder_logits.argtypes = [vp, vp, vp, vp, C.POINTER(i32), i32, i32, vp, vp]
    lib.texocr_decoder_generate.argtypes = [vp, vp, i32, vp, C.POINTER(i32), i32, i32, vp, C.POINTER(i32), vp]
    lib.texocr_generate.argtypes = [vp, vp, C.POINTER(i32), i32, i32, vp, C.POINTER(i32), vp]
    lib.texocr_cross_entropy.argtypes = [vp, vp, vp, i64, vp, vp]
    lib.texocr_kernel_launches.argtypes = [vp]
    lib.texocr_kernel_launches.restype = i64
    lib.texocr_profile_enable.argtypes = [vp, i32]
    lib.texocr_profile_read.argtypes = [vp, C.POINTER(ProfileRow), i32]
    lib.texocr_set_option.argtypes = [vp, C.c_char_p, i64]
    lib.texocr_set_sampling.argtypes = [vp, C.c_double, C.c_double, C.c_uint64]
    lib.texocr_preprocess_u8.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp]
    lib.texocr_debug_sample_step.argtypes = [vp, vp, i32, i32, C.c_uint32, vp]
    lib.texocr_debug_read.argtypes = [vp, C.c_char_p, vp, i64]
    lib.texocr_debug_read.restype = i64
    lib.texocr_debug_gemm.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, i32, i32, vp, vp, vp]
    lib.texocr_debug_attn_decode.argtypes = [vp, i32, vp, i32, vp, vp, i32, vp, i64, i32, i32, i32, vp, vp, vp, i32, i32, i32, vp]
    lib.texocr_debug_fold_absorbed.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.texocr_debug_attn_abs.argtypes = [vp, vp, vp, i64, vp, vp, i32, vp, vp, i32, vp]
    for name in EXPORTS:
        getattr(lib, name)
    _lib = lib
    return lib


def _i32_array(values: Sequence[int]):
    return (C.c_int32 * len(values))(*[int(v) for v in values])


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


ImagesArg = Union[torch.Tensor, Sequence[torch.Tensor]]


class Engine:
    """One native handle: packed weights, workspaces, KV cache and the decode-step CUDA graph."""

    def __init__(self, dims: ModelDims, state_dict: Dict[str, torch.Tensor], precision: str = "fp32", device: int = 0):
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        if not torch.cuda.is_available():
            raise RuntimeError("texocr_b200 needs a CUDA (sm_100a) device: there is no CPU fallback")
        self.lib = load_library()
        self.dims = dims
        self.precision = precision
        self.device = torch.device("cuda", device)
        cfg = TexocrConfig(ABI_VERSION, dims.vocab, dims.max_length, dims.enc_layers, dims.dec_layers, dims.bos, dims.eos,
                           dims.pad, 0 if dims.encoder_kind == "hybrid" else 1, 1 if precision == "bf16" else 0)
        h = C.c_void_p()
        rc = self.lib.texocr_create(C.byref(cfg), device, C.byref(h))
        if rc != 0:
            raise RuntimeError(f"texocr_create failed ({rc}): {self.lib.texocr_last_error(None).decode()}")
        self.h = h
        for name, t in state_dict.items():
            if ".block." in name:       # nn.Sequential alias of block_list.* (model/resnet.py:130-139)
                continue
            tt = _f32c(t).cpu()
            shape = (C.c_int64 * tt.dim())(*tt.shape)
            self._check(self.lib.texocr_set_weight(self.h, name.encode(), tt.data_ptr(), tt.dim(), shape))
        self._check(self.lib.texocr_finalize_weights(self.h))

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int):
        if rc < 0:
            raise RuntimeError(f"texocr error {rc}: {self.lib.texocr_last_error(self.h).decode()}")
        return rc

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def close(self):
        if getattr(self, "h", None):
            self.lib.texocr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _pack_images(images: ImagesArg) -> Tuple[torch.Tensor, List[int]]:
        """(B,1,H,W) tensor or a list of (1,H,W)/(H,W) tensors -> flat float32 buffer + [H0,W0,H1,W1,...]."""
        if isinstance(images, torch.Tensor):
            if images.dim() != 4 or images.shape[1] != 1:
                raise ValueError("src must be (B,1,H,W) single-channel images")
            B, _, H, W = images.shape
            return _f32c(images).reshape(-1), [v for _ in range(B) for v in (H, W)]
        hw: List[int] = []
        flat = []
        for im in images:
            im = im.reshape(im.shape[-2], im.shape[-1])
            hw += [im.shape[0], im.shape[1]]
            flat.append(_f32c(im).reshape(-1))
        return torch.cat(flat), hw

    @staticmethod
    def token_counts(hw: Sequence[int]) -> List[int]:
        return [(hw[2 * i] // 16) * (hw[2 * i + 1] // 16) + 1 for i in range(len(hw) // 2)]

    # ------------------------------------------------------------------ entry points
    def encode_packed(self, images: ImagesArg, out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, List[int]]:
        flat, hw = self._pack_images(images)
        counts = self.token_counts(hw)
        if out is None:
            out = torch.empty((sum(counts), 256), dtype=torch.float32, device=self.device)
        self._check(self.lib.texocr_encode(self.h, flat.data_ptr(), _i32_array(hw), len(counts), out.data_ptr(), self._stream()))
        return out, counts

    def decoder_logits(self, ids: torch.Tensor, mask: Optional[torch.Tensor], enc: torch.Tensor, enc_len: Sequence[int],
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
        B, T = ids.shape
        ids = ids.detach().to(torch.int64).contiguous()
        m = None if mask is None else mask.detach().to(torch.uint8).contiguous()
        enc = _f32c(enc).reshape(-1, 256)
        if out is None:
            out = torch.empty((B, T, self.dims.vocab), dtype=torch.float32, device=self.device)
        self._check(self.lib.texocr_decoder_logits(self.h, ids.data_ptr(), None if m is None else m.data_ptr(), enc.data_ptr(),
                                                   _i32_array(enc_len), B, T, out.data_ptr(), self._stream()))
        return out

    def decoder_generate(self, start_tokens: torch.Tensor, eos_tok: Optional[int], enc: torch.Tensor, enc_len: Sequence[int],
                         max_len: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        B = start_tokens.shape[0]
        start = start_tokens.detach().to(torch.int64).reshape(B).contiguous()
        enc = _f32c(enc).reshape(-1, 256)
        if out is None:
            out = torch.empty((B, max_len), dtype=torch.int64, device=self.device)
        n = C.c_int32(0)
        self._check(self.lib.texocr_decoder_generate(self.h, start.data_ptr(), -1 if eos_tok is None else int(eos_tok),
                                                     enc.data_ptr(), _i32_array(enc_len), B, int(max_len), out.data_ptr(),
                                                     C.byref(n), self._stream()))
        return out[:, :n.value]

    def generate(self, images: ImagesArg, max_len: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Images (host or device) -> greedy token ids (B, n_steps) int64; `out` may be a pinned host tensor."""
        flat, hw = self._pack_images(images)
        B = len(hw) // 2
        if out is None:
            out = torch.empty((B, max_len), dtype=torch.int64, device=self.device)
        n = C.c_int32(0)
        self._check(self.lib.texocr_generate(self.h, flat.data_ptr(), _i32_array(hw), B, int(max_len), out.data_ptr(), C.byref(n),
                                             self._stream()))
        return out[:, :n.value]

    def cross_entropy(self, logits: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        logits = _f32c(logits).reshape(-1, self.dims.vocab)
        targets = targets.detach().to(torch.int64).reshape(-1).contiguous()
        out = torch.empty((), dtype=torch.float32, device=self.device)
        self._check(self.lib.texocr_cross_entropy(self.h, logits.data_ptr(), targets.data_ptr(), logits.shape[0], out.data_ptr(),
                                                  self._stream()))
        return out

    # ------------------------------------------------------------------ instrumentation
    def kernel_launches(self) -> int:
        return int(self.lib.texocr_kernel_launches(self.h))

    def profile_enable(self, on: bool):
        self._check(self.lib.texocr_profile_enable(self.h, 1 if on else 0))

    def profile_read(self) -> List[dict]:
        rows = (ProfileRow * 32)()
        n = self._check(self.lib.texocr_profile_read(self.h, rows, 32))
        return [dict(name=rows[i].name.decode(), launches=rows[i].launches, ms=rows[i].ms, bytes=rows[i].bytes,
                     flops=rows[i].flops) for i in range(n)]

    def set_option(self, name: str, value: int):
        self._check(self.lib.texocr_set_option(self.h, name.encode(), int(value)))

    def preprocess_u8(self, images: Sequence[torch.Tensor], pad_multiple: int = 16) -> List[torch.Tensor]:
        """uint8 images (H, W) / (H, W, 1) / (H, W, 3), host or device -> list of float32 (1, Hp, Wp) device tensors
        (ToTensor -> Grayscale -> Invert of data_wrangling/dataset.py:365-371, zero-padded to multiples of pad_multiple)."""
        flat, hwc = [], []
        for im in images:
            im = torch.as_tensor(im)
            if im.dtype != torch.uint8 or im.ndim not in (2, 3):
                raise ValueError("preprocess_u8 takes uint8 images of shape (H, W) or (H, W, C)")
            c = 1 if im.ndim == 2 else im.shape[2]
            hwc += [im.shape[0], im.shape[1], c]
            flat.append(im.reshape(-1))
        dev = all(t.is_cuda for t in flat)
        buf = torch.cat([t if dev else t.cpu() for t in flat]).contiguous()
        B = len(images)
        pm = int(pad_multiple)
        sizes = [((hwc[3 * b] + pm - 1) // pm * pm, (hwc[3 * b + 1] + pm - 1) // pm * pm) for b in range(B)]
        out = torch.empty((sum(h * w for h, w in sizes),), dtype=torch.float32, device=self.device)
        out_hw = (C.c_int32 * (2 * B))()
        self._check(self.lib.texocr_preprocess_u8(self.h, buf.data_ptr(), _i32_array(hwc), B, pm, out.data_ptr(), out_hw, self._stream()))
        res, off = [], 0
        for b in range(B):
            h, w = out_hw[2 * b], out_hw[2 * b + 1]
            res.append(out[off:off + h * w].view(1, h, w))
            off += h * w
        return res

    def set_sampling(self, temp: float, threshold: float = 0.9, seed: int = 0):
        """temp > 0: later generate calls sample like model/decoder.py:103-108 (top-k filter, softmax(/temp), one draw);
        temp <= 0: greedy.  Draws are reproducible for a given seed."""
        self._check(self.lib.texocr_set_sampling(self.h, float(temp), float(threshold), int(seed) & (2 ** 64 - 1)))

    def debug_sample_step(self, logits: torch.Tensor, step: int = 0, call: int = 0) -> torch.Tensor:
        """Test hook: the token-selection kernel on float32 device logits (rows, vocab) -> int64 (rows,)."""
        logits = _f32c(logits)
        out = torch.empty((logits.shape[0],), dtype=torch.int64, device=self.device)
        self._check(self.lib.texocr_debug_sample_step(self.h, logits.data_ptr(), logits.shape[0], int(step), int(call), out.data_ptr()))
        return out

    def debug_gemm(self, A, W, C, epi=0, bias=None, res=None, use_tc=True, A2=None, W2=None, ldc=None):
        """Test hook: C = epi(A @ W.T) through the engine's GEMM kernels (tensors on the device, row-major).
        ``ldc`` overrides C's leading dimension (epi 4: in {max, index} pairs)."""
        M, K = A.shape
        N = W.shape[0]
        dt_a = 1 if A.dtype == torch.bfloat16 else 0
        dt_c = 1 if C.dtype == torch.bfloat16 else 0
        ptr = lambda t: None if t is None else t.data_ptr()
        self._check(self.lib.texocr_debug_gemm(self.h, A.data_ptr(), W.data_ptr(), C.data_ptr(), M, N, K, A.stride(0), W.stride(0),
                                               C.stride(0) if ldc is None else int(ldc), epi, dt_a, dt_c, ptr(bias), ptr(res), 0 if res is None else res.stride(0),
                                               1 if use_tc else 0, ptr(A2), ptr(W2), self._stream()))
        return C

    def debug_attn_decode(self, self_attn, q, knew, vnew, kv, col0, tcap, k_off, step, batch, max_keys, use_tma):
        """Test hook: one decode-attention launch; returns bf16 [batch, 512]."""
        out = torch.zeros((batch, 512), dtype=torch.bfloat16, device=self.device)
        ptr = lambda t: None if t is None else t.data_ptr()
        self._check(self.lib.texocr_debug_attn_decode(
            self.h, 1 if self_attn else 0, q.data_ptr(), q.stride(0), ptr(knew), ptr(vnew), 0 if knew is None else knew.stride(0),
            kv.data_ptr(), kv.shape[0], kv.stride(0), col0, tcap, ptr(k_off), ptr(step), out.data_ptr(), batch, max_keys,
            1 if use_tma else 0, self._stream()))
        return out

    def debug_attn_abs(self, q: torch.Tensor, latent: torch.Tensor, k_off: Optional[torch.Tensor] = None,
                       znew: Optional[torch.Tensor] = None, tcap: int = 0, step: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Test hook: absorbed decode attention; q bf16 [B, 2048], latent bf16 [rows, 256]; cross: k_off int32 [B+1] (device);
        self: znew bf16 [B, 256], tcap, step int32 [1] (device) -> bf16 [B, 2048]."""
        out = torch.zeros((q.shape[0], 2048), dtype=torch.bfloat16, device=self.device)
        ptr = lambda t: None if t is None else t.data_ptr()
        self._check(self.lib.texocr_debug_attn_abs(self.h, q.data_ptr(), latent.data_ptr(), latent.shape[0], ptr(k_off), ptr(znew),
                                                   int(tcap), ptr(step), out.data_ptr(), q.shape[0], self._stream()))
        return out

    def debug_read(self, name: str, numel: int) -> torch.Tensor:
        out = torch.empty(numel, dtype=torch.float32, device=self.device)
        n = self.lib.texocr_debug_read(self.h, name.encode(), out.data_ptr(), numel)
        if n < 0:
            self._check(int(n))
        return out[:n]
