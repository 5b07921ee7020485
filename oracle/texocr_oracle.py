"""CPU oracle for the TeXOCR inference hot path -- TEST INFRASTRUCTURE ONLY.

This file restates, function by function, what the reference (olibridge01/TeXOCR,
pure Python on PyTorch ATen) computes on the path SURVEY.md section 8 scopes:
ResNetV2-hybrid / patch ViT encoder -> transformer decoder -> greedy generate loop.
It is the checker for the CUDA path, never the thing measured or shipped: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product package ``texocr_b200`` never does.

Arithmetic: float32 on torch CPU ops (``F.conv2d``, ``F.linear``, ``torch.einsum``,
``F.group_norm``, ``F.layer_norm``, exact-erf ``F.gelu``) -- the third-party arithmetic
the reference itself bottoms out in (ATen / oneDNN / MKL; the reference pins no version,
``requirements.txt:1-2``; this image has torch 2.11.0).  State is a plain ``state_dict``
with the reference's key names.

Parity pin: the reference holds no tests, golden vectors or checkpoints,
so the oracle is pinned against outputs of the *unmodified reference modules* executed in
the build container: ``tests/golden/make_golden.py`` imports ``/root/reference`` as package
``TeXOCR``, loads the same seeded weights and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against them.

Greedy decoding: the reference samples (top-k -> softmax(/temp) -> multinomial,
``model/decoder.py:104-108``).  "Greedy" (BASELINE.json) is the temp -> 0 limit, i.e.
argmax of the last-position logits: top-k keeps the maximum and softmax is monotone.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

GRID_W = 63           # 1008 // 16  (model/encoder.py:137-141, create_encoder img_size=(160,1008))
GRID_H = 10
HEADS = 8
DIM_HEAD = 64         # model/attention.py:76


# --------------------------------------------------------------------------- helpers
def same_pad_amount(x: int, k: int, s: int, d: int = 1) -> int:
    """utils.py:98-100 get_same_padding."""
    return max((math.ceil(x / s) - 1) * s + (k - 1) * d + 1 - x, 0)


def pad_same(x: torch.Tensor, k: int, s: int, value: float = 0.0) -> torch.Tensor:
    """utils.py:116-123: TF-'SAME' dynamic padding, split p//2 before / p - p//2 after."""
    ih, iw = x.shape[-2:]
    ph, pw = same_pad_amount(ih, k, s), same_pad_amount(iw, k, s)
    if ph > 0 or pw > 0:
        x = F.pad(x, [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2], value=value)
    return x


def standardise_weight(w: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """model/resnet.py:61-64: per-out-channel (w - mean) / sqrt(biased var + 1e-6)."""
    flat = w.reshape(w.shape[0], -1)
    mean = flat.mean(dim=1, keepdim=True)
    var = flat.var(dim=1, unbiased=False, keepdim=True)
    return ((flat - mean) / torch.sqrt(var + eps)).reshape_as(w)


def std_conv2d(x: torch.Tensor, w: torch.Tensor, k: int, stride: int) -> torch.Tensor:
    """model/resnet.py:38-66 StdConv2d.forward; padding rule utils.py:93-114 (static iff stride 1)."""
    if stride == 1:
        pad = (k - 1) // 2
    else:
        x = pad_same(x, k, stride)
        pad = 0
    return F.conv2d(x, standardise_weight(w), None, stride, pad)


def group_norm_act(x: torch.Tensor, sd: SD, prefix: str, act: bool) -> torch.Tensor:
    """model/resnet.py:14-35 GroupNormAct: 32 groups, eps 1e-5, affine, optional ReLU."""
    x = F.group_norm(x, 32, sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-5)
    return F.relu(x) if act else x


# --------------------------------------------------------------------------- backbone
def bottleneck(x: torch.Tensor, sd: SD, p: str, stride: int, has_ds: bool) -> torch.Tensor:
    """model/resnet.py:141-149 Bottleneck.forward (stride on the 3x3 and on the downsample conv)."""
    res = x
    if has_ds:
        res = std_conv2d(x, sd[p + ".downsample.conv.weight"], 1, stride)
        res = group_norm_act(res, sd, p + ".downsample.norm", act=False)
    y = std_conv2d(x, sd[p + ".block_list.0.weight"], 1, 1)
    y = group_norm_act(y, sd, p + ".block_list.1", act=True)
    y = std_conv2d(y, sd[p + ".block_list.2.weight"], 3, stride)
    y = group_norm_act(y, sd, p + ".block_list.3", act=True)
    y = std_conv2d(y, sd[p + ".block_list.4.weight"], 1, 1)
    y = group_norm_act(y, sd, p + ".block_list.5", act=False)
    return F.relu(y + res)


def resnetv2(img: torch.Tensor, sd: SD, bb: str = "encoder.patch_embed.backbone_net",
             taps: Optional[dict] = None) -> torch.Tensor:
    """model/resnet.py:200-254 ResNetV2([2,4,6]).forward: stem (7x7 s2, GN+ReLU, maxpool 3x3 s2 with
    -inf SAME pad, model/resnet.py:69-79) then three stages of bottlenecks (model/resnet.py:152-197)."""
    x = std_conv2d(img, sd[bb + ".stem.0.weight"], 7, 2)
    x = group_norm_act(x, sd, bb + ".stem.1", act=True)
    x = F.max_pool2d(pad_same(x, 3, 2, value=-float("inf")), 3, 2)
    if taps is not None:
        taps["stem"] = x
    for s, depth in enumerate((2, 4, 6)):
        for b in range(depth):
            stride = (1 if s == 0 else 2) if b == 0 else 1
            x = bottleneck(x, sd, f"{bb}.stages.{s}.stage_blocks.{b}", stride, has_ds=(b == 0))
            if taps is not None:
                taps[f"s{s}b{b}"] = x
    return x


# --------------------------------------------------------------------------- transformer blocks
def _count_sublayers(sd: SD, prefix: str) -> int:
    """Number of entries of AttentionLayers.layers (each holds [norm, block, residual], model/attention.py:221)."""
    n = 0
    while f"{prefix}.layers.{n}.0.weight" in sd:
        n += 1
    return n


def multi_head_attention(x: torch.Tensor, sd: SD, p: str, causal: bool, enc: Optional[torch.Tensor] = None,
                         mask: Optional[torch.Tensor] = None, enc_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """model/attention.py:101-180 MultiHeadAttention.forward (without the debug clones)."""
    B, N, _ = x.shape
    kv = enc if enc is not None else x
    q = F.linear(x, sd[p + ".q.weight"])
    k = F.linear(kv, sd[p + ".k.weight"])
    v = F.linear(kv, sd[p + ".v.weight"])
    split = lambda t: t.reshape(t.shape[0], t.shape[1], HEADS, DIM_HEAD).permute(0, 2, 1, 3)   # 'b n (h d) -> b h n d'
    q, k, v = split(q), split(k), split(v)
    energy = torch.einsum("bhid,bhjd->bhij", q, k) * (DIM_HEAD ** -0.5)      # scale AFTER the product (line 148)
    fill = -torch.finfo(energy.dtype).max                                      # utils.py:81-83
    if mask is not None or enc_mask is not None:                               # lines 130-145
        q_mask = mask if mask is not None else torch.ones(B, N, dtype=torch.bool)
        k_mask = q_mask if enc is None else enc_mask
        if k_mask is None:
            k_mask = torch.ones(B, k.shape[-2], dtype=torch.bool)
        allowed = q_mask[:, None, :, None] & k_mask[:, None, None, :]
        energy = energy.masked_fill(~allowed, fill)
    if causal:                                                                 # lines 158-164
        i, j = energy.shape[-2:]
        r_i = torch.arange(i)[:, None]
        r_j = torch.arange(j)[None, :]
        future = r_j > r_i + (j - i)
        energy = energy.masked_fill(future[None, None], fill)
    attn = F.softmax(energy, dim=-1)
    out = torch.einsum("bhij,bhjd->bhid", attn, v)
    out = out.permute(0, 2, 1, 3).reshape(B, N, HEADS * DIM_HEAD)
    y = F.linear(out, sd[p + ".fc_out.0.weight"], sd[p + ".fc_out.0.bias"])
    return F.glu(y, dim=-1)                                                    # nn.GLU (lines 96-99)


def mlp_geglu(x: torch.Tensor, sd: SD, p: str) -> torch.Tensor:
    """model/attention.py:9-17 GeGLU + 41-67 MLP: fc -> a * gelu_erf(gate) -> fc_out."""
    u = F.linear(x, sd[p + ".fc_in.fc.weight"], sd[p + ".fc_in.fc.bias"])
    a, g = u.chunk(2, dim=-1)
    return F.linear(a * F.gelu(g), sd[p + ".fc_out.weight"], sd[p + ".fc_out.bias"])


def attention_layers(x: torch.Tensor, sd: SD, p: str, kinds: Tuple[str, ...], causal: bool,
                     enc: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
                     enc_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """model/attention.py:223-269 AttentionLayers.forward: ONE LayerNorm shared by the stack, applied
    before every block and again after every block but the last (lines 242-259)."""
    g, b = sd[p + ".layers.0.0.weight"], sd[p + ".layers.0.0.bias"]
    dim = x.shape[-1]
    for i, kind in enumerate(kinds):
        residual = x
        x = F.layer_norm(x, (dim,), g, b, 1e-5)
        lp = f"{p}.layers.{i}.1"
        if kind == "self":
            out = multi_head_attention(x, sd, lp, causal, mask=mask)
        elif kind == "cross":
            out = multi_head_attention(x, sd, lp, False, enc=enc, mask=mask, enc_mask=enc_mask)
        else:
            out = mlp_geglu(x, sd, lp)
        x = out + residual
        if i != len(kinds) - 1:
            x = F.layer_norm(x, (dim,), g, b, 1e-5)
    return x


# --------------------------------------------------------------------------- encoder
def patch_tokens(img: torch.Tensor, sd: SD, kind: str, taps: Optional[dict] = None) -> torch.Tensor:
    """model/encoder.py:65-72 (hybrid: backbone then 1x1 projection conv) or 25-28 (16x16 s16 conv);
    both end with flatten(2).transpose(1,2): row-major tokens over the feature grid."""
    w, b = sd["encoder.patch_embed.proj.weight"], sd["encoder.patch_embed.proj.bias"]
    if kind == "hybrid":
        feat = resnetv2(img, sd, taps=taps)
        if taps is not None:
            taps["backbone"] = feat
        x = F.conv2d(feat, w, b, 1)
    else:
        x = F.conv2d(img, w, b, w.shape[-1])
    return x.flatten(2).transpose(1, 2)


def encoder_forward(sd: SD, img: torch.Tensor, kind: str = "hybrid", taps: Optional[dict] = None) -> torch.Tensor:
    """model/encoder.py:128-152 VisionTransformer.forward with num_classes=0 (head = Identity)."""
    B, _, H, W = img.shape
    x = patch_tokens(img, sd, kind, taps)
    x = torch.cat((sd["encoder.cls_token"].expand(B, -1, -1), x), dim=1)
    h, w = H // 16, W // 16
    grid = torch.arange(GRID_H * GRID_W).reshape(GRID_H, GRID_W)
    pos_ids = torch.cat((torch.zeros(1, dtype=torch.long), grid[:h, :w].reshape(-1) + 1))
    x = x + sd["encoder.pos_embed"][:, pos_ids]
    if taps is not None:
        taps["tokens_in"] = x
    n_layers = _count_sublayers(sd, "encoder.attn_layers") // 2
    x = attention_layers(x, sd, "encoder.attn_layers", ("self", "mlp") * n_layers, causal=False)
    return F.layer_norm(x, (x.shape[-1],), sd["encoder.norm.weight"], sd["encoder.norm.bias"], 1e-5)


# --------------------------------------------------------------------------- decoder
def _dec_kinds(sd: SD) -> Tuple[str, ...]:
    n = _count_sublayers(sd, "decoder.net.attn_layers")
    return ("self", "cross", "mlp") * (n // 3)


def decoder_logits(sd: SD, ids: torch.Tensor, enc: torch.Tensor, mask: Optional[torch.Tensor] = None,
                   return_embeddings: bool = False) -> torch.Tensor:
    """model/decoder.py:41-67 Transformer.forward in eval mode (embed dropout inactive)."""
    T = ids.shape[1]
    x = F.embedding(ids, sd["decoder.net.token_embedding.weight"])
    x = x + sd["decoder.net.pos_embedding.embedding.weight"][:T][None]          # model/attention.py:30-32
    x = attention_layers(x, sd, "decoder.net.attn_layers", _dec_kinds(sd), causal=True, enc=enc, mask=mask)
    x = F.layer_norm(x, (x.shape[-1],), sd["decoder.net.norm.weight"], sd["decoder.net.norm.bias"], 1e-5)
    if return_embeddings:
        return x
    return F.linear(x, sd["decoder.net.to_logits.weight"], sd["decoder.net.to_logits.bias"])


def decoder_loss(sd: SD, trg: torch.Tensor, enc: torch.Tensor, mask: Optional[torch.Tensor] = None):
    """model/decoder.py:124-145 AutoRegressiveDecoder.forward: shift by one, CE without ignore_index."""
    x_in, x_out = trg[:, :-1], trg[:, 1:]
    if mask is not None and mask.shape[1] == trg.shape[1]:
        mask = mask[:, :-1]
    out = decoder_logits(sd, x_in, enc, mask)
    return F.cross_entropy(out.transpose(1, 2), x_out), out


def model_forward(sd: SD, src: torch.Tensor, trg: torch.Tensor, pad: int = 999, kind: str = "hybrid"):
    """model/ocr_model.py:34-44 OCRModel.forward (+ make_trg_mask)."""
    return decoder_loss(sd, trg, encoder_forward(sd, src, kind), trg != pad)


def generate_greedy_recompute(sd: SD, enc: torch.Tensor, max_len: int, bos: int = 998, eos: Optional[int] = 997,
                              gaps: Optional[list] = None) -> torch.Tensor:
    """model/decoder.py:77-122 AutoRegressiveDecoder.generate, the reference algorithm as written:
    every step re-runs the decoder over the whole (windowed) prefix, cross-attention K/V included,
    takes the last-position logits and appends argmax (the greedy limit of lines 104-108).
    Stops after the first step at which every row holds an EOS (lines 115-116); rows are not frozen."""
    B = enc.shape[0]
    dec_max = sd["decoder.net.pos_embedding.embedding.weight"].shape[0]
    output = torch.full((B, 1), bos, dtype=torch.long)
    mask = torch.ones_like(output, dtype=torch.bool)
    for _ in range(max_len):
        x = output[:, -dec_max:]
        mask = mask[:, -dec_max:]
        logits = decoder_logits(sd, x, enc, mask)[:, -1, :]
        nxt = logits.argmax(dim=-1, keepdim=True)
        if gaps is not None:
            top2 = logits.topk(2, dim=-1).values
            gaps.append((top2[:, 0] - top2[:, 1]).clone())
        output = torch.cat((output, nxt), dim=-1)
        mask = F.pad(mask, (0, 1), value=True)
        if eos is not None and (output == eos).any(dim=1).all():
            break
    return output[:, 1:]


def generate_greedy_cached(sd: SD, enc: torch.Tensor, max_len: int, bos: int = 998, eos: Optional[int] = 997,
                           gaps: Optional[list] = None, logits_out: Optional[list] = None, select=None) -> torch.Tensor:
    """Same token sequence as ``generate_greedy_recompute`` while ``max_len <= decoder max_len``
    (token-identical in fp32), with per-layer K/V kept between steps and the encoder
    memory projected once.  Used where the O(T^2) reference loop would take minutes."""
    B, S, D = enc.shape
    kinds = _dec_kinds(sd)
    P = "decoder.net.attn_layers"
    if max_len > sd["decoder.net.pos_embedding.embedding.weight"].shape[0]:
        raise ValueError("cached decode is only position-exact while max_len <= decoder max_len")
    g, b = sd[P + ".layers.0.0.weight"], sd[P + ".layers.0.0.bias"]
    heads = lambda t: t.reshape(B, -1, HEADS, DIM_HEAD).permute(0, 2, 1, 3)
    cross_kv, self_k, self_v = {}, {}, {}
    for i, kind in enumerate(kinds):
        if kind == "cross":
            lp = f"{P}.layers.{i}.1"
            cross_kv[i] = (heads(F.linear(enc, sd[lp + ".k.weight"])), heads(F.linear(enc, sd[lp + ".v.weight"])))
    tok = torch.full((B,), bos, dtype=torch.long)
    out_ids: List[torch.Tensor] = []
    seen_eos = torch.zeros(B, dtype=torch.bool)
    for t in range(max_len):
        x = sd["decoder.net.token_embedding.weight"][tok] + sd["decoder.net.pos_embedding.embedding.weight"][t]
        x = x[:, None, :]
        for i, kind in enumerate(kinds):
            lp = f"{P}.layers.{i}.1"
            residual = x
            x = F.layer_norm(x, (D,), g, b, 1e-5)
            if kind == "mlp":
                out = mlp_geglu(x, sd, lp)
            else:
                q = heads(F.linear(x, sd[lp + ".q.weight"]))
                if kind == "self":
                    k_new, v_new = heads(F.linear(x, sd[lp + ".k.weight"])), heads(F.linear(x, sd[lp + ".v.weight"]))
                    self_k[i] = k_new if t == 0 else torch.cat((self_k[i], k_new), dim=2)
                    self_v[i] = v_new if t == 0 else torch.cat((self_v[i], v_new), dim=2)
                    k, v = self_k[i], self_v[i]
                else:
                    k, v = cross_kv[i]
                attn = F.softmax(torch.einsum("bhid,bhjd->bhij", q, k) * (DIM_HEAD ** -0.5), dim=-1)
                o = torch.einsum("bhij,bhjd->bhid", attn, v).permute(0, 2, 1, 3).reshape(B, 1, HEADS * DIM_HEAD)
                out = F.glu(F.linear(o, sd[lp + ".fc_out.0.weight"], sd[lp + ".fc_out.0.bias"]), dim=-1)
            x = out + residual
            if i != len(kinds) - 1:
                x = F.layer_norm(x, (D,), g, b, 1e-5)
        x = F.layer_norm(x, (D,), sd["decoder.net.norm.weight"], sd["decoder.net.norm.bias"], 1e-5)
        logits = F.linear(x[:, 0], sd["decoder.net.to_logits.weight"], sd["decoder.net.to_logits.bias"])
        tok = logits.argmax(dim=-1) if select is None else select(logits, t)      # select: sampling (below)
        if gaps is not None:
            top2 = logits.topk(2, dim=-1).values
            gaps.append((top2[:, 0] - top2[:, 1]).clone())
        if logits_out is not None:
            logits_out.append(logits.clone())
        out_ids.append(tok)
        if eos is not None:
            seen_eos |= tok == eos
            if bool(seen_eos.all()):
                break
    return torch.stack(out_ids, dim=1)


def model_generate(sd: SD, src: torch.Tensor, max_len: int, bos: int = 998, eos: int = 997,
                   kind: str = "hybrid", cached: bool = False, gaps: Optional[list] = None) -> torch.Tensor:
    """model/ocr_model.py:46-66 OCRModel.generate: encoder once, BOS column, decoder loop."""
    with torch.no_grad():
        enc = encoder_forward(sd, src, kind)
        fn = generate_greedy_cached if cached else generate_greedy_recompute
        return fn(sd, enc, max_len, bos, eos, gaps)


# ------------------------------------------------------------------------------------------------ sampling
def topk_filter(logits: torch.Tensor, threshold: float = 0.9) -> torch.Tensor:
    """utils.py:85-91: keep the k = int((1 - threshold) * vocab) largest logits of every row, -inf elsewhere
    (k = 99 for threshold 0.9, vocab 1000: the product is 99.99999999999997 in double arithmetic)."""
    k = int((1 - threshold) * logits.shape[-1])
    val, ind = torch.topk(logits, k)
    out = torch.full_like(logits, float("-inf"))
    out.scatter_(1, ind, val)
    return out


def sample_probs(logits: torch.Tensor, temp: float, threshold: float = 0.9) -> torch.Tensor:
    """model/decoder.py:104-107: the distribution torch.multinomial draws the next token from."""
    return F.softmax(topk_filter(logits, threshold) / temp, dim=-1)


def philox4x32_10(ctr, key):
    """Philox4x32-10 (Salmon et al., SC'11), the counter-based generator of the CUDA path: 4 x uint32 out."""
    M0, M1, W0, W1, MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
    c0, c1, c2, c3 = (int(v) & MASK for v in ctr)
    k0, k1 = (int(v) & MASK for v in key)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def philox_uniform(seed: int, row: int, step: int, call: int = 0) -> float:
    """u in [0, 1) of (row, step) in sampled generate call number `call` (include/texocr.h, texocr_set_sampling)."""
    x = philox4x32_10((row, step, call, 0), (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))[0]
    return (x >> 8) * (1.0 / 16777216.0)


def sample_inverse_cdf(probs: torch.Tensor, u: float) -> int:
    """First vocabulary index whose cumulative probability exceeds u * total (the CUDA path's draw)."""
    c = torch.cumsum(probs.double(), dim=0)
    idx = int(torch.searchsorted(c, torch.tensor(u * float(c[-1]), dtype=torch.float64), right=True))
    return idx if idx < probs.numel() else int(torch.nonzero(probs > 0).flatten()[-1])      # rounding: last kept index


def make_sampler(temp: float, threshold: float = 0.9, seed: int = 0, call: int = 0):
    """`select` callback for generate_greedy_cached: the reference's top-k / temperature draw with the CUDA path's RNG."""
    def select(logits: torch.Tensor, t: int) -> torch.Tensor:
        probs = sample_probs(logits, temp, threshold)
        return torch.tensor([sample_inverse_cdf(probs[r], philox_uniform(seed, r, t, call)) for r in range(logits.shape[0])],
                            dtype=torch.long)
    return select


def batch_acc(pred: torch.Tensor, target: torch.Tensor, pad_token: int) -> float:
    """eval/eval.py:3-33 token accuracy (the consumer of generate's output), without the debug prints."""
    if pred.shape[1] > target.shape[1]:
        target = torch.cat((target, torch.full((target.shape[0], pred.shape[1] - target.shape[1]), pad_token)), dim=1)
    elif pred.shape[1] < target.shape[1]:
        pred = torch.cat((pred, torch.full((pred.shape[0], target.shape[1] - pred.shape[1]), pad_token)), dim=1)
    m = (pred != pad_token) | (target != pad_token)
    return ((((pred == target) & m).sum(1)).float() / m.sum(1).float()).mean().item()
