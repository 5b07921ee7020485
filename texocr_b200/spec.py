"""Parameter inventory of the TeXOCR inference path.

The drop-in boundary keeps the reference's ``state_dict`` key names, so the whole
host side is driven by one table: ``param_table(config)`` lists every key the
reference model exposes (reference ``model/encoder.py:172-191``,
``model/decoder.py:148-173``, ``model/resnet.py:200-254``,
``model/attention.py:183-221``) together with its shape, its initialiser and --
for the aliased keys (the one LayerNorm shared by a whole stack,
``model/attention.py:200,221``; the ``block_list``/``block`` double registration,
``model/resnet.py:130-139``) -- the key it aliases.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

# fixed by the reference's create_encoder (model/encoder.py:172-191)
IMG_MAX_H, IMG_MAX_W = 160, 1008
STAGE_DEPTHS = (2, 4, 6)
STAGE_CHANNELS = (256, 512, 1024)
STEM_CHANNELS = 64
GN_GROUPS = 32
DIM_HEAD = 64            # model/attention.py:76
VOCAB_DEFAULT = 1000     # tokenizer/tokenizer_clean_1k.txt:1


@dataclass(frozen=True)
class ParamSpec:
    key: str
    shape: Tuple[int, ...]
    init: str                 # conv | linear_w | linear_b | ones | zeros | normal02
    fan_in: int = 0
    alias_of: Optional[str] = None


@dataclass(frozen=True)
class ModelDims:
    patch_size: int
    enc_dim: int
    enc_layers: int
    enc_heads: int
    dec_dim: int
    dec_layers: int
    dec_heads: int
    dec_exp: int
    vocab: int
    max_length: int
    bos: int
    eos: int
    pad: int
    encoder_kind: str = "hybrid"     # "hybrid" (ResNetV2 stem, the default) | "patch"

    @property
    def grid_h(self) -> int:
        return IMG_MAX_H // self.patch_size

    @property
    def grid_w(self) -> int:
        return IMG_MAX_W // self.patch_size


def dims_from_config(config: dict, encoder_kind: str = "hybrid") -> ModelDims:
    """Read exactly the keys the reference reads."""
    if "max_length" not in config:
        raise AssertionError("max_length not loaded into config file!")   # model/decoder.py:150
    if "vocab_size" not in config:
        raise AssertionError("vocab_size not loaded into config file!")   # model/decoder.py:151
    enc, dec = config["encoder"], config["decoder"]
    if not config.get("glu", True):
        raise ValueError("texocr_b200 implements the GeGLU MLP only (config glu: true)")
    if not dec.get("cross_attend", True):
        raise ValueError("texocr_b200 implements the cross-attending decoder only")
    if int(enc.get("n_channels", 1)) != 1:
        raise ValueError("texocr_b200 supports single-channel images only (encoder.n_channels: 1)")
    d = ModelDims(
        patch_size=int(config["patch_size"]),
        enc_dim=int(enc["embed_dim"]), enc_layers=int(enc["num_layers"]), enc_heads=int(enc["heads"]),
        dec_dim=int(dec["embed_dim"]), dec_layers=int(dec["num_layers"]), dec_heads=int(dec["heads"]),
        dec_exp=int(dec["exp_factor"]),
        vocab=int(config["vocab_size"]), max_length=int(config["max_length"]),
        bos=int(config["bos_token"]), eos=int(config["eos_token"]), pad=int(config["trg_pad_idx"]),
        encoder_kind=encoder_kind,
    )
    if d.patch_size != 16:
        raise ValueError("texocr_b200 is built for patch_size 16 (backbone stride 16)")
    if d.enc_dim != 256 or d.dec_dim != 256 or d.enc_heads != 8 or d.dec_heads != 8 or d.dec_exp != 4:
        raise ValueError("texocr_b200 kernels are specialised for embed_dim 256, 8 heads, exp_factor 4")
    return d


def default_config(max_length: int = 256, vocab_size: int = VOCAB_DEFAULT) -> dict:
    """The hot-path subset of the reference's config/config.yml (values from config.yml:1-45)."""
    return {
        "bos_token": 998, "eos_token": 997, "trg_pad_idx": 999, "src_pad_idx": 1,
        "patch_size": 16, "glu": True, "device": "cuda",
        "encoder": {"dropout": 0.1, "embed_dim": 256, "exp_factor": 4, "heads": 8,
                    "n_channels": 1, "num_layers": 4},
        "decoder": {"cross_attend": True, "dropout": 0.1, "embed_dim": 256, "exp_factor": 4,
                    "heads": 8, "num_layers": 4},
        "max_length": max_length, "vocab_size": vocab_size,
    }


def _attention_block(prefix: str, dim: int, heads: int) -> List[ParamSpec]:
    inner = heads * DIM_HEAD
    return [
        ParamSpec(f"{prefix}.q.weight", (inner, dim), "linear_w", dim),
        ParamSpec(f"{prefix}.k.weight", (inner, dim), "linear_w", dim),
        ParamSpec(f"{prefix}.v.weight", (inner, dim), "linear_w", dim),
        ParamSpec(f"{prefix}.fc_out.0.weight", (2 * dim, inner), "linear_w", inner),
        ParamSpec(f"{prefix}.fc_out.0.bias", (2 * dim,), "linear_b", inner),
    ]


def _mlp_block(prefix: str, dim: int, exp: int) -> List[ParamSpec]:
    hid = dim * exp
    return [
        ParamSpec(f"{prefix}.fc_in.fc.weight", (2 * hid, dim), "linear_w", dim),
        ParamSpec(f"{prefix}.fc_in.fc.bias", (2 * hid,), "linear_b", dim),
        ParamSpec(f"{prefix}.fc_out.weight", (dim, hid), "linear_w", hid),
        ParamSpec(f"{prefix}.fc_out.bias", (dim,), "linear_b", hid),
    ]


def _attn_layers(prefix: str, dim: int, heads: int, kinds: Tuple[str, ...], exp: int) -> List[ParamSpec]:
    out: List[ParamSpec] = []
    for i, kind in enumerate(kinds):
        alias_w = None if i == 0 else f"{prefix}.layers.0.0.weight"
        alias_b = None if i == 0 else f"{prefix}.layers.0.0.bias"
        out.append(ParamSpec(f"{prefix}.layers.{i}.0.weight", (dim,), "ones", alias_of=alias_w))
        out.append(ParamSpec(f"{prefix}.layers.{i}.0.bias", (dim,), "zeros", alias_of=alias_b))
        if kind == "mlp":
            out += _mlp_block(f"{prefix}.layers.{i}.1", dim, exp)
        else:
            out += _attention_block(f"{prefix}.layers.{i}.1", dim, heads)
    return out


def backbone_conv_plan() -> List[dict]:
    """The 40 weight-standardised convolutions of ResNetV2([2,4,6]) in execution order.

    Each entry: name (state_dict prefix relative to backbone_net), cin, cout, k, stride,
    gn (prefix of the GroupNorm that follows), act (ReLU after the GN?).
    Follows model/resnet.py:100-149 (Bottleneck), 152-197 (Stage), 200-254 (ResNetV2).
    """
    plan = [dict(name="stem.0", cin=1, cout=STEM_CHANNELS, k=7, stride=2, gn="stem.1", act=True)]
    prev = STEM_CHANNELS
    for s, (depth, cout) in enumerate(zip(STAGE_DEPTHS, STAGE_CHANNELS)):
        mid = cout // 4
        for b in range(depth):
            stride = (1 if s == 0 else 2) if b == 0 else 1
            p = f"stages.{s}.stage_blocks.{b}"
            if b == 0:
                plan.append(dict(name=f"{p}.downsample.conv", cin=prev, cout=cout, k=1, stride=stride,
                                 gn=f"{p}.downsample.norm", act=False))
            plan.append(dict(name=f"{p}.block_list.0", cin=prev, cout=mid, k=1, stride=1,
                             gn=f"{p}.block_list.1", act=True))
            plan.append(dict(name=f"{p}.block_list.2", cin=mid, cout=mid, k=3, stride=stride,
                             gn=f"{p}.block_list.3", act=True))
            plan.append(dict(name=f"{p}.block_list.4", cin=mid, cout=cout, k=1, stride=1,
                             gn=f"{p}.block_list.5", act=False))
            prev = cout
    return plan


def param_table(d: ModelDims) -> List[ParamSpec]:
    t: List[ParamSpec] = []
    E = d.enc_dim
    # hybrid: create_encoder fixes img_size=(160,1008) -> 10x63 grid; the patch variant can only be built with
    # an int img_size (PatchEmbedding does img_size // patch_size, model/encoder.py:23) -> 63x63 grid.  The pos-id
    # formula r*63+c+1 is the same for both (model/encoder.py:137-141).
    n_pos = (d.grid_h if d.encoder_kind == "hybrid" else d.grid_w) * d.grid_w + 1
    t.append(ParamSpec("encoder.cls_token", (1, 1, E), "zeros"))
    t.append(ParamSpec("encoder.pos_embed", (1, n_pos, E), "zeros"))
    if d.encoder_kind == "hybrid":
        bb = "encoder.patch_embed.backbone_net"
        for c in backbone_conv_plan():
            fan = c["cin"] * c["k"] * c["k"]
            shape = (c["cout"], c["cin"], c["k"], c["k"])
            t.append(ParamSpec(f"{bb}.{c['name']}.weight", shape, "conv", fan))
            t.append(ParamSpec(f"{bb}.{c['gn']}.weight", (c["cout"],), "ones"))
            t.append(ParamSpec(f"{bb}.{c['gn']}.bias", (c["cout"],), "zeros"))
            if ".block_list." in c["name"]:     # nn.Sequential(*block_list) re-registers the same modules
                for suffix, nm in ((".weight", c["name"]), (".weight", c["gn"]), (".bias", c["gn"])):
                    src = f"{bb}.{nm}{suffix}"
                    dst = src.replace(".block_list.", ".block.")
                    base = next(p for p in t if p.key == src)
                    t.append(ParamSpec(dst, base.shape, base.init, base.fan_in, alias_of=src))
        t.append(ParamSpec("encoder.patch_embed.proj.weight", (E, STAGE_CHANNELS[-1], 1, 1), "conv", STAGE_CHANNELS[-1]))
        t.append(ParamSpec("encoder.patch_embed.proj.bias", (E,), "linear_b", STAGE_CHANNELS[-1]))
    else:
        ps = d.patch_size
        t.append(ParamSpec("encoder.patch_embed.proj.weight", (E, 1, ps, ps), "conv", ps * ps))
        t.append(ParamSpec("encoder.patch_embed.proj.bias", (E,), "linear_b", ps * ps))
    # encoder MLP always uses the class defaults glu=True, exp_factor=4 (model/encoder.py:182-190)
    t += _attn_layers("encoder.attn_layers", E, d.enc_heads, ("self", "mlp") * d.enc_layers, 4)
    t.append(ParamSpec("encoder.norm.weight", (E,), "ones"))
    t.append(ParamSpec("encoder.norm.bias", (E,), "zeros"))
    D = d.dec_dim
    t.append(ParamSpec("decoder.net.token_embedding.weight", (d.vocab, D), "normal02"))
    t.append(ParamSpec("decoder.net.pos_embedding.embedding.weight", (d.max_length, D), "normal02"))
    t.append(ParamSpec("decoder.net.norm.weight", (D,), "ones"))
    t.append(ParamSpec("decoder.net.norm.bias", (D,), "zeros"))
    t += _attn_layers("decoder.net.attn_layers", D, d.dec_heads, ("self", "cross", "mlp") * d.dec_layers, d.dec_exp)
    t.append(ParamSpec("decoder.net.to_logits.weight", (d.vocab, D), "linear_w", D))
    t.append(ParamSpec("decoder.net.to_logits.bias", (d.vocab,), "linear_b", D))
    return t


def unique_params(table: List[ParamSpec]) -> Dict[str, ParamSpec]:
    return {p.key: p for p in table if p.alias_of is None}
