"""Generate tests/golden/golden_next_v1.npz + golden_tokenizer_v1.json for the SURVEY.md 8(f) rows by running the
UNMODIFIED reference in the build container (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden_next.py

f1 sampling: the reference's own ``generate`` loop (model/decoder.py:77-122: topk -> softmax(/temp) -> multinomial)
with ``torch.multinomial`` replaced by the inverse CDF at the Philox uniforms of include/texocr.h, so the draw is
reproducible; also raw ``topk`` / softmax vectors of utils.py:85-91.
f2 detokeniser: ``RegExTokenizer`` (tokenizer/tokenizer.py) loaded from the reference's trained vocabulary file: the
id -> bytes table, decode() of random id rows, encode()/decode() of LaTeX strings, ``process_output`` pairs (utils.py:73-79).
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import texocr_oracle as O  # noqa: E402
from texocr_b200 import spec, synth  # noqa: E402

LATEX = [
    r"\frac { a } { b } + \sqrt { x ^ { 2 } + y ^ { 2 } }",
    r"\int _ { 0 } ^ { \infty } e ^ { - x ^ { 2 } } d x = \frac { \sqrt { \pi } } { 2 }",
    r"\sum _ { n = 1 } ^ { N } \alpha _ { n } \mathbf { x } _ { n } \leq \lambda _ { \max } ( A )",
    r"E = m c ^ { 2 } , \quad \hbar \omega = k _ { B } T",
    r"\left( \begin{array} { c c } 1 & 0 \\ 0 & 1 \end{array} \right) \psi ( x , t ) = 0",
    "x ∈ ℝ , café \\to 7 890 12",
]


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    tmp = tempfile.mkdtemp(prefix="texocr_ref_")
    os.symlink("/root/reference", os.path.join(tmp, "TeXOCR"))
    sys.path.insert(0, tmp)
    import TeXOCR.model as M  # noqa
    from TeXOCR.tokenizer.tokenizer import RegExTokenizer
    from TeXOCR.utils import process_output, topk

    out = {}
    # ---------------------------------------------------------------- f1: sampling
    rng = np.random.Generator(np.random.PCG64(99))
    logits = torch.from_numpy(rng.standard_normal((6, 1000)).astype(np.float32) * 3.0)
    filt = topk(logits.clone())                              # reference top-k filter (threshold 0.9)
    out["samp_logits"] = logits.numpy()
    out["samp_kept"] = torch.isfinite(filt).numpy()
    out["samp_probs_t03"] = F.softmax(filt / 0.3, dim=-1).numpy()
    out["samp_probs_t10"] = F.softmax(filt / 1.0, dim=-1).numpy()

    cfg = spec.default_config(max_length=256, vocab_size=1000)
    cfg["device"] = "cpu"
    d = spec.dims_from_config(cfg)
    model = M.create_model(cfg)
    model.load_state_dict(synth.seeded_state_dict(d, seed=0), strict=True)
    model.eval()
    seed, temp = 2024, 0.3
    step = {"t": 0}

    def philox_multinomial(p, n, **kw):                     # replaces only the random draw of the reference loop
        t = step["t"]
        step["t"] += 1
        return torch.tensor([[O.sample_inverse_cdf(p[r], O.philox_uniform(seed, r, t, 0))] for r in range(p.shape[0])])

    img8 = synth.synth_images(8, 64, 384, seed=1234)
    orig = torch.multinomial
    torch.multinomial = philox_multinomial
    try:
        with torch.no_grad():
            toks = model.generate(img8, max_len=40, temp=temp)
    finally:
        torch.multinomial = orig
    out["samp_gen_tokens"] = toks.numpy().astype(np.int16)
    out["samp_gen_meta"] = np.array([seed, 40, 8])
    out["samp_gen_temp"] = np.array(temp)
    # distance of every draw to the nearest CDF boundary (tie-awareness of the comparison): recompute the probabilities
    with torch.no_grad():
        enc8 = model.encoder(img8)
        full = model.decoder.net(torch.cat((torch.full((8, 1), d.bos), toks[:, :-1]), 1),
                                 mask=torch.ones(8, toks.shape[1], dtype=torch.bool), enc=enc8)
    margins = np.zeros(toks.shape, dtype=np.float64)
    for r in range(8):
        for t in range(toks.shape[1]):
            p = F.softmax(topk(full[r:r + 1, t].clone()) / temp, dim=-1)[0].double()
            c = torch.cumsum(p, 0)
            u = O.philox_uniform(seed, r, t, 0) * float(c[-1])
            k = int(toks[r, t])
            lo = float(c[k] - p[k])
            margins[r, t] = min(u - lo, float(c[k]) - u)
    out["samp_gen_margin"] = margins
    np.savez_compressed(os.path.join(HERE, "golden_next_v1.npz"), **out)

    # ---------------------------------------------------------------- f2: detokeniser
    tok = RegExTokenizer()
    tok.load("/root/reference/tokenizer/tokenizer_clean_1k.txt")
    vocab = {int(i): list(tok.vocab[i]) for i in sorted(tok.vocab)}
    ids_rng = np.random.Generator(np.random.PCG64(5))
    rows = [[int(v) for v in ids_rng.integers(0, 997, size=n)] for n in (1, 7, 40, 128)]
    enc_rows = [tok.encode(s) for s in LATEX]
    js = {
        "vocab_size": tok.vocab_size,
        "special_tokens": tok.special_tokens,
        "vocab_bytes": vocab,
        "bp_merges": [[int(i), int(j), int(t)] for (i, j), t in tok.bp_merges.items()],
        "random_ids": rows,
        "random_decoded": [tok.decode(r) for r in rows],
        "latex": LATEX,
        "latex_ids": enc_rows,
        "latex_decoded": [tok.decode(r) for r in enc_rows],
        "latex_processed": [process_output(tok.decode(r)) for r in enc_rows],
        "process_in": [" \\alpha   x \\beta 2 \n y", "a  b\tc", "\\frac { 1 } { 2 } \\pi r", "\\mathbf  A \\cdot 3"],
    }
    js["process_out"] = [process_output(s) for s in js["process_in"]]
    with open(os.path.join(HERE, "golden_tokenizer_v1.json"), "w") as f:
        json.dump(js, f)
    # ---------------------------------------------------------------- f4: input side
    from PIL import Image
    from torchvision import transforms
    from TeXOCR.data_wrangling.dataset import BucketBatchSampler, img_transform
    det = transforms.Compose(img_transform.transforms[1:])          # ToTensor -> Grayscale(1) -> Invert (RandomAffine skipped)
    assert [type(t).__name__ for t in det.transforms] == ["ToTensor", "Grayscale", "Invert"]
    prep = {}
    irng = np.random.Generator(np.random.PCG64(17))
    for i, (h, w, c) in enumerate([(37, 50, 3), (64, 384, 1), (48, 200, 3), (16, 16, 1)]):
        arr = irng.integers(0, 256, size=(h, w, c), dtype=np.uint8)
        pil = Image.fromarray(arr[..., 0], mode="L") if c == 1 else Image.fromarray(arr, mode="RGB")
        prep[f"img{i}_u8"] = arr
        prep[f"img{i}_f32"] = det(pil).numpy()
    class FakeSet:                                                   # what BucketBatchSampler reads from a dataset
        def __init__(self, sizes):
            from collections import defaultdict
            self.sizes = defaultdict(list)
            for i, s_ in enumerate(sizes):
                self.sizes[s_].append(i)
        def __len__(self):
            return sum(len(v) for v in self.sizes.values())
    size_pool = [(384, 64), (208, 48), (1008, 160), (128, 32)]
    sizes = [size_pool[int(v)] for v in irng.integers(0, 4, size=41)]
    prep["bucket_sizes"] = np.array(sizes)
    for name, kw in {"plain": dict(keep_small=True, shuffle=False), "drop": dict(keep_small=False, shuffle=False),
                     "shuf": dict(keep_small=True, shuffle=True)}.items():
        smp = BucketBatchSampler(FakeSet(sizes), batch_size=4, drop_last=False, seed=3, **kw)
        for epoch in range(2):
            batches = list(iter(smp))
            flat = np.full((len(batches), 4), -1, dtype=np.int64)
            for bi, b in enumerate(batches):
                flat[bi, :len(b)] = b
            prep[f"bucket_{name}_e{epoch}"] = flat
        prep[f"bucket_{name}_len"] = np.array(len(smp))
    np.savez_compressed(os.path.join(HERE, "golden_prep_v1.npz"), **prep)
    print("wrote golden_next_v1.npz", {k: v.shape for k, v in out.items()})
    print("tokenizer vocab", len(vocab), "latex rows", [len(r) for r in enc_rows])
    print("sampled tokens row0", toks[0, :16].tolist(), "min margin", margins.min())


if __name__ == "__main__":
    main()
