"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

Run here only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference uses absolute ``TeXOCR.*`` imports, so it is imported through a symlink
``<tmp>/TeXOCR -> /root/reference``; nothing is copied.  Weights are the numpy-seeded
``texocr_b200.synth.seeded_state_dict`` (seed 0, re-randomised cls/pos/LN/GN, seed 123) loaded
with ``strict=True`` -- which also pins the state_dict key names and shapes of
``texocr_b200.spec.param_table`` against the reference.  Greedy = ``torch.multinomial`` replaced by
argmax around the reference's own ``generate``.
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from texocr_b200 import spec, synth  # noqa: E402


def import_reference():
    tmp = tempfile.mkdtemp(prefix="texocr_ref_")
    os.symlink("/root/reference", os.path.join(tmp, "TeXOCR"))
    sys.path.insert(0, tmp)
    import TeXOCR.model as M  # noqa
    return M


def greedy(fn, *a, **kw):
    orig = torch.multinomial
    torch.multinomial = lambda p, n, **k: p.argmax(-1, keepdim=True)
    try:
        return fn(*a, **kw)
    finally:
        torch.multinomial = orig


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    M = import_reference()
    cfg = spec.default_config(max_length=256, vocab_size=1000)
    cfg["device"] = "cpu"
    d = spec.dims_from_config(cfg)
    sd = synth.seeded_state_dict(d, seed=0)
    model = M.create_model(cfg)
    ref_keys = list(model.state_dict().keys())
    assert sorted(ref_keys) == sorted(sd.keys()), "state_dict key mismatch vs reference"
    model.load_state_dict(sd, strict=True)
    model.eval()
    out = {}
    with torch.no_grad():
        # --- encoder pins: three shapes incl. a non-64-multiple width and the maximum size
        for name, (B, H, W, dense) in {"a": (2, 64, 384, False), "b": (1, 48, 208, True), "c": (1, 160, 1008, False)}.items():
            img = synth.synth_images(B, H, W, seed=1234 + ord(name), dense=dense)
            enc = model.encoder(img)
            if name == "c":
                enc = enc[:, ::6]
            out[f"enc_{name}"] = enc.numpy()
            out[f"enc_{name}_shape"] = np.array([B, H, W, int(dense), 1234 + ord(name)])
            if name == "a":
                feat = model.encoder.patch_embed.backbone_net(img)
                out["backbone_a_sub"] = feat[:, ::16].numpy()          # every 16th channel
                enc_a, img_a = model.encoder(img), img
        # --- teacher-forced logits + loss (config-4 style, small): labels with padding
        trg = synth.synth_labels(2, 33, d, seed=4321, min_len=8)
        loss, logits = model.decoder(trg, enc=enc_a, mask=model.make_trg_mask(trg), return_out=True)
        out["tf_trg"] = trg.numpy()
        out["tf_logits"] = logits.numpy()
        out["tf_loss"] = np.array(loss.item())
        out["fwd_loss"] = np.array(model(img_a, trg).item())
        # a label batch where one row is entirely padding after BOS (fully-masked query rows)
        trg2 = trg.clone()
        trg2[1, 3:] = d.pad
        _, logits2 = model.decoder(trg2, enc=enc_a, mask=model.make_trg_mask(trg2), return_out=True)
        out["tf2_trg"] = trg2.numpy()
        out["tf2_logits_sub"] = logits2[:, :, ::8].numpy()
        # --- greedy generate, BASELINE config 1: B=8, 64x384, max_len 256
        img8 = synth.synth_images(8, 64, 384, seed=1234)
        tokens = greedy(model.generate, img8, max_len=256)
        out["gen8_tokens"] = tokens.numpy().astype(np.int16)
        # top-2 gaps of the reference's own decisions (for the tie-aware comparison)
        enc8 = model.encoder(img8)
        full = model.decoder.net(torch.cat((torch.full((8, 1), d.bos), tokens[:, :-1]), 1),
                                 mask=torch.ones(8, 256, dtype=torch.bool), enc=enc8)
        top2 = full.topk(2, dim=-1).values
        out["gen8_gaps"] = (top2[..., 0] - top2[..., 1]).numpy()
        assert (full.argmax(-1) == tokens).float().mean() > 0.99
        # --- early exit: pick as EOS a token every row emits early, loop must stop there
        first = tokens[:, :64]
        common = []
        for t in sorted(set(first[0].tolist())):
            occ = [(first[r] == t).nonzero() for r in range(8)]
            if all(len(o) for o in occ):
                done = max(int(o[0]) for o in occ)          # step at which the last row first emits t
                if 4 <= done <= 48 and t != d.bos:
                    common.append((done, int(t)))
        common = [t for _, t in sorted(common)]
        if common:
            eos = common[0]
            early = greedy(model.decoder.generate, start_tokens=torch.full((8, 1), d.bos), eos_tok=eos,
                           max_len=256, temp=0.3, enc=enc8)
            out["early_eos"] = np.array(eos)
            out["early_tokens"] = early.numpy().astype(np.int16)
        # --- pure patch-embed ViT variant (model/encoder.py:11-28)
        torch.manual_seed(0)
        pe = M.VisionEncoder(img_size=1008, patch_size=16, in_channels=1, embed_dim=256, num_layers=4, heads=8)
        d_p = spec.dims_from_config(cfg, encoder_kind="patch")
        sd_p = {k[len("encoder."):]: v for k, v in synth.seeded_state_dict(d_p, seed=0).items() if k.startswith("encoder.")}
        pe.load_state_dict(sd_p, strict=True)
        pe.eval()
        imgp = synth.synth_images(2, 64, 384, seed=99)
        out["enc_patch"] = pe(imgp).numpy()
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, v.dtype)
    print("n_steps early:", out.get("early_tokens", np.zeros((0, 0))).shape)


if __name__ == "__main__":
    main()
