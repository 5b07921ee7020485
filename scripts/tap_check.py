import sys, torch
sys.path.insert(0, "/root/repo")
import texocr_b200
from texocr_b200 import spec, synth
cfg = spec.default_config(max_length=256); cfg["device"]="cuda:0"
d = spec.dims_from_config(cfg)
m = texocr_b200.create_model(cfg, precision="bf16"); m.load_state_dict(synth.seeded_state_dict(d, seed=0))
eng = m.engine()
steps = 4
for o in sys.argv[1:]:
    k, v = o.split("=")
    if k == "steps": steps = int(v)
    else: eng.set_option(k, int(v))
img = synth.synth_images(512, 64, 384, seed=1234).cuda()
B = 512
sizes = {"logits": B * 1000, "x": B * 256, "s": B * 256, "xn": B * 128, "qkv": B * 768, "o": B * 256, "hid": B * 512,
         "kvcache": 4 * B * steps * 512}
ref = None
for it in range(25):
    tok = m.generate(img, steps)
    taps = {k: eng.debug_read(k, n).view(torch.int32).clone() for k, n in sizes.items()}
    taps["tok"] = tok.clone()
    if ref is None: ref = taps; continue
    diffs = {k: int((taps[k] != ref[k]).sum()) for k in taps}
    if any(diffs.values()):
        msg = {k: v for k, v in diffs.items() if v}
        extra = ""
        if diffs["kvcache"]:
            idx = (taps["kvcache"] != ref["kvcache"]).nonzero().flatten()
            w = idx[0].item()          # word index in [layer][b][head][t][64 words]
            per_l = B * steps * 512
            l, r = divmod(w, per_l); b, r = divmod(r, steps * 512); hh, r = divmod(r, steps * 64); t, c = divmod(r, 64)
            extra = f" first kv diff: layer {l} seq {b} head {hh} key {t} word {c} ({'K' if c < 32 else 'V'}); n={len(idx)}"
        print("run", it, "differs:", msg, extra)
print("done", sys.argv[1:])
