"""CPU tier: host logic, the C-ABI library's exports, and the N>1 sharding path over gloo (world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    from texocr_b200 import _lib
    header = open(os.path.join(ROOT, "include", "texocr.h")).read()
    declared = set(re.findall(r"TEXOCR_API\s+[\w\s\*]+?\b(texocr_\w+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    if not os.path.exists(_lib.LIB_PATH):
        from texocr_b200.build import build
        build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name


def test_create_without_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from texocr_b200 import _lib
    lib = _lib.load_library()
    cfg = _lib.TexocrConfig(_lib.ABI_VERSION, 1000, 256, 4, 4, 998, 997, 999, 0, 0)
    h = ctypes.c_void_p()
    rc = lib.texocr_create(ctypes.byref(cfg), 0, ctypes.byref(h))
    assert rc == -5 and b"no CPU fallback" in lib.texocr_last_error(None)


def test_model_surface_and_state_dict_roundtrip(dims, sd):
    import texocr_b200
    from texocr_b200 import spec
    cfg = spec.default_config()
    cfg["device"] = "cpu"
    m = texocr_b200.create_model(cfg)
    assert set(m.state_dict().keys()) == set(sd.keys())
    m.load_state_dict(sd, strict=True)
    assert m.bos_token == 998 and m.eos_token == 997 and m.trg_pad_idx == 999
    assert m.decoder.max_len == 256
    assert torch.equal(m.state_dict()["decoder.net.to_logits.bias"], sd["decoder.net.to_logits.bias"])
    trg = torch.tensor([[998, 5, 997, 999]])
    assert m.make_trg_mask(trg).tolist() == [[True, True, True, False]]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.generate(torch.zeros(1, 1, 64, 384), 4)
    with pytest.raises(AssertionError):
        spec.dims_from_config({k: v for k, v in cfg.items() if k != "max_length"})
    with pytest.raises(ValueError):
        spec.dims_from_config({**cfg, "glu": False})


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under texocr_b200/ may import or execute it."""
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle[/.]texocr_oracle|texocr_oracle", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "texocr_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), (dirpath, f)


def test_shard_plan_and_gloo_gather_world2(tmp_path):
    """bench.py's data-parallel plumbing on CPU: contiguous shards per rank, gather of token ids == concatenation."""
    script = tmp_path / "w2.py"
    script.write_text(f"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {ROOT!r})
from bench import shard_range, gather_tokens
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
lo, hi = shard_range(1000, r, w)
tok = torch.arange(lo * 4, hi * 4, dtype=torch.int64).reshape(hi - lo, 4)
allt = gather_tokens(tok, w)
if r == 0:
    assert allt.shape == (1000, 4) and torch.equal(allt.reshape(-1), torch.arange(4000)), allt
    print("GATHER_OK")
dist.destroy_process_group()
""")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert "GATHER_OK" in out.stdout, out.stdout + out.stderr


def test_strong_scaling_job_plan_and_verify_logic_world2(tmp_path):
    """BASELINE configs[4] plumbing on CPU (gloo, world 2): a fixed job is cut into batches with fixed boundaries, every rank
    takes a contiguous run of whole batches, the gathered block of the peer rank equals what rank 0 computes itself for that
    batch (the check bench.py --verify does with the real decode) -- with a deterministic stand-in for the decode."""
    from bench import job_plan
    for world in (1, 2, 4, 8):
        plans = [job_plan(65536, 512, r, world) for r in range(world)]
        assert sum(plans, []) == list(range(128))                       # every batch exactly once, in order, contiguous per rank
        assert max(len(p) for p in plans) - min(len(p) for p in plans) <= 1
    assert [len(job_plan(1000, 512, r, 4)) for r in range(4)] == [1, 1, 0, 0]
    script = tmp_path / "v2.py"
    script.write_text(f"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {ROOT!r})
from bench import job_plan, gather_tokens, BATCH_SEEDS
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
def decode(seed):           # stand-in for model.generate on the batch of that seed: a pure function of the seed
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 1000, (16, 8), generator=g)
mine = job_plan(16 * 6, 16, r, w)
blocks = [gather_tokens(decode(1234 + j % BATCH_SEEDS), w) for j in mine]
if r == 0:
    peer = w - 1
    pj = job_plan(16 * 6, 16, peer, w)[0]
    assert torch.equal(blocks[0][peer * 16:(peer + 1) * 16], decode(1234 + pj % BATCH_SEEDS))
    assert torch.equal(blocks[0][:16], decode(1234 + mine[0] % BATCH_SEEDS))
    print("VERIFY_OK")
dist.destroy_process_group()
""")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert "VERIFY_OK" in out.stdout, out.stdout + out.stderr


def test_generate_pipeline_host_logic_with_stand_in_engines():
    """GeneratePipeline's host side (queue, worker threads, ordering, result buffers, error propagation) with stand-in engines:
    no device and no native library involved -- the GPU behaviour is covered by test_pipeline_of_batches_in_flight_matches_generate."""
    import threading
    import time
    import torch
    from texocr_b200.pipeline import GeneratePipeline

    class FakeEngine:
        def __init__(self, tag, delay):
            self.tag, self.delay, self.calls, self.threads = tag, delay, 0, set()

        def generate(self, images, max_len, out=None):
            self.calls += 1
            self.threads.add(threading.get_ident())
            time.sleep(self.delay * (1 + (int(images.sum()) % 3)))            # finish out of order
            if int(images.flatten()[0]) < 0:
                raise RuntimeError("bad batch")
            res = images.reshape(images.shape[0], -1)[:, :1].to(torch.int64).repeat(1, max_len)
            if out is not None:
                out[:, :max_len] = res
                return out[:, :max_len]
            return res

        def kernel_launches(self):
            return self.calls

    engines = [FakeEngine(i, 0.002) for i in range(3)]
    pipe = GeneratePipeline(None, engines=engines)
    batches = [torch.full((2 + i % 3, 1, 2, 2), float(i)) for i in range(11)]
    got = list(pipe.generate_batches(batches, 5))
    assert [int(g[0, 0]) for g in got] == list(range(11)) and all(g.shape == (2 + i % 3, 5) for i, g in enumerate(got))     # submission order
    assert sum(e.calls for e in engines) == 11 and all(e.calls > 0 for e in engines)           # work was shared
    assert all(len(e.threads) == 1 for e in engines) and len({next(iter(e.threads)) for e in engines}) == 3   # one thread per engine
    outs = [torch.zeros((b.shape[0], 8), dtype=torch.int64) for b in batches]
    got2 = list(pipe.generate_batches(batches, 8, outs=outs))
    assert all(g.data_ptr() == o.data_ptr() and int(o[0, 7]) == i for i, (g, o) in enumerate(zip(got2, outs)))
    pipe.warm_up(batches[0], 3)
    assert pipe.kernel_launches() == 11 + 11 + 3
    bad = [batches[0], torch.full((1, 1, 2, 2), -1.0), batches[2]]
    with pytest.raises(RuntimeError, match="bad batch"):
        list(pipe.generate_batches(bad, 4))
    assert int(list(pipe.generate_batches(batches[:2], 2))[1][0, 0]) == 1                      # still usable after an error
    pipe.close()
    assert pipe.engines == []


def test_absorbed_attention_weight_folding_is_the_same_function():
    """The library's own folding code (host only, no device): attention computed from the folded matrices -- Q' = x Wqk^T scored
    against the 256-wide latent rows, C = P Z, y = C Wvo^T -- equals the reference formulation q K^T / softmax / P V / Wo
    (model/attention.py:114-180) on the same bf16-rounded weights, in float64."""
    import ctypes as C
    import numpy as np
    from texocr_b200 import _lib
    lib = _lib.load_library()
    rng = np.random.default_rng(5)
    bf = lambda a: torch.from_numpy(a).to(torch.bfloat16).to(torch.float32).numpy()
    wq, wk, wv = (rng.standard_normal((512, 256)).astype(np.float32) * 0.06 for _ in range(3))
    wo = rng.standard_normal((512, 512)).astype(np.float32) * 0.04
    wqk = np.empty((2048, 256), np.float32)
    wvo = np.empty((512, 2048), np.float32)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.texocr_debug_fold_absorbed(ptr(wq), ptr(wk), ptr(wv), ptr(wo), ptr(wqk), ptr(wvo)) == 0
    x = rng.standard_normal((5, 256))             # query-side inputs (LayerNorm'd rows)
    z = rng.standard_normal((7, 256))             # latent rows (encoder memory / cached LayerNorm'd inputs)
    q64, k64, v64, o64 = (bf(w).astype(np.float64) for w in (wq, wk, wv, wo))
    q = (x @ q64.T).reshape(5, 8, 64)
    k = (z @ k64.T).reshape(7, 8, 64)
    v = (z @ v64.T).reshape(7, 8, 64)
    sc = np.einsum("nhd,mhd->hnm", q, k) * 0.125
    p = np.exp(sc - sc.max(-1, keepdims=True)); p /= p.sum(-1, keepdims=True)
    y = np.einsum("hnm,mhd->nhd", p, v).reshape(5, 512) @ o64.T                      # [5, 512]: value half | gate half
    y_il = np.empty_like(y); y_il[:, 0::2] = y[:, :256]; y_il[:, 1::2] = y[:, 256:]  # the GLU epilogue's (value, gate) interleave
    qa = (x @ wqk.astype(np.float64).T).reshape(5, 8, 256)
    sa = np.einsum("nhc,mc->hnm", qa, z) * 0.125
    pa = np.exp(sa - sa.max(-1, keepdims=True)); pa /= pa.sum(-1, keepdims=True)
    ca = np.einsum("hnm,mc->nhc", pa, z).reshape(5, 2048)
    ya = ca @ wvo.astype(np.float64).T
    assert np.abs(sa - sc).max() < 1e-4 * np.abs(sc).max()                            # fp32 storage of the folded matrices
    assert np.abs(ya - y_il).max() < 1e-4 * np.abs(y_il).max()


def test_bench_roofline_groups_roles_of_one_kernel_and_in_flight_rule():
    """bench.py host logic: the `roofline` object names the dominant KERNEL -- the query projection and the block-diagonal value
    projection are one tc_gemm_kernel instantiation, so their classes are added before the maximum is taken -- and the number of
    batches in flight is chosen so that the job's batches divide over the replicas."""
    sys.path.insert(0, ROOT) if ROOT not in sys.path else None
    import bench
    assert [bench.auto_in_flight(k) for k in (12, 20, 16, 30, 7, 25)] == [6, 5, 8, 6, 7, 5]
    peaks = {"hbm_gbs": 6551.4, "bf16_tflops": 1653.6, "bf16_tflops_sustained": 1372.5}
    mk = lambda name, ms, n, by, fl: {"name": name, "ms": ms, "launches": n, "bytes": by * n, "flops": fl * n}
    rows = [mk("dec_gemm_wo", 21.8, 2048, 2.1e6, 268e6), mk("dec_gemm_q", 19.2, 2048, 3.4e6, 537e6), mk("dec_gemm_vproj", 18.0, 2048, 2.9e6, 134e6),
            mk("dec_attn_self", 17.7, 1024, 33e6, 0.5e9), mk("dec_attn_cross", 16.4, 1024, 25e6, 0.4e9), mk("conv_gemm", 5.2, 39, 6e8, 3.8e10)]
    tot = sum(r["ms"] for r in rows)
    shares = {r["name"]: round(r["ms"] / tot, 4) for r in rows}
    r = bench.dominant_kernel_roofline(rows, bench.class_table(rows, peaks), shares, tot, peaks, "measured", {})
    assert r["class"] == "gemm64_store_bf16" and r["launches"] == 4096 and r["bound"] == "tensor"
    assert abs(r["achieved"] - (537e6 + 134e6) * 2048 / 37.2e-3 / 1e12) < 0.1 and abs(r["frac"] - r["achieved"] / 1372.5) < 1e-3
    assert abs(r["share_of_batch_kernel_time"] - 37.2 / tot) < 1e-3 and r["traffic"] is None
    # one role clearly ahead: no grouping needed
    rows[0]["ms"] = 60.0
    tot = sum(x["ms"] for x in rows)
    r = bench.dominant_kernel_roofline(rows, bench.class_table(rows, peaks), {x["name"]: x["ms"] / tot for x in rows}, tot, peaks, "measured",
                                       {"dec_gemm_wo": {"dram_bytes_per_launch": 1612032}})
    assert r["class"] == "dec_gemm_wo" and r["traffic"] == 1612032
