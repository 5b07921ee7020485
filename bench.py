#!/usr/bin/env python
"""Benchmark of the TeXOCR inference hot path on B200 (BASELINE.json metric: decoded equations/sec,
greedy, max_len 256).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference ...                            # the unmodified reference on the host cores (baseline/_ref)

A step = one pass of the hot path over one batch: encoder -> cross-K/V -> 256 greedy decode steps for
B=512 synthetic 64x384 images per GPU (BASELINE.json configs[2]).  The K timed steps are K independent batches; up to
--in-flight of them are decoded concurrently on the GPU (texocr_b200/pipeline.py: one engine handle, host thread and
stream per batch in flight -- one batch alone is bound by the latency of its kernel chain, not by the machine); the
one-batch-at-a-time number is reported next to it.  N>1: one process per GPU (torchrun), the image
list is sharded contiguously, every rank decodes its shard independently and the token ids are gathered
with one NCCL all_gather per step ("scaling": "weak").  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")      # one hardware queue per stream (branches x batches in flight)

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "decoded equations/sec (greedy, max_len 256)"
UNIT = "equations/s"
H, W, MAX_LEN = 64, 384, 256
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
HBM_CLASSES = {"dec_attn_self", "dec_attn_cross", "dec_rowwise", "dec_argmax", "gn_stats", "gn_apply", "enc_rowwise",
               "stem_conv", "tf_rowwise", "misc"}


# ----------------------------------------------------------------------------- data-parallel plumbing (also tested on gloo)
def shard_range(n_items: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of rank `rank`."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


def gather_tokens(tokens: torch.Tensor, world: int) -> torch.Tensor:
    """Token ids of every rank's shard, concatenated in rank order (the path's only collective)."""
    if world == 1:
        return tokens
    import torch.distributed as dist
    outs = [torch.empty_like(tokens) for _ in range(world)]
    dist.all_gather(outs, tokens.contiguous())
    return torch.cat(outs, dim=0)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []
        self.lo, self.hi = 0, None

    def mark_begin(self):      # the timed region starts here: only samples between mark_begin and mark_end are reported
        self.lo = len(self.lines)

    def mark_end(self):
        self.hi = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[self.lo:self.hi]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return {k: float(d[k]) for k in FALLBACK_PEAKS if k in d}, "measured"
        except Exception:
            pass
    return dict(FALLBACK_PEAKS), "fallback"


# ----------------------------------------------------------------------------- CPU arms (the reference itself, else its oracle port)
def _greedy_patch():
    """the reference has no greedy switch; argmax == the temp -> 0 limit of its top-k/softmax/multinomial
    draw, obtained by replacing torch.multinomial around the call (the reference's code is not touched)."""
    orig = torch.multinomial
    torch.multinomial = lambda p, n, **k: p.argmax(-1, keepdim=True)
    return orig


def cpu_reference_eq_per_s(n_eq: int, steps: int = 1, warmup: int = 0):
    """The reference's own CPU path on all host cores: the UNMODIFIED reference from baseline/_ref/TeXOCR
    (`create_model(config).generate(src, max_len)`, model/ocr_model.py:46-66 -- encoder + O(T^2) full-prefix loop without KV
    cache), same seeded weights and synthetic images as the B200 arm.  If that copy is not importable, the oracle port of the
    same algorithm (oracle/texocr_oracle.py) is timed instead and `kind` says so.
    Returns (eq/s, cores, seconds per step, kind, note)."""
    from texocr_b200 import spec, synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = spec.default_config(max_length=MAX_LEN)
    cfg["device"] = "cpu"
    d = spec.dims_from_config(cfg)
    sd = synth.seeded_state_dict(d, seed=0)
    img = synth.synth_images(n_eq, H, W, seed=1234)
    kind, note, run = "reference", "unmodified reference (baseline/_ref/TeXOCR) OCRModel.generate, torch.multinomial -> argmax", None
    try:
        from baseline.install_ref import import_reference
        M = import_reference()
        model = M.create_model(cfg)
        model.load_state_dict(sd, strict=True)
        model.eval()

        def run():
            orig = _greedy_patch()
            try:
                return model.generate(img, max_len=MAX_LEN)
            finally:
                torch.multinomial = orig
    except Exception as ex:      # the copy did not travel / a dependency of the reference is missing: time the port, and say so
        from oracle import texocr_oracle as O
        kind, note = "port", f"oracle port of the reference algorithm (reference not importable: {type(ex).__name__}: {ex})"

        def run():
            return O.model_generate(sd, img, MAX_LEN, d.bos, d.eos, cached=False)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            tok = run()
            dt = time.perf_counter() - t0
            assert tok.shape[0] == n_eq
            if i >= warmup:
                times.append(dt)
    per_step = sum(times) / len(times)
    return n_eq / per_step, cores, per_step, kind, note


def run_reference_arm(args, rank: int):
    if rank != 0:
        return
    total = max(1, args.steps + args.warmup)
    n_eq = max(1, min(8, 160 // total))        # B = 8 (BASELINE config 1) at ~5-10 s per step on 16 cores unless many steps are asked for
    v, cores, per_step, kind, note = cpu_reference_eq_per_s(n_eq, steps=args.steps, warmup=args.warmup)
    sample = (f"B={n_eq} synthetic {H}x{W} images, full {MAX_LEN}-step greedy loop without KV cache, fp32 torch CPU, "
              f"{per_step:.1f} s per step; {note}")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"greedy generate, {H}x{W}, max_len {MAX_LEN}, default config.yml model, random-init weights",
                   "batch_per_step": n_eq},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- the B200 arm
def measure_next_rows(model, img_dev, peaks, stream):
    """SURVEY.md section 8f rows measured next to the headline (not part of the timed region): sampled generate, the
    GPU image transform against the HBM roofline, the host detokeniser."""
    import time
    out = {}
    B = img_dev.shape[0]

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ms = timed(lambda: model.generate(img_dev, MAX_LEN, temp=0.3, sample=True, seed=1), 2)
    out["sampled_generate"] = {"value": B / (ms / 1e3), "unit": UNIT, "ms": ms,
                               "workload": f"same as the headline with top-k(0.9)/temp 0.3 sampling (model/decoder.py:103-108), batch {B}"}
    # image transform: 512 RGB uint8 images 64x384 resident on the device -> float32; bytes = 3 read + 4 written per pixel
    u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=img_dev.device)
    lst = [u for u in u8]
    eng = model.engine()
    ms = timed(lambda: eng.preprocess_u8(lst, 16), 5)
    eng.profile_enable(True)
    eng.preprocess_u8(lst, 16)
    rows = {r["name"]: r for r in eng.profile_read()}
    eng.profile_enable(False)
    k_ms = rows.get("misc", {}).get("ms", float("nan"))
    nbytes = B * H * W * 7
    out["preprocess"] = {"value": B / (ms / 1e3), "unit": "images/s", "ms_call": ms, "kernel_ms": k_ms,
                         "kernel_gbs": nbytes / (k_ms / 1e3) / 1e9, "frac_of_hbm_peak": nbytes / (k_ms / 1e3) / 1e9 / peaks["hbm_gbs"],
                         "workload": f"{B} RGB uint8 {H}x{W} images on the device -> float32 (ToTensor/Grayscale/Invert); the call "
                                     "time includes the Python-side concatenation of the image list"}
    # detokeniser on the host: a synthetic 1k byte-pair vocabulary, (B, MAX_LEN) ids with an EOS per row
    from texocr_b200.detok import Detokenizer
    merges = [(i % 256, (i * 7) % 256, 256 + i) for i in range(741)]
    tok = Detokenizer.from_merges(merges, {"<PAD>": 999, "<BOS>": 998, "<EOS>": 997})
    ids = torch.randint(32, 997, (B, MAX_LEN))
    ids[:, MAX_LEN - 8] = 997
    t0 = time.perf_counter()
    for _ in range(3):
        texts = tok.decode_batch(ids, eos_token=997)
    dt = (time.perf_counter() - t0) / 3
    out["detokenise"] = {"value": B / dt, "unit": "strings/s", "ms": dt * 1e3, "cores": 1,
                         "workload": f"{B} rows x {MAX_LEN - 8} tokens -> text + process_output, host Python"}
    assert len(texts) == B
    return out


def job_plan(total: int, batch: int, rank: int, world: int):
    """Strong-scaling job (BASELINE configs[4]): `total` equations = ceil(total / batch) batches with FIXED boundaries, dealt to the
    ranks as contiguous runs of whole batches, so every batch is decoded by the same kernels on the same rows whatever the
    world size.  Returns the list of global batch indices of this rank."""
    n_batches = (total + batch - 1) // batch
    lo, hi = shard_range(n_batches, rank, world)
    return list(range(lo, hi))


BATCH_SEEDS = 8      # distinct synthetic batches of a strong-scaling job: global batch j holds the images of seed 1234 + j % BATCH_SEEDS


def class_table(rows, peaks):
    """Per kernel class of one eagerly launched batch: time, achieved GB/s and TFLOP/s on the engine's algorithmic bytes / FLOPs
    (DESIGN.md section 5), against both measured peaks."""
    out = {}
    for r in sorted(rows, key=lambda r: -r["ms"]):
        sec = r["ms"] / 1e3
        gbs, tf = r["bytes"] / sec / 1e9, r["flops"] / sec / 1e12
        out[r["name"]] = {"ms": round(r["ms"], 4), "launches": int(r["launches"]), "avg_us": round(r["ms"] * 1e3 / max(1, r["launches"]), 3),
                          "gbs": round(gbs, 1), "frac_hbm": round(gbs / peaks["hbm_gbs"], 4),
                          "tflops": round(tf, 2), "frac_tensor": round(tf / peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]), 4)}
    return out


def auto_in_flight(steps: int) -> int:
    """Batches in flight for a job of `steps` batches per GPU: 6 unless another count in 5..8 leaves fewer replicas idle in the last
    round (ties: nearest to 6).  20 batches over 6 replicas end with two replicas busy for a whole batch latency; over 5 they divide."""
    return min((5, 6, 7, 8), key=lambda n: ((-steps) % n, abs(n - 6)))


def dominant_kernel_roofline(rows, by_class, shares, tot, peaks, peaks_src, traffic_db):
    """`roofline` of the bench line: the kernel with the largest share of one eagerly launched batch (roles that are the same kernel
    instantiation grouped), its algorithmic FLOPs / bytes per launch over its CUDA-event time, against the measured peak."""
    kernel_of = {"dec_gemm_q": "tc_gemm_kernel<64, EPI_STORE, bf16, 1, 0, 2> (absorbed query projection, N = 2048, K = 256; 8 launches per decode step)",
                 "dec_gemm_wo": "tc_gemm_kernel<32, EPI_GLU_RES, float, 1, 0, 2> (attention out-projection + GLU + residual, K = 512; 8 per step)",
                 "dec_gemm_w1": "tc_gemm_kernel<64, EPI_GEGLU, bf16, 1, 0, 2> (MLP in + GeGLU, N = 2048; 4 per step)",
                 "dec_gemm_w2": "tc_gemm_kernel<32, EPI_BIAS_RES, float, 1, 0, 2> (MLP out + residual, K = 1024; 4 per step)",
                 "dec_gemm_vproj": "tc_gemm_kernel<64, EPI_STORE, bf16, 1, 0, 2> block-diagonal (per-head value projection; 8 per step)",
                 "dec_attn_self": "attn_seq_kernel<self> (absorbed decode self-attention, one warp per sequence; 4 per step)",
                 "dec_attn_cross": "attn_seq_kernel<cross> (absorbed decode cross-attention, one warp per sequence; 4 per step)",
                 "conv_gemm": "tc_gemm_persistent_kernel<128 | 256 | 64, EPI_STORE, float, bf16x3> (backbone convolutions, GroupNorm partial sums in the epilogue)",
                 "gn_apply": "gn_apply8_kernel / gn_apply_kernel (GroupNorm apply + residual + ReLU, split-bf16 out)"}
    # the dominant KERNEL: classes are roles, and two roles can be the same kernel instantiation (the absorbed query projection and
    # the block-diagonal value projection are both tc_gemm_kernel<64, EPI_STORE, bf16>; the self / cross attention one template):
    # group by kernel, take the group with the largest share (this is also the top entry of the ncu launch list of a decode step)
    groups = {"dec_gemm_q": "gemm64_store_bf16", "dec_gemm_vproj": "gemm64_store_bf16", "dec_attn_self": "attn_seq", "dec_attn_cross": "attn_seq"}
    by_kernel = {}
    for r in rows:
        by_kernel.setdefault(groups.get(r["name"], r["name"]), []).append(r)
    top_key = max(by_kernel, key=lambda k: sum(r["ms"] for r in by_kernel[k]))
    members = by_kernel[top_key]
    top = {"name": max(members, key=lambda r: r["ms"])["name"], "ms": sum(r["ms"] for r in members), "launches": sum(r["launches"] for r in members),
           "bytes": sum(r["bytes"] for r in members), "flops": sum(r["flops"] for r in members)}
    if len(members) > 1:
        by_class_top = class_table([dict(top, name=top_key)], peaks)[top_key]
        shares[top_key] = round(top["ms"] / tot, 4)
        kernel_of[top_key] = {"gemm64_store_bf16": "tc_gemm_kernel<64, EPI_STORE, bf16, 1, 0, 2> (absorbed query projections N = 2048 and block-diagonal per-head value "
                                                   "projections, K = 256; 16 launches per decode step)",
                              "attn_seq": "attn_seq_kernel<self> + attn_seq_kernel<cross> (absorbed decode attention; 8 launches per decode step)"}[top_key]
        tc = by_class_top
        top["name"] = top_key
    else:
        tc = by_class[top["name"]]
    tensor_bound = top["flops"] > 0 and top["name"] not in HBM_CLASSES
    # measured DRAM bytes per launch of that kernel from the committed ncu --set full capture (profiles/roofline_traffic.json)
    t_ent = traffic_db.get({"dec_attn_self": "attn_seq_kernel", "dec_attn_cross": "attn_seq_kernel", "attn_seq": "attn_seq_kernel"}.get(top["name"], top["name"])) or {}
    ratio = t_ent.get("dram_bytes_over_algorithmic")
    if ratio is None and t_ent.get("dram_bytes_per_launch"):
        ratio = t_ent["dram_bytes_per_launch"] / (top["bytes"] / max(1, top["launches"]))
    if ratio is None and len(members) > 1 and all((traffic_db.get(r["name"]) or {}).get("dram_bytes_per_launch") for r in members):
        ratio = sum(traffic_db[r["name"]]["dram_bytes_per_launch"] * r["launches"] for r in members) / max(1.0, top["bytes"])
    roofline = {"bound": "tensor" if tensor_bound else "hbm", "kernel": kernel_of.get(top["name"], top["name"]), "class": top["name"],
                "achieved": tc["tflops"] if tensor_bound else tc["gbs"],
                "peak": peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]) if tensor_bound else peaks["hbm_gbs"],
                "unit": "TFLOP/s" if tensor_bound else "GB/s",
                "frac": tc["frac_tensor"] if tensor_bound else tc["frac_hbm"],
                "traffic": (ratio * top["bytes"] / max(1, top["launches"])) if ratio else None,
                "peak_source": (peaks_src + " (MEASURED_PEAKS.json, sustained bf16 / hbm_gbs)") if peaks_src == "measured" else "fallback",
                "launches": int(top["launches"]), "avg_us": tc["avg_us"], "share_of_batch_kernel_time": shares[top["name"]],
                "algorithmic_flops_per_launch": top["flops"] / max(1, top["launches"]),
                "algorithmic_bytes_per_launch": top["bytes"] / max(1, top["launches"]),
                "also_gbs": tc["gbs"], "also_frac_hbm": tc["frac_hbm"],
                "method": "CUDA events around every launch of one eagerly launched batch (single branch, nothing else resident); the "
                          "kernel is one 128 x BN tile per CTA: TMA -> 16 tcgen05.mma -> TMEM -> 8 epilogue warps -> store, 4-5 us of dependent latency "
                          "per launch, so its tensor-pipe fraction is small by construction (profiles/r02_*; DESIGN.md section 5)"}
    return roofline


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="equations per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--in-flight", type=int, default=0, help="batches decoded concurrently per GPU (1 = one at a time; 0 = choose 5..8 so "
                    "that the job's batches divide evenly over the replicas: a last round with idle replicas costs more than one replica more or less)")
    ap.add_argument("--branches", type=int, default=1, help="decode branches per batch when several batches are in flight")
    ap.add_argument("--total", type=int, default=0, help="strong scaling (BASELINE configs[4]): a fixed job of this many equations in "
                    "batches of --batch, contiguous runs of batches per rank; --steps is then the number of batches per rank")
    ap.add_argument("--verify", action="store_true", help="rank 0 re-decodes another rank's batch and checks the gathered block")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the roofline / encoder / next-rows measurements after the timed region")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if world != args.gpus and world > 1:
        args.gpus = world
    args.warmup = max(args.warmup, 3)

    import texocr_b200
    from texocr_b200 import spec, synth
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cfg = spec.default_config(max_length=MAX_LEN)
    cfg["device"] = f"cuda:{local_rank}"
    d = spec.dims_from_config(cfg)
    model = texocr_b200.create_model(cfg, precision=args.precision)
    model.load_state_dict(synth.seeded_state_dict(d, seed=0))      # same random-init weights on every rank
    model.eval()
    eng = model.engine()
    B = args.batch
    strong = args.total > 0
    if strong:
        my_batches = job_plan(args.total, B, rank, world)
        n_max = len(job_plan(args.total, B, 0, world))              # rank 0 holds the longest run; shorter ranks pad with repeats
        args.steps = n_max
        seeds = [1234 + (j % BATCH_SEEDS) for j in my_batches] + [1234] * (n_max - len(my_batches))
        pool = {sd_: synth.synth_images(B, H, W, seed=sd_).pin_memory() for sd_ in sorted(set(seeds))}
        host_batches = [pool[sd_] for sd_ in seeds]
    else:
        lo, hi = shard_range(B * world, rank, world)                # this rank's contiguous shard of the job's image list
        host_batches = [synth.synth_images(hi - lo, H, W, seed=1234 + rank).pin_memory()] * args.steps
    dev_pool = {}
    for hb in host_batches:
        if hb.data_ptr() not in dev_pool:
            dev_pool[hb.data_ptr()] = hb.cuda()
    dev_batches = [dev_pool[hb.data_ptr()] for hb in host_batches]
    img_host, img_dev = host_batches[0], dev_batches[0]
    rows_per_batch = img_host.shape[0]
    out_host = torch.empty((rows_per_batch, MAX_LEN), dtype=torch.int64).pin_memory()
    stream = torch.cuda.current_stream()
    if args.in_flight <= 0:
        args.in_flight = auto_in_flight(args.steps)
    n_fly = max(1, min(args.in_flight, args.steps))
    pipe = None
    if n_fly > 1:
        from texocr_b200.pipeline import GeneratePipeline
        pipe = GeneratePipeline(model, in_flight=n_fly, branches=args.branches)
        outs_host = [torch.empty((rows_per_batch, MAX_LEN), dtype=torch.int64).pin_memory() for _ in range(args.steps)]
        # result buffers of the device-resident pass, allocated once: a cudaMalloc from a worker thread in the middle of the timed
        # region (the caching allocator growing its pool for 20 kept results) serialises against the batches in flight
        outs_dev = [torch.empty((rows_per_batch, MAX_LEN), dtype=torch.int64, device="cuda") for _ in range(args.steps)]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, whole=False):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        if whole:
            fn(steps)          # runs all `steps` batches (several in flight); returns after the last one has completed
        else:
            for i in range(steps):
                fn(i)
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    gathered = {}              # step -> (world * rows, MAX_LEN) token ids of the last timed pass (kept for --verify)

    def step_device(i=0):
        tok = model.generate(dev_batches[i % len(dev_batches)], max_len=MAX_LEN)             # public API; inputs resident in HBM
        gathered[i] = gather_tokens(tok, world)

    def step_e2e(i=0):
        tok = eng.generate(host_batches[i % len(host_batches)], MAX_LEN, out=out_host)         # C-ABI with HOST buffers: H2D + D2H inside the call
        if world > 1:
            gather_tokens(tok.cuda(non_blocking=True), world)

    def steps_device(k):       # k batches resident in HBM through the pipeline (public API: GeneratePipeline.generate_batches)
        for i, tok in enumerate(pipe.generate_batches(dev_batches[:k], MAX_LEN, outs=outs_dev[:k])):
            gathered[i] = gather_tokens(tok, world)

    def steps_e2e(k):          # k batches from pinned HOST memory, token ids back to pinned host memory
        for tok in pipe.generate_batches(host_batches[:k], MAX_LEN, outs=outs_host[:k]):
            if world > 1:
                gather_tokens(tok.cuda(non_blocking=True), world)

    # one sampler per job (rank 0's GPU), not per rank: eight nvidia-smi pollers next to eight ranks perturb the launch path.  It is
    # started before the warm-up (nvidia-smi's own start-up takes driver locks for several hundred ms, which used to land inside the
    # timed region and cost up to 20 % of it) and keeps polling until after the end-to-end region; the reported samples are those
    # taken between the two marks around the device-timed region.
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    for i in range(args.warmup):
        step_device(i)
    if pipe is not None:
        pipe.warm_up(img_dev, MAX_LEN)
        for _ in range(args.warmup):
            steps_device(n_fly)
    clocks.mark_begin()
    serial = None
    if pipe is None:
        l0 = eng.kernel_launches()
        ms = timed(step_device, args.steps)
        clocks.mark_end()
        launches = eng.kernel_launches() - l0
    else:
        l0 = pipe.kernel_launches()
        ms = timed(steps_device, args.steps, whole=True)
        clocks.mark_end()
        launches = pipe.kernel_launches() - l0
        if not strong:
            ms_1 = timed(step_device, min(args.steps, 6))
            serial = {"value": world * B * min(args.steps, 6) / (ms_1 / 1e3), "unit": UNIT, "ms_per_step": ms_1 / min(args.steps, 6),
                      "note": "the same batches, one model.generate call at a time (one decode branch per 128 rows): what a caller of the "
                              "reference's own loop (test.py:27-40) gets without the pipeline"}
    done_eq = sum(len(job_plan(args.total, B, r, world)) for r in range(world)) * B if strong else world * B * args.steps
    value = done_eq / (ms / 1e3)

    # ---- --verify: the gathered block of another rank equals a re-decode of that rank's images on THIS GPU (the
    # N-GPU result must equal the 1-GPU result for the same images bit for bit)
    verify = None
    if args.verify:
        if pipe is not None:
            steps_device(min(args.steps, n_fly))       # refresh `gathered` from a pass whose blocks are all kept
        else:
            step_device(0)
        if rank == 0:
            peer = world - 1
            if strong:
                peer_batches = job_plan(args.total, B, peer, world)
                seed = 1234 + (peer_batches[0] % BATCH_SEEDS)
            else:
                seed = 1234 + peer
            again = model.generate(synth.synth_images(rows_per_batch, H, W, seed=seed).cuda(), max_len=MAX_LEN)
            block = gathered[0][peer * rows_per_batch:(peer + 1) * rows_per_batch]
            same = bool(torch.equal(again, block))
            verify = {"ok": same, "peer_rank": peer, "rows": rows_per_batch,
                      "what": "rank 0 re-decoded the first batch of the peer rank from the same seed and compared it with the block the all_gather delivered"}
            if not same:
                raise SystemExit("verify failed: gathered block differs from the re-decode: " + json.dumps(verify))

    e2e = None
    if not args.no_e2e:
        if pipe is None:
            step_e2e()
            ms_e = timed(step_e2e, args.steps)
        else:
            steps_e2e(n_fly)
            ms_e = timed(steps_e2e, args.steps, whole=True)
        e2e = {"value": done_eq / (ms_e / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": int(img_host.numel() * 4 * world), "d2h_bytes_per_step": int(out_host.numel() * 8 * world)}
    clk = clocks.stop()
    if pipe is not None:
        pipe.close()

    # ---- rooflines.  One extra batch is launched eagerly (no CUDA graph, one decode branch, nothing else on the GPU) with CUDA
    # events around every launch on the launching stream (texocr_profile_*); the engine accounts algorithmic bytes / FLOPs per
    # launch.  `roofline` = the kernel with the largest share of that batch's kernel time;
    # `roofline_by_class` = every class against both measured peaks; `roofline_attention` = the HBM-bound decode attention in detail.
    peaks, peaks_src = load_peaks()
    roofline = roofline_attn = by_class = shares = job = None
    encoder = next_rows = None
    if rank == 0 and not args.no_extras:
        eng.set_option("decode_branches", 1)
        eng.set_option("attn_trace", 1)            # in-kernel %globaltimer phase sums of the attention kernel (see below)
        eng.profile_enable(True)
        model.generate(img_dev, max_len=MAX_LEN)
        rows = eng.profile_read()
        eng.profile_enable(False)
        phases, window_s = None, None
        try:
            allraw = eng.debug_read("attn_trace", 16 * 3 * 2048 * 2 + 128).view(torch.int64).cpu()
            tr = allraw[:3 * 2048].reshape(3, 256, 8).double()          # branch 0: entry (min over CTAs), ready (min), end (max)
            ok = tr[2] > 0
            window_s = float(((tr[2] - tr[1]) * ok).sum()) * 1e-9         # launch-wide: first CTA past its dependency -> last CTA done
            raw = allraw[16 * 3 * 2048:].double()
            phases = {k: {"first_stage_us": float(raw[o + 1] / max(1.0, float(raw[o + 4])) / 1e3),
                          "stage_loop_us": float(raw[o + 2] / max(1.0, float(raw[o + 4])) / 1e3),
                          "epilogue_us": float(raw[o + 3] / max(1.0, float(raw[o + 4])) / 1e3)} for k, o in (("self", 0), ("cross", 8))}
        except Exception:
            phases = None
        eng.set_option("attn_trace", 0)
        eng.set_option("decode_branches", 0)
        tot = sum(r["ms"] for r in rows) or 1.0
        shares = {r["name"]: round(r["ms"] / tot, 4) for r in sorted(rows, key=lambda r: -r["ms"])}
        by_class = class_table(rows, peaks)
        traffic_db = {}
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:
                traffic_db = json.load(open(tp))
            except Exception:
                traffic_db = {}
        roofline = dominant_kernel_roofline(rows, by_class, shares, tot, peaks, peaks_src, traffic_db)
        attn = [r for r in rows if r["name"] in ("dec_attn_self", "dec_attn_cross")]
        a_ms = sum(r["ms"] for r in attn)
        a_bytes = sum(r["bytes"] for r in attn)
        a_n = sum(r["launches"] for r in attn)
        if a_n:
            ach = a_bytes / (a_ms / 1e3) / 1e9
            ratio = (traffic_db.get("attn_seq_kernel") or {}).get("dram_bytes_over_algorithmic")
            roofline_attn = {"bound": "hbm", "kernel": "attn_seq_kernel<self> + attn_seq_kernel<cross> (absorbed decode attention, one warp per sequence, 8 launches per decode step)",
                             "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                             "traffic": ratio * a_bytes / a_n if ratio else None, "launches": a_n, "avg_us": a_ms * 1e3 / a_n,
                             "algorithmic_bytes_per_launch": a_bytes / a_n, "share_of_batch_kernel_time": round(a_ms / tot, 4),
                             "bytes_basis": "512 B per key and layer (the 256-wide bf16 latent row all 8 heads share); SURVEY.md 8d counts "
                                            "2,048 B per key and layer for projected K/V -- this formulation moves a quarter of that"}
            if phases and args.precision == "bf16" and window_s:
                by = {r["name"]: r for r in attn}
                loop_s = sum(by[n]["launches"] * phases[k]["stage_loop_us"] * 1e-6 for n, k in (("dec_attn_self", "self"), ("dec_attn_cross", "cross")) if n in by)
                roofline_attn["in_kernel"] = {"achieved": a_bytes / window_s / 1e9, "unit": "GB/s", "frac": a_bytes / window_s / 1e9 / peaks["hbm_gbs"],
                                              "method": "%globaltimer, per launch: first CTA released by its dependency -> last CTA finished (engine option attn_trace)",
                                              "mean_cta_phases_us": phases,
                                              "mean_cta_stage_loop_gbs": a_bytes / loop_s / 1e9 if loop_s > 0 else None}
        # ---- the whole job against the HBM roofline: bytes this implementation has to move per batch (absorbed attention: 2,048 B per
        # cached position / memory token and step = 4 layers x 512 B; weights once per step) and SURVEY.md 8d's bytes for the same work
        s_tok = synth.encoder_tokens(H, W)
        absorbed = args.precision == "bf16"
        row_b = 2048 if absorbed else 16384
        # per-step decoder weights of the absorbed formulation: per layer 2 x Wqk [2048,256] + 2 x Wv [512,256] + 2 x Wo [512,512] + W1 + W2
        w_step = (4 * 2621440 + 257000 + 1024) * 2 if absorbed else 15222736 * 2
        moved = sum(synth.decode_step_bytes(B, t, s_tok, w_step=w_step, kv_row=row_b, mem_row=row_b) for t in range(1, MAX_LEN + 1))
        survey = sum(synth.decode_step_bytes(B, t, s_tok) for t in range(1, MAX_LEN + 1)) * (1.0 if absorbed else 2.0)
        sec_per_batch = ms / 1e3 / args.steps
        job = {"decode_bytes_moved_per_batch": moved, "decode_bytes_survey_per_batch": survey,
               "hbm_frac_bytes_moved": moved / sec_per_batch / 1e9 / peaks["hbm_gbs"],
               "hbm_frac_survey_effective": survey / sec_per_batch / 1e9 / peaks["hbm_gbs"],
               "note": "bytes_moved = what the decode loop of this formulation reads per batch (latent rows + per-step weights); "
                       "survey_effective divides SURVEY.md 8d's projected-K/V bytes by the same time: an algorithmic saving, not HBM utilisation. "
                       "Both use the whole step time, encoder included"}
        # secondary metric of BASELINE.json: encoder img/s -- configs[1] (256 mixed-width images, one ragged batch) and the headline's own batch
        def enc_rate(batch, n_img, flops):
            for _ in range(2):
                model.encoder(batch)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(3):
                model.encoder(batch)
            e1.record(stream)
            torch.cuda.synchronize()
            t_ms = e0.elapsed_time(e1) / 3
            tf = flops / (t_ms / 1e3) / 1e12
            return {"value": n_img / (t_ms / 1e3), "unit": "images/s", "ms": t_ms, "algorithmic_tflops": tf,
                    "frac_of_bf16_peak": tf / peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])}
        widths = synth.synth_widths(256, seed=77)
        rag = [synth.synth_images(1, H, w, seed=500 + i)[0].cuda() for i, w in enumerate(widths)]
        encoder = {"metric": "encoder img/s",
                   "config2_ragged": dict(enc_rate(rag, 256, sum(synth.encoder_flops(H, w) for w in widths)),
                                          workload="BASELINE configs[1]: 256 images, H=64, widths 128..1008 (multiples of 16), one ragged batch"),
                   "uniform": dict(enc_rate(img_dev, rows_per_batch, synth.encoder_flops(H, W) * rows_per_batch),
                                   workload=f"{rows_per_batch} images {H}x{W} (the headline batch)"),
                   "note": "frac_of_bf16_peak is on ALGORITHMIC FLOPs; the backbone issues 3 MMAs per product (bf16x3), ceiling 0.41 (SURVEY.md 8d)"}
        encoder["value"] = encoder["config2_ragged"]["value"]
        next_rows = measure_next_rows(model, img_dev, peaks, stream)

    # ---- what each kernel class costs a batch WITH the batches in flight (the eager per-launch figures above carry ~5 us of launch
    # latency each and say little about the pipelined path): the same pipelined run with classes of decode kernels switched off
    # (engine option dbg_skip: 1 self-attention, 2 cross-attention, 4 LayerNorm, 8 decode GEMMs; results are garbage, timing only)
    if rank == 0 and not args.no_extras and n_fly > 1 and world == 1 and roofline_attn is not None:
        from texocr_b200.pipeline import GeneratePipeline
        ab = {}
        with GeneratePipeline(model, in_flight=n_fly, branches=args.branches) as p2:
            p2.warm_up(img_dev, MAX_LEN)
            for name, skip in (("all", 0), ("no_attention", 3), ("no_gemm_ln", 12), ("encoder_and_token_kernels_only", 15)):
                for e_ in p2.engines:
                    e_.set_option("dbg_skip", skip)
                k = 2 * n_fly
                list(p2.generate_batches([img_dev] * n_fly, MAX_LEN))
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                list(p2.generate_batches([img_dev] * k, MAX_LEN))
                torch.cuda.synchronize()
                ab[name] = (time.perf_counter() - t0) * 1e3 / k
        attn_ms = ab["no_gemm_ln"] - ab["encoder_and_token_kernels_only"]
        a_total = sum(r["bytes"] for r in rows if r["name"] in ("dec_attn_self", "dec_attn_cross"))
        g_flops = sum(r["flops"] for r in rows if r["name"].startswith("dec_gemm"))
        chain_ms = ab["no_attention"] - ab["encoder_and_token_kernels_only"]
        if roofline is not None and (roofline.get("class", "").startswith("dec_gemm") or roofline.get("class") == "gemm64_store_bf16") and chain_ms > 0:
            roofline["in_flight"] = {
                "class": "all decode GEMMs (33 launches per step) + the LayerNorm launches between them", "ms_per_batch": chain_ms,
                "achieved": g_flops / (chain_ms / 1e3) / 1e12, "unit": "TFLOP/s",
                "frac": g_flops / (chain_ms / 1e3) / 1e12 / peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]),
                "method": f"{n_fly} batches in flight, wall clock per batch with the decode attention switched off minus the same with the decode "
                          "GEMMs and LayerNorms switched off as well; FLOPs = the algorithmic FLOPs of one batch's decode GEMM launches.  The eager "
                          "per-launch figure above carries the launch latency of a 4-5 us kernel measured alone"}
        roofline_attn["in_flight"] = {
            "ms_per_batch": ab, "attention_ms_per_batch": attn_ms, "gemm_ln_chain_ms_per_batch": ab["no_attention"] - ab["encoder_and_token_kernels_only"],
            "attention_gbs": a_total / (attn_ms / 1e3) / 1e9, "attention_frac_hbm": a_total / (attn_ms / 1e3) / 1e9 / peaks["hbm_gbs"],
            "method": f"{n_fly} batches in flight, wall clock per batch with decode kernel classes switched off; attention = (attention only) - (neither); "
                      "bytes = the algorithmic latent-row bytes of one batch's attention launches"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, per_step, kind, note = cpu_reference_eq_per_s(8)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                        "sample": f"B=8 synthetic {H}x{W} images (BASELINE config 1), full {MAX_LEN}-step greedy loop without KV cache, "
                                  f"fp32 torch CPU, {per_step:.1f} s; {note}"}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    workload = (f"BASELINE configs[4]: data-parallel sweep of {args.total} synthetic equations in batches of {B} (fixed batch boundaries, "
                f"contiguous runs of batches per rank), " if strong else "BASELINE configs[2]: ") + \
               (f"full greedy generate with KV cache, batch {B} per GPU, {H}x{W} images, max_len {MAX_LEN}, default config.yml model "
                f"(ResNetV2-hybrid ViT encoder + 4-layer decoder), random-init weights")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": workload,
                   "batch_per_gpu": B, "max_len": MAX_LEN, "image": [H, W], "precision": args.precision,
                   "parallelism": f"dp{world} (independent shards, token-id all_gather)",
                   "batches_in_flight": n_fly, "decode_branches_per_batch": args.branches if n_fly > 1 else max(1, min(8, (B + 64) // 128)),
                   "total_equations": done_eq if strong else None,
                   "l2_policy": "working set per batch in flight (0.25 GB latent attention cache + 0.2 GB encoder memory / decode buffers + >= 3 GB encoder activations) exceeds the 126 MB L2 many times over; no explicit flush"},
        "e2e": e2e, "one_batch_at_a_time": serial, "gpu_launches": int(launches), "clocks": clk, "roofline": roofline,
        "roofline_by_class": by_class, "roofline_attention": roofline_attn, "job_hbm": job, "kernel_time_shares": shares,
        "cpu_baseline": cpu_baseline, "encoder": encoder, "next_rows": next_rows, "verify": verify,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
