"""GPU tier, two GPUs (skipped on a one-GPU box): the data-parallel path of bench.py -- one process per GPU over NCCL -- returns
exactly the tokens one GPU returns for the same images (pure data parallelism, bitwise equal)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_bench_verifies_against_single_gpu_decode():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29621", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "6", "--warmup", "3", "--verify",
           "--no-cpu-baseline", "--no-extras", "--total", str(512 * 12)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert lines, out.stdout[-2000:] + out.stderr[-2000:]
    rec = json.loads(lines[-1])
    assert rec["n_gpus"] == 2 and rec["scaling"] == "strong" and rec["verify"]["ok"] is True
    assert rec["config"]["total_equations"] == 512 * 12
