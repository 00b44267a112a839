"""Multi-GPU correctness on hardware: frames are independent -- LO
phase, filter state and Welch mean restart per chunk (pypanadapter_spectrum.py:
2092, 2098, 2111) -- so the rows N ranks compute for their shards and gather to
rank 0 over NCCL must EQUAL, bit for bit, the rows one GPU computes for the same
frames.  bench.py performs that check after its timed regions whenever N > 1
(``multi_gpu_check``); this test launches it the way the driver launches the
scaling runs.  Needs >= 2 GPUs on the box (``gpurun --gpus 2``); skipped otherwise."""
from __future__ import annotations

import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus() -> int:
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.parametrize("workload,mode", [("cfg2", "fast"), ("cfg1", "exact")])
def test_gathered_rows_equal_single_gpu_rows(workload, mode):
    n = _gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "bench.py"),
           "--gpus", str(world), "--steps", "3", "--warmup", "3", "--frames", "64", "--sustain-s", "0",
           "--no-cpu-baseline", "--workload", workload, "--mode", mode]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    line = json.loads(lines[0])
    assert line["n_gpus"] == world
    chk = line["multi_gpu_check"]
    assert chk["rows_equal_single_gpu"] is True and chk["max_abs_diff"] == 0.0, chk
    assert chk["frames"] == world * 48
