"""bench.py's reference arm (the CPU leg the driver runs as `--impl reference`):
one JSON line with the contract's keys on rank 0, nothing and exit 0 on the
other ranks; the GPU arm refuses to run without a CUDA device (no CPU fallback)."""
from __future__ import annotations

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, **env):
    e = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        e.pop(k, None)
    e.update({k: str(v) for k, v in env.items()})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=e,
                          capture_output=True, text=True, timeout=600)


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "3"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["metric"].startswith("input Msamples/s") and line["unit"] == "Msamples/s"
    assert line["steps"] == 1 and line["warmup"] >= 3 and line["n_gpus"] == 1
    assert line["value"] > 0 and line["ms_per_step"] > 0
    assert line["config"]["workload"].startswith("cfg2") and "model" not in line["config"]
    assert line["config"]["sample_dtype"] == "u8" and line["config"]["fft_size"] == 4096
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"],
                           "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_reference_arm_other_ranks_do_nothing():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "3"],
             RANK=1, LOCAL_RANK=1, WORLD_SIZE=2)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = _run(["--steps", "1", "--warmup", "3", "--no-cpu-baseline"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stdout + r.stderr)
