"""N > 1 path on CPU: world_size-2 gloo job (torch.distributed.run) sharding
frames across ranks and gathering rows to rank 0; plus the pure sharding math."""
import os
import socket
import subprocess
import sys

import pytest

from pypanadapter_b200 import dist as zdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,world", [(0, 1), (1, 4), (5, 2), (64, 8), (9536, 8), (7, 3)])
def test_shard_ranges_partition(n, world):
    blocks = [zdist.shard_range(n, r, world) for r in range(world)]
    assert blocks[0][0] == 0 and blocks[-1][1] == n
    for a, b in zip(blocks, blocks[1:]):
        assert a[1] == b[0]
    sizes = [hi - lo for lo, hi in blocks]
    assert max(sizes) - min(sizes) <= 1
    assert zdist.shard_counts(n, world) == sizes


def test_shard_range_rejects_bad_rank():
    with pytest.raises(ValueError):
        zdist.shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_gloo_gather(tmp_path, emu_lib):
    out = tmp_path / "result.txt"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_worker.py"), str(out)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert out.read_text() == "ok"


def test_cpulist_and_gpu_local_cpus(tmp_path):
    from pypanadapter_b200 import dist as zdist
    assert zdist.parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert zdist.parse_cpulist("") == set()
    dev = tmp_path / "bus" / "pci" / "devices" / "0000:1b:00.0"
    dev.mkdir(parents=True)
    (dev / "local_cpulist").write_text("0-1\n")
    assert zdist.gpu_local_cpus("0000:1B:00.0", str(tmp_path)) == {0, 1}
    assert zdist.gpu_local_cpus("00000000:1b:00.0", str(tmp_path)) == {0, 1}      # nvml's 8-digit domain
    assert zdist.gpu_local_cpus("0000:3c:00.0", str(tmp_path)) == set()


def test_bind_host_to_gpu_is_fail_safe(tmp_path):
    """Binding narrows the affinity to the GPU's local CPUs when sysfs names some
    inside the cpuset, and otherwise reports why not -- it never raises."""
    import os
    from pypanadapter_b200 import dist as zdist
    before = os.sched_getaffinity(0)
    try:
        r = zdist.bind_host_to_gpu("0000:1b:00.0", str(tmp_path))               # nothing in sysfs
        assert r["bound"] is False and os.sched_getaffinity(0) == before
        dev = tmp_path / "bus" / "pci" / "devices" / "0000:1b:00.0"
        dev.mkdir(parents=True)
        (dev / "local_cpulist").write_text("100000-100003\n")                     # outside the cpuset
        r = zdist.bind_host_to_gpu("0000:1b:00.0", str(tmp_path))
        assert r["bound"] is False and os.sched_getaffinity(0) == before
        (dev / "local_cpulist").write_text(",".join(str(c) for c in sorted(before)) + "\n")
        r = zdist.bind_host_to_gpu("0000:1b:00.0", str(tmp_path))               # whole cpuset: one domain
        assert r["bound"] is False and r["why"] == "single NUMA domain"
        if len(before) >= 2:
            keep = sorted(before)[:len(before) // 2]
            (dev / "local_cpulist").write_text(",".join(str(c) for c in keep) + "\n")
            r = zdist.bind_host_to_gpu("0000:1b:00.0", str(tmp_path))
            assert r == {"bound": True, "cpus": len(keep), "of": len(before)}
            assert os.sched_getaffinity(0) == set(keep)
    finally:
        os.sched_setaffinity(0, before)
