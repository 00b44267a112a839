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
