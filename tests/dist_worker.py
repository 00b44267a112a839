"""Worker for tests/test_dist_gloo.py: one rank of a world_size-2 gloo job on
CPU.  Uses the emulated engine (test infrastructure) to produce rows, shards
frames with pypanadapter_b200.dist and checks the gathered result on rank 0."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist
    from pypanadapter_b200 import _lib, dist as zdist, synth
    from pypanadapter_b200.engine import ZoomPSD
    from tests.emu import build_emu

    out_path = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = _lib.load_library(build_emu.OUT)
    w = synth.CFG1
    n = 2048 * 10
    frames = synth.make_frames(w, 5, n=n)               # 5 frames over 2 ranks: ragged blocks
    with ZoomPSD(0, lib=lib) as eng:
        eng.configure(w.fs, w.fft_size, w.fft_ratio, n, w.window, crop="thread")
        rows = zdist.process_frames_sharded(eng, frames)
        if rank == 0:
            single = eng.process(frames)
            ok = rows is not None and rows.shape == single.shape and np.array_equal(rows, single)
            with open(out_path, "w") as f:
                f.write("ok" if ok else "mismatch")
        else:
            assert rows is None
        # channel sharding (cfg4-style): 7 channels over 2 ranks
        mine = list(zdist.channels_for_rank(7, rank, world))
        assert mine == ([0, 1, 2, 3] if rank == 0 else [4, 5, 6])
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
