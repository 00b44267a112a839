"""Does the cfg2 step run faster when the batch is cut into slabs that go through the chain
concurrently (kernels of different slabs overlap: FIR chain of one beside the last stage /
strips / Welch of another, intermediates of a slab stay in L2)?  Measured with NE independent
engines on NE streams, each given F/NE frames per step -- a probe for slab pipelining inside
the engine, not a product path.

    python -m tests.tools.concurrency_probe [--workload cfg2] [--frames 512]
"""
from __future__ import annotations

import argparse
import json

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--frames", type=int, default=512)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--engines", default="1,2,4")
    ap.add_argument("--priority", type=int, default=0, help="1: alternate stream priorities")
    ap.add_argument("--set", action="append", dest="sets", metavar="OPTION=VALUE", help="zfb_set_option on every engine")
    args = ap.parse_args()

    import torch
    from pypanadapter_b200 import synth
    from pypanadapter_b200.engine import ZoomPSD

    w = synth.WORKLOADS[args.workload]
    F = args.frames
    host = synth.make_frames(w, F, distinct=min(F, 8))
    d_in = torch.from_numpy(host.view(np.uint8).reshape(F, -1)).cuda()
    row_bytes = d_in.shape[1]
    for ne in [int(x) for x in args.engines.split(",")]:
        per = F // ne
        engs, streams, rows = [], [], []
        for i in range(ne):
            e = ZoomPSD(0)
            lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
            st = torch.cuda.Stream(priority=(-1 if (args.priority and i % 2) else 0))
            e.set_stream(st.cuda_stream)
            for it in args.sets or []:
                name, _, val = it.partition("=")
                e.set_option(name, int(val))
            e.configure(w.fs, w.fft_size, w.fft_ratio, w.frame_len, w.window, dtype=w.dtype, flip=w.flip,
                        f_demod=w.f_demod, crop=w.crop, ema_alpha=w.ema_alpha, mode="fast")
            engs.append(e)
            streams.append(st)
            rows.append(torch.empty((per, e.row_width), dtype=torch.float32, device="cuda"))

        def step():
            for i, e in enumerate(engs):
                e.process_device(d_in.data_ptr() + i * per * row_bytes, per, rows[i].data_ptr())

        for _ in range(5):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main_stream = torch.cuda.current_stream()
        e0.record(main_stream)
        for st in streams:
            st.wait_event(e0)
        for _ in range(args.steps):
            step()
        for st in streams:
            ev = torch.cuda.Event()
            ev.record(st)
            main_stream.wait_event(ev)
        e1.record(main_stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        print(json.dumps({"engines": ne, "frames_per_engine": per, "ms_per_step": ms,
                          "gsamples_per_s": F * w.frame_len / ms / 1e6, "priority": args.priority,
                          "options": args.sets or []}), flush=True)
        for e in engs:
            e.close()


if __name__ == "__main__":
    main()
