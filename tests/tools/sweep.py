#!/usr/bin/env python
"""BASELINE configs[4]: decimation / FFT-size sweep (R 1..64 x N 1024..262144,
avg = max(R, 16), complex64, Hamming) -- device-resident throughput of the
engine next to the CPU oracle port (one core, one frame) on the same box.

    python tests/tools/sweep.py [--out profiles/sweep.jsonl] [--budget-mb 256] [--no-cpu]

Prints one JSON line per grid point and a markdown table at the end.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--budget-mb", type=float, default=256.0, help="input bytes per timed launch")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--ratios", default="1,2,4,8,16,32,64")
    ap.add_argument("--sizes", default="1024,4096,16384,65536,262144")
    ap.add_argument("--mode", default="fast")
    ap.add_argument("--set", action="append", dest="sets", metavar="OPTION=VALUE", help="zfb_set_option before the sweep")
    ap.add_argument("--lib", default=None, help="an alternative sm_100a build of the same sources (compile-time A/B)")
    args = ap.parse_args()

    import torch
    from oracle import golden_cases as gc
    from oracle import zoompsd_oracle as zo
    from pypanadapter_b200.engine import ZoomPSD

    fs = 2.4e6
    if args.lib:
        from pypanadapter_b200 import _lib
        eng = ZoomPSD(0, lib=_lib.load_library(args.lib))
    else:
        eng = ZoomPSD(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    for it in args.sets or []:
        name, _, val = it.partition("=")
        eng.set_option(name, int(val))
    results = []
    for R in [int(v) for v in args.ratios.split(",")]:
        for N in [int(v) for v in args.sizes.split(",")]:
            avg = max(R, 16)
            n = N * avg
            frame_bytes = n * 8
            frames = int(max(1, min(4096, args.budget_mb * 1e6 // frame_bytes)))
            x = gc.tone_noise(n, fs, [(0.013 * fs / R, 0.4), (-0.02 * fs / R, 0.03)], 2e-3, 7, np.complex64)
            eng.configure(fs, N, R, n, "hamming", crop="thread", mode=args.mode)
            host = np.ascontiguousarray(np.broadcast_to(x, (frames, n)))
            d_in = torch.from_numpy(host.view(np.float32).reshape(frames, -1)).cuda()
            d_rows = torch.empty((frames, eng.row_width), dtype=torch.float32, device="cuda")
            for _ in range(3):
                eng.process_device(d_in.data_ptr(), frames, d_rows.data_ptr())
            torch.cuda.synchronize()
            reps = 5
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                eng.process_device(d_in.data_ptr(), frames, d_rows.data_ptr())
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            gpu = frames * n / (ms * 1e-3) / 1e6
            row0 = d_rows[0].cpu().numpy().astype(np.float64)
            rec = dict(R=R, N=N, frame_len=n, frames=frames, mode="fast" if eng.fast_active else "exact",
                       gpu_msamples_s=gpu, gpu_ms=ms, gbytes_s=gpu * 8e-3)
            if not args.no_cpu:
                t0 = time.perf_counter()
                want = zo.zoom_psd(x, fs, N, R, "hamming", crop="thread")
                dt = time.perf_counter() - t0
                rec["cpu_msamples_s_1core"] = n / dt / 1e6
                m = want > want.max() - 160
                rec["max_abs_diff_db20"] = float(np.abs(row0 - want)[m].max())
                rec["argmax_equal"] = bool(int(np.argmax(row0)) == int(np.argmax(want)))
            results.append(rec)
            print(json.dumps(rec), flush=True)
            del d_in, d_rows
    if args.out:
        with open(args.out, "w") as f:
            for r in results:
                f.write(json.dumps(r) + "\n")
    sizes = sorted({r["N"] for r in results})
    print("\n| R \\\\ N | " + " | ".join(str(s) for s in sizes) + " |")
    print("|---|" + "---|" * len(sizes))
    for R in sorted({r["R"] for r in results}):
        cells = []
        for N in sizes:
            r = next((q for q in results if q["R"] == R and q["N"] == N), None)
            cells.append("—" if r is None else "%.1f" % (r["gpu_msamples_s"] / 1e3))
        print("| %d | " % R + " | ".join(cells) + " |")
    print("(GPU Gsamples/s, device-resident, complex64)")


if __name__ == "__main__":
    main()
