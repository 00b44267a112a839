#!/usr/bin/env python
"""Single-frame latency of the drop-in surface (the live use case: 10 rows/s
of ~0.24 M samples): wall time of one call, host to host.

    python tests/tools/latency.py
"""
from __future__ import annotations

import json
import os
import sys
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def stats(ts):
    ts = np.array(ts) * 1e6
    return {"median_us": float(np.median(ts)), "p90_us": float(np.percentile(ts, 90)), "min_us": float(ts.min())}


def main():
    from oracle import zoompsd_oracle as zo
    from pypanadapter_b200 import synth
    from pypanadapter_b200.buffers import PSD, Data
    from pypanadapter_b200.engine import ZoomPSD, zoom_psd

    out = {}
    eng = ZoomPSD(0)
    for w in (synth.CFG1, synth.CFG2):
        x = synth.make_frame(w, 0)
        for mode in ("fast", "exact"):
            kw = dict(flip=w.flip, crop="thread", mode=mode, engine=eng)
            for _ in range(5):
                zoom_psd(x, w.fs, w.fft_size, w.fft_ratio, w.window, **kw)
            ts = []
            for _ in range(200):
                t0 = time.perf_counter()
                zoom_psd(x, w.fs, w.fft_size, w.fft_ratio, w.window, **kw)
                ts.append(time.perf_counter() - t0)
            out["zoom_psd %s %s" % (w.name, mode)] = stats(ts)
        t0 = time.perf_counter()
        zo.zoom_psd(x, w.fs, w.fft_size, w.fft_ratio, w.window, flip=w.flip)
        out["cpu oracle %s (1 core)" % w.name] = {"median_us": (time.perf_counter() - t0) * 1e6}
    # Data ring + PSD.update: samples are already on the device when update() runs
    w = synth.CFG1
    x = synth.make_frame(w, 0)
    state = types.SimpleNamespace(fft_size=w.fft_size, fft_ratio=w.fft_ratio, fft_tapering=w.window,
                                  panadapter=types.SimpleNamespace(SampleRate=w.fs))
    d = Data(engine=eng).new_complex()
    psd = PSD(d, state)
    t_add, t_upd = [], []
    for it in range(60):
        for i in range(0, len(x) - d.chunk_size + 1, d.chunk_size):
            t0 = time.perf_counter()
            d.add(x[i:i + d.chunk_size])
            t_add.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        psd.update()
        if it >= 5:
            t_upd.append(time.perf_counter() - t0)
    out["Data.add (16392 samples, pinned copy + async H2D)"] = stats(t_add[50:])
    out["PSD.update cfg1 (ring resident)"] = stats(t_upd)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
