#!/usr/bin/env python
"""Waterfall display path on the device next to the reference's
host arithmetic: image assembly + level mapping (one streaming kernel, reads
4 B and writes 1 B per pixel) and the autolevel percentiles (three radix
histogram sweeps, 4 B per pixel and sweep), per image size.

    python tests/tools/image_bench.py [--widths 1024,8192,32768] [--reps 20]

Device times come from CUDA events on the engine's stream (zfb_set_profiling);
the HBM peak is MEASURED_PEAKS.json's.  The CPU column times the oracle's
restatement of what the reference + pyqtgraph do per displayed row: np.roll of
the whole image (S:1652), the level mapping, np.percentile (S:1676).
torch is used only to own the device output buffer.
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


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--widths", default="1024,8192,32768")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    import torch
    from oracle import zoompsd_oracle as zo
    from pypanadapter_b200.engine import ZoomPSD

    peak, peak_src = hbm_peak()
    eng = ZoomPSD(0)
    rng = np.random.default_rng(7)
    for w in [int(v) for v in a.widths.split(",")]:
        h = w // 4
        # any configuration whose row width is w: R = 1, crop = w
        eng.configure(2.4e6, w, 1, w * 2, "hann", crop=w)
        eng.ring_configure(h)
        rows = (-150.0 + 10.0 * rng.standard_normal((h, w))).astype(np.float32)
        for i in range(0, h, 256):
            eng.push_rows(rows[i:i + 256])
        seen = h + 5
        d_out = torch.empty(h * w, dtype=torch.uint8, device="cuda")
        eng.set_stream(torch.cuda.current_stream().cuda_stream)
        for _ in range(3):
            eng.ring_image_device(d_out.data_ptr(), h, 1, seen, "u8")
        eng.ring_quantiles(h, 1, seen, [0.02, 0.98])
        torch.cuda.synchronize()
        eng.set_profiling(True)
        eng.profile()
        flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
        for _ in range(a.reps):
            flush.sum()                                     # images smaller than L2: evict between reps (clean lines)
            eng.ring_image_device(d_out.data_ptr(), h, 1, seen, "u8")
        torch.cuda.synchronize()
        prof = eng.profile()
        img_ms = prof["waterfall_image"][0] / prof["waterfall_image"][1]
        for _ in range(max(2, a.reps // 4)):
            flush.sum()
            q, n = eng.ring_quantiles(h, 1, seen, [0.02, 0.98])
        prof = eng.profile()
        sel_ms = prof["autolevel_select"][0] / prof["autolevel_select"][1]
        eng.set_profiling(False)
        t0 = time.perf_counter()
        idx = eng.ring_image(h, 1, seen, "u8", levels=(-220, -120))
        e2e_img = time.perf_counter() - t0
        t0 = time.perf_counter()
        eng.ring_quantiles(h, 1, seen, [0.02, 0.98])
        e2e_sel = time.perf_counter() - t0
        px = h * w
        line = {
            "width": w, "height": h, "pixels": px,
            "image_u8": {"ms": img_ms, "algorithmic_bytes": 5 * px, "gbs": 5 * px / img_ms / 1e6,
                         "frac_hbm": 5 * px / img_ms / 1e6 / peak, "host_call_ms_incl_d2h": e2e_img * 1e3},
            "autolevel": {"ms_per_sweep": sel_ms, "sweeps": 3, "algorithmic_bytes_per_sweep": 4 * px,
                          "gbs": 4 * px / sel_ms / 1e6, "frac_hbm": 4 * px / sel_ms / 1e6 / peak,
                          "host_call_ms": e2e_sel * 1e3, "levels": [float(q[0]), float(q[1])], "count": n},
            "hbm_peak_gbs": peak, "peak_source": peak_src,
        }
        if not a.no_cpu and px <= (1 << 26):
            img = eng.ring_image(h, 1, seen, "f32").astype(np.float64)
            t0 = time.perf_counter()
            img2 = np.roll(img, -1, 0)
            t_roll = time.perf_counter() - t0
            t0 = time.perf_counter()
            ref_idx = zo.waterfall_indices(img2, -220, -120)
            t_map = time.perf_counter() - t0
            t0 = time.perf_counter()
            ref_q = zo.waterfall_autolevel(img)
            t_pct = time.perf_counter() - t0
            line["cpu"] = {"roll_ms": t_roll * 1e3, "level_map_ms": t_map * 1e3, "percentile_ms": t_pct * 1e3,
                           "indices_equal": bool(np.array_equal(idx, zo.waterfall_indices(img, -220, -120))),
                           "levels_equal": bool(ref_q == (float(q[0]), float(q[1])))}
            del ref_idx
        print(json.dumps(line), flush=True)
        del d_out, flush


if __name__ == "__main__":
    main()
