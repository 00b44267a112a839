#!/usr/bin/env python
"""BASELINE configs[2] at full size: batched offline waterfall over 1e10
synthetic complex64 samples, 65536-pt FFT, 50 % overlap Hann, rows of 2^20
samples (9536 rows).  Samples are generated ON THE DEVICE from a counter-based
integer hash of the sample index (splitmix64 -> two uniforms -> sum of 4 =
near-Gaussian noise) plus a tone, reproducible on the host in numpy; a
deterministic subset of rows is regenerated on the host and checked against
the oracle (0.01 dB20 above the floor, peak bin exact).

    python tests/tools/cfg3_job.py [--samples 1e10] [--batch 64] [--check 6]

torch is used only to fill device buffers; every row comes from the engine.
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

FS = 2.4e6
N = 65536
ROW = 1 << 20
TONE_HZ, TONE_A, SIGMA = 301234.5, 0.5, 1e-3
M64 = (1 << 64) - 1


def _mix_np(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def host_row(row_index: int) -> np.ndarray:
    """The same samples as device_rows(), in numpy."""
    with np.errstate(over="ignore"):
        k = np.arange(ROW, dtype=np.uint64) + np.uint64(row_index) * np.uint64(ROW)
        h = _mix_np(k)
        g = _mix_np(h)
        parts = [(h & np.uint64(0xFFFF)), ((h >> np.uint64(16)) & np.uint64(0xFFFF)),
                 ((h >> np.uint64(32)) & np.uint64(0xFFFF)), (h >> np.uint64(48)),
                 (g & np.uint64(0xFFFF)), ((g >> np.uint64(16)) & np.uint64(0xFFFF)),
                 ((g >> np.uint64(32)) & np.uint64(0xFFFF)), (g >> np.uint64(48))]
    u = [p.astype(np.float64) / 65536.0 for p in parts]
    # sum of 4 uniforms, centred, variance 4/12 -> scaled to unit variance
    nr = (u[0] + u[1] + u[2] + u[3] - 2.0) * np.sqrt(3.0)
    ni = (u[4] + u[5] + u[6] + u[7] - 2.0) * np.sqrt(3.0)
    ph = (k.astype(np.float64) * (TONE_HZ / FS)) % 1.0
    x = TONE_A * np.exp(2j * np.pi * ph) + SIGMA * (nr + 1j * ni)
    return x.astype(np.complex64)


def device_rows(torch, first_row: int, nrows: int):
    """(nrows, ROW) complex64 on the GPU, same arithmetic in torch."""
    dev = "cuda"
    k = (torch.arange(nrows * ROW, dtype=torch.int64, device=dev) + first_row * ROW)

    def mix(z):
        z = z + (-7046029254386353131)                                  # 0x9E3779B97F4A7C15 as int64
        z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * (-4658895280553007687)  # logical shift via mask
        z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * (-7723592293110705685)
        return z ^ ((z >> 31) & ((1 << 33) - 1))

    h = mix(k)
    g = mix(h)

    def parts(z):
        return [(z & 0xFFFF), ((z >> 16) & 0xFFFF), ((z >> 32) & 0xFFFF), ((z >> 48) & 0xFFFF)]

    u = [p.to(torch.float64) / 65536.0 for p in parts(h) + parts(g)]
    s3 = float(np.sqrt(3.0))
    nr = (u[0] + u[1] + u[2] + u[3] - 2.0) * s3
    ni = (u[4] + u[5] + u[6] + u[7] - 2.0) * s3
    ph = torch.remainder(k.to(torch.float64) * (TONE_HZ / FS), 1.0) * (2.0 * np.pi)
    re = TONE_A * torch.cos(ph) + SIGMA * nr
    im = TONE_A * torch.sin(ph) + SIGMA * ni
    x = torch.complex(re.to(torch.float32), im.to(torch.float32))
    return x.reshape(nrows, ROW)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=float, default=1e10)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--check", type=int, default=6, help="rows regenerated on the host and checked")
    args = ap.parse_args()

    import torch
    from oracle import zoompsd_oracle as zo
    from pypanadapter_b200.engine import ZoomPSD
    from tests import parity

    nrows = int(args.samples) // ROW
    eng = ZoomPSD(0)
    eng.configure(FS, N, 1, ROW, "hann", crop=None)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    check = sorted(set(int(v) for v in np.linspace(0, nrows - 1, args.check)))
    kept = {}
    d_rows = torch.empty((args.batch, N), dtype=torch.float32, device="cuda")
    t_engine = 0.0
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for first in range(0, nrows, args.batch):
        nb = min(args.batch, nrows - first)
        x = device_rows(torch, first, nb)
        xv = torch.view_as_real(x).contiguous()
        e0.record(stream)
        eng.process_device(xv.data_ptr(), nb, d_rows.data_ptr())
        e1.record(stream)
        torch.cuda.synchronize()
        t_engine += e0.elapsed_time(e1) * 1e-3
        for r in check:
            if first <= r < first + nb:
                kept[r] = d_rows[r - first].cpu().numpy().astype(np.float64)
        del x, xv
    wall = time.perf_counter() - t0
    gen_diff = float(np.abs(device_rows(torch, check[1], 1)[0].cpu().numpy() - host_row(check[1])).max())
    floor = parity.floor_db20(FS, "hann", N, False)
    worst = 0.0
    for r in check:
        want = zo.zoom_psd(host_row(r), FS, N, 1, "hann", crop=None)
        worst = max(worst, parity.assert_row_parity(kept[r], want, floor, "cfg3 row %d" % r))
    total = nrows * ROW
    print(json.dumps({
        "workload": "cfg3: %.3g complex64 samples, 65536-pt FFT, 50%% overlap Hann, rows of 2^20" % total,
        "rows": nrows, "engine_seconds": t_engine, "wall_seconds_incl_generation": wall,
        "gsamples_per_s_engine": total / t_engine / 1e9, "rows_per_s_engine": nrows / t_engine,
        "gbytes_per_s_algorithmic": total * 8 / t_engine / 1e9,
        "rows_checked_against_oracle": check, "max_abs_diff_db20": worst,
        "device_vs_host_generator_max_abs": gen_diff}))


if __name__ == "__main__":
    main()
