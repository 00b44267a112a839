#!/usr/bin/env python
"""Wide randomised parity sweep against the oracle, beyond tests/engine_suite.random_configs:
N = 32 .. 16384 (four-step FFT included), R = 1 .. 64, 1 .. 6 Welch segments (single-segment rows
included), uint8 / complex64 / int16 wire formats, flip, ten windows, every crop kind, arbitrary
software-LO frequency, both decimator modes, batches of 1 .. 3 frames.

    python tests/tools/wide_sweep.py FIRST_SEED LAST_SEED [SECONDS] [--emu]

``--emu``: the CPU emulation build of the same kernel sources (tests/emu; logic check in the
GPU-less container) instead of the product library on cuda:0.  One line per failing seed.
The floor is the strict one (every bin above -100 dBFS); rows of up to six segments take the
engine's fp64 path (zfb_precise.cuh).  ``--lift-floor`` restores round 1's relaxed floor (bins
within 85 dB of the row's peak), ``--more-segments`` draws 1 .. 12 segments instead of 1 .. 6.
"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pypanadapter_b200 import _lib, synth
from pypanadapter_b200.engine import ZoomPSD
from tests import parity
from oracle import zoompsd_oracle as zo
argv = [a for a in sys.argv[1:] if not a.startswith("--")]
LIFT = "--lift-floor" in sys.argv        # round-1 behaviour: only bins within 85 dB of the row's peak
MAXSEG = 13 if "--more-segments" in sys.argv else 7
if "--emu" in sys.argv:
    from tests.emu import build_emu
    eng = ZoomPSD(0, lib=_lib.load_library(build_emu.build()))
else:
    eng = ZoomPSD(0)
lo, hi = int(argv[0]), int(argv[1])
budget = float(argv[2]) if len(argv) > 2 else 1200
windows = ["hamming", "hann", "boxcar", "blackmanharris", ("kaiser", 8.6), ("tukey", 0.5), "flattop", "bartlett",
           ("gaussian", 300.0), "nuttall"]
fails = 0; done = 0
t0 = time.time()
for seed in range(lo, hi):
    rng = np.random.default_rng(seed)
    N = int(2 ** rng.integers(5, 15))
    R = int(2 ** rng.integers(0, 7))
    segs = int(rng.integers(1, MAXSEG))
    n = int(N * R * (segs + 1) // 2 + rng.integers(0, 2 * R + 3))
    if n > 3_000_000:
        continue
    if R > 1:
        n = max(n, 28 * R + 5)
    dtype = ["u8", "c64", "cs16"][int(rng.integers(3))]
    flip = bool(rng.random() < 0.4)
    window = windows[int(rng.integers(len(windows)))]
    crop = [None, "thread", int(2 * rng.integers(1, N // 2 + 1))][int(rng.integers(3))]
    if crop == "thread" and N < 2 * R:
        crop = None
    f_demod = 1.0 if rng.random() < 0.4 else float(rng.uniform(-0.5, 0.5) * 1e6)
    mode = "fast" if rng.random() < 0.6 else "exact"
    nframes = int(rng.integers(1, 4))
    fs = 1e6
    k = np.arange(n)
    rows_in = []
    for f in range(nframes):
        ftone = f_demod + rng.uniform(-0.3, 0.3) * fs / R / max(1, R if crop == "thread" else 1)
        x = 0.45 * np.exp(2j * np.pi * ftone / fs * k) + 0.2 * np.exp(2j * np.pi * rng.uniform(-0.5, 0.5) * k)
        x = x + 5e-3 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
        rows_in.append(synth.quantise_u8(x * 0.9) if dtype == "u8" else synth.quantise_cs16(x * 0.9) if dtype == "cs16"
                       else x.astype(np.complex64))
    wire = np.stack(rows_in)
    what = "seed %d: N=%d R=%d n=%d %s flip=%s win=%s crop=%s f=%g %s frames=%d" % (
        seed, N, R, n, dtype, flip, window, crop, f_demod, mode, nframes)
    try:
        eng.configure(fs, N, R, n, window, dtype=dtype, flip=flip, f_demod=f_demod, crop=crop, mode=mode)
        rows = eng.process(wire).astype(np.float64)
        for f in range(nframes):
            want = zo.zoom_psd(wire[f], fs, N, R, window, f_demod=f_demod, crop=crop, flip=flip)
            floor = parity.floor_db20(fs, window, eng.geometry["nperseg"], R > 1)
            if LIFT:
                floor = max(floor, want.max() - 170.0)
            parity.assert_row_parity(rows[f], want, floor, what)
        done += 1
    except AssertionError as exc:
        fails += 1
        print("FAIL", what, "::", str(exc)[:300], flush=True)
    except Exception as exc:
        fails += 1
        print("ERROR", what, "::", repr(exc)[:300], flush=True)
    if time.time() - t0 > budget:
        print("stopped at seed", seed, flush=True); break
print("wide sweep seeds %d..%d: %d ok, %d failing, %.0f s" % (lo, seed, done, fails, time.time() - t0), flush=True)
