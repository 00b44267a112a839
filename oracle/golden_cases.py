"""Golden-vector case list -- TEST INFRASTRUCTURE ONLY.

Shared by oracle/make_golden.py (which runs the REFERENCE's own methods on
these inputs, in the build container) and tests/ (which regenerate the same
inputs from their seeds and compare oracle / CUDA outputs with the recorded
rows).  Inputs are never stored, only parameters + reference outputs.
"""
from __future__ import annotations

import numpy as np

from pypanadapter_b200 import synth


def tone_noise(n, fs, tones, sigma, seed, dtype=np.complex128):
    k = np.arange(n, dtype=np.float64)
    x = np.zeros(n, dtype=np.complex128)
    for f, a in tones:
        x += a * np.exp(2j * np.pi * (f / fs) * k)
    rng = np.random.default_rng(seed)
    x += sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return x.astype(dtype)


def real_tone_noise(n, fs, tones, sigma, seed, dtype=np.float64):
    """Real samples, as AudioPan's callback delivers them (S:712-714)."""
    k = np.arange(n, dtype=np.float64)
    x = np.zeros(n, dtype=np.float64)
    for f, a in tones:
        x += a * np.cos(2 * np.pi * (f / fs) * k + 0.3)
    x += sigma * np.random.default_rng(seed).standard_normal(n)
    return x.astype(dtype)


def ref_sine(size, cycles=10.0):
    """The reference's own Waveform.Sine test signal (S:992-994)."""
    a = np.linspace(0, 2 * np.pi * cycles, size)
    return np.sin(a) + 1j * np.cos(a)


def ref_random(size, seed):
    """Same distribution as the reference's Waveform.Random (S:1001-1002):
    uniform complex in [-1,1)^2 (seeded here; the reference is unseeded)."""
    rng = np.random.default_rng(seed)
    return (2 * rng.random(size) - 1) + 1j * (2 * rng.random(size) - 1)


def make_input(case: dict) -> np.ndarray:
    """Chunk exactly as handed to the reference method (complex128/complex64;
    u8 cases are converted + flipped by the oracle's K0/K1 before the
    reference sees them, see make_golden.py)."""
    kind = case["input"]
    if kind == "workload":
        w = synth.WORKLOADS[case["workload"]]
        return synth.make_frame(w, case.get("frame", 0), case.get("n"))
    if kind == "tone_noise":
        return tone_noise(case["n"], case["fs"], case["tones"], case["sigma"],
                          case["seed"],
                          np.complex64 if case.get("c64") else np.complex128)
    if kind == "real_tone_noise":
        return real_tone_noise(case["n"], case["fs"], case["tones"], case["sigma"], case["seed"],
                               np.float32 if case.get("f32") else np.float64)
    if kind == "ref_sine":
        return ref_sine(case["n"], case.get("cycles", 10.0))
    if kind == "ref_random":
        return ref_random(case["n"], case["seed"])
    if kind == "dc":
        return np.full(case["n"], 0.25 + 0.125j, dtype=np.complex128)
    if kind == "impulse":
        x = np.zeros(case["n"], dtype=np.complex128)
        x[case["at"]] = 1.0
        return x
    raise ValueError(kind)


# path: 'S' = ApplicationDisplay.update (crop = n_win), 'T' = PSD.update
CASES = [
    # --- BASELINE configs[0] (cfg1), both call paths -----------------------
    dict(name="cfg1_T", path="T", input="workload", workload="cfg1",
         fs=2.4e6, N=2048, R=8, window="hamming"),
    dict(name="cfg1_S1024", path="S", n_win=1024, input="workload",
         workload="cfg1", fs=2.4e6, N=2048, R=8, window="hamming"),
    dict(name="cfg1_S256", path="S", n_win=256, input="workload",
         workload="cfg1", fs=2.4e6, N=2048, R=8, window="hamming"),
    # --- BASELINE configs[1] (cfg2): u8 + flip, 4 frames (EMA checked on top)
    *[dict(name="cfg2_T_f%d" % i, path="T", input="workload", workload="cfg2",
           frame=i, u8=True, flip=True, fs=3.2e6, N=4096, R=16,
           window="hamming") for i in range(4)],
    # --- reference defaults (S:1492-1497): R=2, N=2048, hamming ------------
    dict(name="defaults_S", path="S", n_win=1024, input="tone_noise",
         n=2048 * 32, fs=2.56e6, tones=[(123456.0, 0.3), (-400e3, 0.01)],
         sigma=1e-3, seed=11, N=2048, R=2, window="hamming"),
    # --- no zoom (R=1): pure Welch ------------------------------------------
    dict(name="r1_hann_T", path="T", input="tone_noise", n=2048 * 16,
         fs=2.4e6, tones=[(300e3, 0.5)], sigma=1e-4, seed=12, N=2048, R=1,
         window="hann"),
    dict(name="r1_c64_S", path="S", n_win=1024, input="tone_noise", c64=True,
         n=1024 * 16, fs=2.4e6, tones=[(-250e3, 0.7)], sigma=1e-3, seed=13,
         N=1024, R=1, window="hamming"),
    # --- every zoom depth the UI offers on a short chunk --------------------
    *[dict(name="zoom_R%d" % r, path="T", input="tone_noise", n=1024 * 4 * r,
           fs=2.4e6, tones=[(0.11 * 2.4e6 / r / r, 0.4), (-900.0, 0.02)],
           sigma=2e-3, seed=20 + r, N=1024, R=r, window="hamming")
      for r in (2, 4, 8, 16, 32, 64)],
    # --- windows incl. tuple forms (S:1222-1243, S:1354-1363) ---------------
    *[dict(name="win_%s" % (w if isinstance(w, str) else w[0].replace(" ", "_")),
           path="T", input="tone_noise", n=512 * 24, fs=1e6,
           tones=[(7000.0, 0.5)], sigma=1e-3, seed=40, N=512, R=4, window=w)
      for w in ("boxcar", "blackmanharris", "flattop", "bartlett",
                ("kaiser", 14.0), ("gaussian", 7.0), ("tukey", 0.3),
                ("chebwin", 100.0), ("general gaussian", 1.5, 7.0))],
    # --- ragged lengths: T path takes whatever is in Data (T:1516-1518) ----
    dict(name="ragged_T", path="T", input="tone_noise", n=100003, fs=2.4e6,
         tones=[(3000.0, 0.5)], sigma=1e-3, seed=50, N=2048, R=8,
         window="hamming"),
    dict(name="ragged_odd_T", path="T", input="tone_noise", n=33333, fs=2.4e6,
         tones=[(-2000.0, 0.5)], sigma=1e-3, seed=51, N=1024, R=4,
         window="hamming"),
    # --- R > avg: decimated chunk shorter than N (welch zero-pads) ----------
    dict(name="short_T", path="T", input="tone_noise", n=2048 * 4, fs=2.4e6,
         tones=[(1000.0, 0.5)], sigma=1e-3, seed=52, N=2048, R=8,
         window="hamming"),
    # --- exactly one segment -------------------------------------------------
    dict(name="oneseg_T", path="T", input="tone_noise", n=1024 * 4, fs=2.4e6,
         tones=[(5000.0, 0.5)], sigma=1e-3, seed=53, N=1024, R=4,
         window="hamming"),
    # --- reference-native generators, DC and impulse ------------------------
    dict(name="refsine_T", path="T", input="ref_sine", n=2048 * 24, fs=2.4e6,
         N=2048, R=4, window="hamming"),
    dict(name="refrandom_S", path="S", n_win=512, input="ref_random",
         n=1024 * 16, seed=60, fs=2.4e6, N=1024, R=2, window="hamming"),
    dict(name="dc_R1_T", path="T", input="dc", n=1024 * 8, fs=2.4e6, N=1024,
         R=1, window="hamming"),
    dict(name="impulse_T", path="T", input="impulse", n=1024 * 16, at=5000,
         fs=2.4e6, N=1024, R=4, window="hamming"),
    # --- big FFTs ------------------------------------------------------------
    dict(name="n8192_T", path="T", input="tone_noise", n=8192 * 16, fs=20e6,
         tones=[(7000.0, 0.1)], sigma=1e-3, seed=70, N=8192, R=4,
         window="hamming"),
    dict(name="n16384_T", path="T", input="tone_noise", n=16384 * 6, fs=2.4e6,
         tones=[(17000.0, 0.1)], sigma=1e-3, seed=71, N=16384, R=2,
         window="hann"),
    dict(name="n65536_R1_T", path="T", input="tone_noise", n=65536 * 4,
         fs=2.4e6, tones=[(301234.5, 0.5)], sigma=1e-3, seed=72, N=65536,
         R=1, window="hann"),
    dict(name="n32_T", path="T", input="tone_noise", n=32 * 64, fs=48e3,
         tones=[(700.0, 0.5)], sigma=1e-3, seed=73, N=32, R=2,
         window="hamming"),
    # --- real samples (AudioPan, S:712-714; Data.new_real T:1413-1417): without
    #     zoom welch is one-sided and the reference fftshifts/crops N/2+1 bins
    #     (S path; T path only with Data.new_real -- its own set-up stores
    #     everything in the complex buffer, T:1807, hence two-sided rows)
    dict(name="audio_R1_T", path="T", input="real_tone_noise", n=2048 * 12,
         fs=48e3, tones=[(5000.0, 0.5), (11250.0, 0.02)], sigma=1e-3, seed=90,
         N=2048, R=1, window="hamming"),
    dict(name="audio_R1_Treal", path="T", data_real=True, input="real_tone_noise",
         n=2048 * 12, fs=48e3, tones=[(5000.0, 0.5), (11250.0, 0.02)], sigma=1e-3,
         seed=90, N=2048, R=1, window="hamming"),
    dict(name="audio_R1_S", path="S", n_win=1024, input="real_tone_noise",
         f32=True, n=2048 * 12, fs=48e3, tones=[(3000.0, 0.5)], sigma=1e-3,
         seed=91, N=2048, R=1, window="hann"),
    dict(name="audio_R4_T", path="T", input="real_tone_noise", n=1024 * 24,
         fs=48e3, tones=[(900.0, 0.4)], sigma=1e-3, seed=92, N=1024, R=4,
         window="hamming"),
]

# zoomfft (mix + decimate cascade) outputs: S:2088-2100
ZOOMFFT_CASES = [
    dict(name="zoomfft_R8", input="tone_noise", n=2048 * 8, fs=2.4e6,
         tones=[(5300.0, 0.5), (400e3, 0.3)], sigma=3e-3, seed=80, N=2048,
         R=8),
    dict(name="zoomfft_R2_short", input="tone_noise", n=64, fs=1e6,
         tones=[(1000.0, 0.5)], sigma=1e-2, seed=81, N=32, R=2),
]


def case_by_name(name):
    for c in CASES + ZOOMFFT_CASES:
        if c["name"] == name:
            return c
    raise KeyError(name)


def waterfall_rows_ramp():
    """70 rows of width 256 (wraps the 64-row image): row i = -100 - i + 0.01 x."""
    return [np.full(256, -100.0 - i) + np.arange(256) * 0.01 for i in range(70)]


def waterfall_rows_noise(n=40, width=256, seed=20260202):
    """n < 64 rows of float32-representable dB values around -150 (the -500
    fill is still part of the image): device rows equal the reference's."""
    rng = np.random.default_rng(seed)
    return [(-150.0 + 10.0 * rng.standard_normal(width)).astype(np.float32).astype(np.float64)
            for _ in range(n)]

