"""Synthetic IQ workloads shared by tests and bench.py.

Pure numpy host code: deterministic tone + noise frames for the BASELINE.json
configs.  Nothing here touches the GPU or the oracle.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Workload:
    name: str
    fs: float
    fft_size: int
    fft_ratio: int
    fft_avg: int
    window: object = "hamming"
    dtype: str = "c64"            # "c64" | "u8"
    flip: bool = False
    f_demod: float = 1.0
    crop: object = "thread"
    ema_alpha: float | None = None
    tones: tuple = ()             # ((freq_hz, amplitude), ...)
    sigma: float = 3e-3
    full_scale: float = 1.0       # peak scaling applied before quantisation
    seed: int = 20260101
    description: str = ""

    @property
    def frame_len(self) -> int:
        return self.fft_size * self.fft_avg

    @property
    def bytes_per_sample(self) -> int:
        return {"u8": 2, "cs16": 4}.get(self.dtype, 8)

    @property
    def row_width(self) -> int:
        if self.crop is None:
            return self.fft_size
        if self.crop == "thread":
            return 2 * int(.5 * self.fft_size / self.fft_ratio)
        return 2 * (int(self.crop) // 2)


# BASELINE.json configs[0]: the reference's own CPU-runnable case.
CFG1 = Workload(
    name="cfg1", fs=2.4e6, fft_size=2048, fft_ratio=8, fft_avg=int(2.4e6 / 2048 / 10),
    window="hamming", dtype="c64", tones=((5300.0, 0.5), (-11100.0, 0.05)),
    description="synthetic IQ tone+noise 2.4 MS/s, decim 8, 2048-pt FFT, complex64")

# BASELINE.json configs[1]: RTL-SDR replay, uint8 IQ, flip, EMA 0.3.
CFG2 = Workload(
    name="cfg2", fs=3.2e6, fft_size=4096, fft_ratio=16, fft_avg=int(3.2e6 / 4096 / 10),
    window="hamming", dtype="u8", flip=True, ema_alpha=0.3,
    tones=((2300.0, 0.5), (-4100.0, 0.05)), full_scale=0.8,
    description="RTL-SDR v3 replay 3.2 MS/s uint8 IQ, decim 16, 4096-pt FFT, "
                "EMA alpha 0.3, flip")

# BASELINE.json configs[2]: offline waterfall, no zoom, 65536-pt Hann.
CFG3 = Workload(
    name="cfg3", fs=2.4e6, fft_size=65536, fft_ratio=1, fft_avg=16,
    window="hann", dtype="c64", crop=None, tones=((301234.5, 0.5),),
    description="batched offline waterfall, complex64, 65536-pt FFT, 50% "
                "overlap Hann, rows of 2^20 samples")

# BASELINE.json configs[3]: one of 64 virtual receivers over a 20 MS/s stream.
CFG4 = Workload(
    name="cfg4", fs=20e6, fft_size=8192, fft_ratio=16, fft_avg=int(20e6 / 8192 / 10),
    window="hamming", dtype="c64",
    tones=tuple((-9.5e6 + i * 19e6 / 63 + 7000.0, 0.1) for i in range(64)),
    description="64 virtual receivers over one 20 MS/s stream, 8192-pt FFT")

# cfg2's geometry fed with SoapySDR CS16 (interleaved int16 IQ, 4 B/sample): the wire format
# the reference lets SoapySDR widen to CF32 on the host (S:602) crosses PCIe as it left the device
CFG2_CS16 = Workload(
    name="cfg2cs16", fs=3.2e6, fft_size=4096, fft_ratio=16, fft_avg=int(3.2e6 / 4096 / 10),
    window="hamming", dtype="cs16", flip=False, ema_alpha=0.3,
    tones=((2300.0, 0.5), (-4100.0, 0.05)), full_scale=0.8,
    description="cfg2's geometry with SoapySDR CS16 int16 IQ samples (4 B/sample), EMA alpha 0.3")

WORKLOADS = {w.name: w for w in (CFG1, CFG2, CFG3, CFG4, CFG2_CS16)}


def cfg4_centres() -> np.ndarray:
    return -9.5e6 + np.arange(64) * 19e6 / 63


def make_frame_complex(w: Workload, frame_index: int = 0, n: int | None = None
                       ) -> np.ndarray:
    """One frame of the tone+noise model as complex128 (before any
    quantisation / dtype cast).  Tone phase is continuous across frames; the
    noise stream of frame i is seeded (seed, i)."""
    n = w.frame_len if n is None else int(n)
    k = np.arange(n, dtype=np.float64) + float(frame_index) * n
    x = np.zeros(n, dtype=np.complex128)
    for f, a in w.tones:
        x += a * np.exp(2j * np.pi * (f / w.fs) * k)
    rng = np.random.default_rng([w.seed, frame_index])
    x += w.sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return x


def quantise_u8(x: np.ndarray) -> np.ndarray:
    """complex -> interleaved uint8 I,Q offset binary: clip(rint(127.5*(1+s)))."""
    iq = np.empty(2 * len(x), dtype=np.float64)
    iq[0::2] = x.real
    iq[1::2] = x.imag
    return np.clip(np.rint(127.5 * (1.0 + iq)), 0, 255).astype(np.uint8)


def quantise_cs16(x: np.ndarray) -> np.ndarray:
    """complex -> interleaved int16 I,Q (SoapySDR CS16): clip(rint(32768*s))."""
    iq = np.empty(2 * len(x), dtype=np.float64)
    iq[0::2] = x.real
    iq[1::2] = x.imag
    return np.clip(np.rint(32768.0 * iq), -32768, 32767).astype(np.int16)


def make_frame(w: Workload, frame_index: int = 0, n: int | None = None
               ) -> np.ndarray:
    """One frame in the workload's wire dtype (complex64 or interleaved u8)."""
    x = make_frame_complex(w, frame_index, n)
    if w.dtype in ("u8", "cs16"):
        peak = sum(a for _f, a in w.tones) + 4 * w.sigma
        if peak > w.full_scale:
            x = x * (w.full_scale / peak)
        return quantise_u8(x) if w.dtype == "u8" else quantise_cs16(x)
    return x.astype(np.complex64)


def make_frames(w: Workload, nframes: int, distinct: int | None = None,
                n: int | None = None) -> np.ndarray:
    """(nframes, frame_len[*2]) array.  ``distinct`` < nframes repeats the
    first ``distinct`` generated frames (cheap way to fill >L2-sized batches
    for bench.py; the arithmetic does not depend on the values)."""
    distinct = nframes if distinct is None else min(distinct, nframes)
    base = np.stack([make_frame(w, i, n) for i in range(distinct)])
    if distinct == nframes:
        return base
    reps = -(-nframes // distinct)
    return np.ascontiguousarray(np.tile(base, (reps, 1))[:nframes])
