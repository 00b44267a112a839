"""scipy-shaped entry points: the two scipy.signal calls the reference's hot
path makes (pypanadapter_spectrum.py:2098, 2111; pypanadapter_thread.py:1534,
1536, 1538), executed by the sm_100a engine.

    decimate(x, 2)                                   -> ndarray
    welch(x, fs, window=..., nperseg=N, nfft=N)      -> (freqs, Pxx)

Only the argument combinations the reference uses are accepted; anything else
raises (there is no silent CPU path).
"""
from __future__ import annotations

import numpy as np

from .engine import ZoomPSD, default_engine


def _as_complex_1d(x):
    x = np.asarray(x)
    if x.ndim != 1:
        raise ValueError("only 1-D sample chunks are supported")
    if not np.issubdtype(x.dtype, np.number):
        raise TypeError("numeric samples expected")
    return x


def decimate(x, q, n=None, ftype="iir", axis=-1, zero_phase=True, *, engine: ZoomPSD | None = None):
    """``scipy.signal.decimate(x, 2)``: order-8 Chebyshev-I, 0.05 dB, zero
    phase (sosfiltfilt), every 2nd sample (scipy:_signaltools.py:5206-5369).
    Powers of two run the cascade ``decimate(., 2)`` log2(q) times, which is
    how the reference reaches its zoom ratios (S:2096-2098)."""
    if n is not None or ftype != "iir" or axis != -1 or not zero_phase:
        raise NotImplementedError("only scipy.signal.decimate(x, q) with its defaults is accelerated")
    q = int(q)
    if q < 2 or q & (q - 1):
        raise NotImplementedError("decimation factor must be a power of two >= 2")
    x = _as_complex_1d(x)
    eng = engine or default_engine()
    # a minimal Welch plan rides along (the chain always ends in one); its cost
    # is negligible next to the filter
    eng.configure(1.0, 32, q, len(x), "boxcar", crop=None, no_lo=True, linear=True)
    eng.process(x)
    y = eng.read_decimated()
    out_dtype = np.complex128 if x.dtype in (np.complex128, np.float64) else np.complex64
    y = y.astype(out_dtype)
    if np.isrealobj(x):
        return y.real.copy()
    return y


def welch(x, fs=1.0, window="hann", nperseg=None, noverlap=None, nfft=None, detrend="constant",
          return_onesided=True, scaling="density", axis=-1, average="mean", *,
          engine: ZoomPSD | None = None):
    """``scipy.signal.welch(x, fs, window=w, nperseg=N, nfft=N)``: two-sided
    density in natural FFT order for complex input, one-sided (N/2+1 bins)
    for real input (scipy:_spectral_py.py:515, :915-916)."""
    if noverlap is not None or detrend != "constant" or scaling != "density" or axis != -1 \
            or average != "mean":
        raise NotImplementedError("only the welch defaults the reference relies on are accelerated")
    real = np.isrealobj(np.asarray(x))
    if real and not return_onesided:
        raise NotImplementedError("two-sided spectra of real input are not accelerated")
    x = _as_complex_1d(x)
    if nperseg is None:
        nperseg = 256
    if nfft is None:
        nfft = nperseg
    if nfft != nperseg:
        raise NotImplementedError("nfft must equal nperseg (S:2111)")
    eng = engine or default_engine()
    if real:
        eng.configure(fs, nfft, 1, len(x), window, crop=None, linear=True, onesided=True)
        p = eng.process(x)[0]                       # fftshifted one-sided bins, as the reference crops them
        out_dtype = np.float32 if np.asarray(x).dtype == np.float32 else np.float64
        return np.fft.rfftfreq(nfft, 1.0 / fs), np.fft.ifftshift(p).astype(out_dtype)
    eng.configure(fs, nfft, 1, len(x), window, crop=None, linear=True)
    p = eng.process(x)[0]
    pxx = np.fft.ifftshift(p)
    freqs = np.fft.fftfreq(nfft, 1.0 / fs)
    out_dtype = np.float64 if x.dtype == np.complex128 else np.float32      # scipy:_spectral_py.py:969
    return freqs, pxx.astype(out_dtype)
