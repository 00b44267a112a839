"""Taper (window) design and the taper dialog's preview on the device.

Mirrors what ``FFTTaperingControl.ShowCurve`` computes for its two plots
(pypanadapter_spectrum.py:1354-1379):

    taperdata = scipy.signal.get_window(AppState.fft_tapering, 51)
    fft = np.fft.fft(taperdata, 2048) / (len(taperdata) / 2.0)
    taperfft = 20 * np.log10(np.abs(fft / np.max(np.abs(fft))))

``get_window`` below takes the dialog's ``str`` / ``(name, p0[, p1])`` forms.
The closed-form families are evaluated on the GPU (fp64, C ABI
``zfb_taper_design``); ``chebwin``, ``dpss`` and ``slepian`` (polynomial design /
eigenproblem) are built by scipy on the host and only their preview spectrum
runs on the device.  The PSD path itself keeps scipy's tables
(``engine.window_table``): the same call the reference's welch makes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .engine import ZoomPSD, default_engine

# name -> (device kind, number of shape parameters)
_DEVICE = {
    "boxcar": (0, 0), "box": (0, 0), "ones": (0, 0), "rect": (0, 0), "rectangular": (0, 0),
    "triang": (1, 0), "triangle": (1, 0), "tri": (1, 0),
    "bartlett": (2, 0), "bart": (2, 0), "brt": (2, 0),
    "hann": (3, 0), "han": (3, 0), "hamming": (4, 0), "hamm": (4, 0), "ham": (4, 0),
    "blackman": (5, 0), "black": (5, 0), "blk": (5, 0),
    "nuttall": (6, 0), "nutl": (6, 0), "nut": (6, 0),
    "blackmanharris": (7, 0), "blackharr": (7, 0), "bkh": (7, 0),
    "flattop": (8, 0), "flat": (8, 0), "flt": (8, 0),
    "bohman": (9, 0), "bman": (9, 0), "bmn": (9, 0),
    "barthann": (10, 0), "brthan": (10, 0), "bth": (10, 0),
    "parzen": (11, 0), "parz": (11, 0), "par": (11, 0),
    "kaiser": (12, 1), "ksr": (12, 1),
    "gaussian": (13, 1), "gauss": (13, 1), "gss": (13, 1),
    "general gaussian": (14, 2), "general_gaussian": (14, 2), "general gauss": (14, 2),
    "general_gauss": (14, 2), "ggs": (14, 2),
    "exponential": (15, 2), "poisson": (15, 2),
    "tukey": (16, 1), "tuk": (16, 1),
}


def on_device(window) -> bool:
    name = window if isinstance(window, str) else window[0]
    return name in _DEVICE


def get_window(window, nx: int, fftbins: bool = True, *, engine: ZoomPSD | None = None) -> np.ndarray:
    """``scipy.signal.get_window(window, nx, fftbins)`` as float64[nx]."""
    eng = engine or default_engine()
    name, args = (window, ()) if isinstance(window, str) else (window[0], tuple(window[1:]))
    if name not in _DEVICE:
        from scipy.signal import get_window as host_get_window
        return np.asarray(host_get_window(window, int(nx), fftbins=fftbins), dtype=np.float64)
    kind, npar = _DEVICE[name]
    if npar == 0 and args:
        raise ValueError("the %r window takes no parameter" % name)
    if kind in (12, 13, 16) and len(args) != 1:
        raise ValueError("The '%s' window needs one parameter -- pass a tuple." % name)
    p0 = p1 = 0.0
    if kind == 14:
        if len(args) != 2:
            raise ValueError("The 'general gaussian' window needs (power, std)")
        p0, p1 = float(args[0]), float(args[1])
    elif kind == 15:
        # scipy: exponential(M, center=None, tau=1.0), parameters given positionally
        M = nx + 1 if fftbins else nx
        center = (M - 1) / 2 if (len(args) < 1 or args[0] is None) else float(args[0])
        if not fftbins and len(args) >= 1 and args[0] is not None:
            raise ValueError("If sym==True, center must be None.")
        p0, p1 = center, (float(args[1]) if len(args) > 1 else 1.0)
    elif npar == 1:
        p0 = float(args[0])
    out = np.empty(int(nx), dtype=np.float64)
    eng._check(eng._lib.zfb_taper_design(eng._h, kind, p0, p1, int(nx), 1 if fftbins else 0,
                                         out.ctypes.data_as(C.POINTER(C.c_double))), "zfb_taper_design")
    return out


def preview_spectrum(taperdata, nfft: int = 2048, *, engine: ZoomPSD | None = None) -> np.ndarray:
    """``20*log10(abs(fft(taperdata, nfft)) / max)`` (S:1374-1376), float32[nfft]."""
    eng = engine or default_engine()
    t = np.ascontiguousarray(taperdata, dtype=np.float64)
    out = np.empty(int(nfft), dtype=np.float32)
    eng._check(eng._lib.zfb_taper_preview(eng._h, t.ctypes.data_as(C.POINTER(C.c_double)), t.size, int(nfft),
                                          out.ctypes.data_as(C.POINTER(C.c_float))), "zfb_taper_preview")
    return out


def show_curve(window, taper_size: int = 51, fft_size: int = 2048, *, engine: ZoomPSD | None = None):
    """The two curves of the dialog: (taperdata[taper_size], taperfft[fft_size])."""
    taperdata = get_window(window, taper_size, engine=engine)
    return taperdata, preview_spectrum(taperdata, fft_size, engine=engine)
