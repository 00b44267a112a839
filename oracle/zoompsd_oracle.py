"""CPU oracle for pypanadapter's zoom-FFT PSD path -- TEST INFRASTRUCTURE ONLY.

A numpy/scipy fp64 restatement of the ~30 lines the reference runs per frame.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this file; the product never does.

Citations: ``S:`` = /root/reference/pypanadapter_spectrum.py,
``T:`` = /root/reference/pypanadapter_thread.py, ``scipy:`` = scipy 1.18.1
``scipy/signal/`` (the arithmetic lives there; the reference pins no version).

Pinned by tests/golden/*.npz, which were produced by the reference's own
methods (oracle/ref_harness.py + oracle/make_golden.py) with numpy 2.3.5 /
scipy 1.18.1.  PARITY UNPINNED for the functions that restate libraries which are
neither installed nor vendored: ``rtlsdr_bytes_to_iq`` (pyrtlsdr), ``cs16_to_iq``
(SoapySDR's CS16 -> CF32 converter) and ``waterfall_indices`` / ``colormap_lut``
(pyqtgraph's level -> colour-index mapping and ColorMap table).

Two levels are provided for every stage:
* ``*_ref``      -- the same scipy call the reference makes (fast; used for
                    timing the CPU baseline and for full-size parity), and
* ``*_explicit`` -- the algorithm spelled out (odd extension, steady-state
                    initial conditions, biquad recurrences, segmenting,
                    detrend, window, FFT, mean, scale) so that the CUDA
                    kernels have a line-by-line statement to be checked
                    against.  tests/test_oracle.py asserts both agree.
"""
from __future__ import annotations

import math
import warnings

import numpy as np
import scipy.signal

# --------------------------------------------------------------------------
# constants of the reference's decimator: scipy.signal.decimate(x, 2)
#   scipy:_signaltools.py:5317-5319  n = 8 ; cheby1(n, 0.05, 0.8 / q)
# --------------------------------------------------------------------------
DECIM_ORDER = 8
DECIM_RIPPLE_DB = 0.05
DECIM_Q = 2
#   scipy:_signaltools.py:5186-5187  ntaps = 2*n_sections + 1 = 9
#   scipy:_signaltools.py:4938       edge  = ntaps * 3       = 27
DECIM_PADLEN = 27


def decim_sos() -> np.ndarray:
    """The 4-section SOS the reference's decimator uses (fp64)."""
    return scipy.signal.cheby1(DECIM_ORDER, DECIM_RIPPLE_DB, 0.8 / DECIM_Q,
                               output="sos")


# --------------------------------------------------------------------------
# K0/K1: sample conversion and the RTL "IQ inversion" flip
# --------------------------------------------------------------------------
def rtlsdr_bytes_to_iq(raw: np.ndarray) -> np.ndarray:
    """uint8 interleaved I,Q -> complex128, as pyrtlsdr's packed_bytes_to_iq.

    Called by the reference through ``driver.read_samples`` (S:543, T:460) and
    the async callback (S:448).  Upstream pyrtlsdr (un-vendored, unpinned,
    README.md:4):  ``iq = bytes.astype(float64).view(complex128);
    iq /= 127.5; iq -= (1 + 1j)``.  PARITY UNPINNED (library not installed).
    """
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    if raw.size % 2:
        raise ValueError("uint8 IQ stream must hold an even number of bytes")
    iq = raw.astype(np.float64).view(np.complex128)
    iq = iq / 127.5
    iq = iq - (1 + 1j)
    return iq


def cs16_to_iq(raw: np.ndarray) -> np.ndarray:
    """int16 interleaved I,Q (SoapySDR ``SOAPY_SDR_CS16``) -> complex128.

    The reference asks SoapySDR for ``SOAPY_SDR_CF32`` (S:602) and lets the
    library convert the device's native int16 on the host; SoapySDR's
    CS16 -> CF32 converter scales by 1/32768.  SoapySDR is not installed and not
    vendored: PARITY UNPINNED for this one function.
    """
    raw = np.ascontiguousarray(raw, dtype=np.int16)
    if raw.size % 2:
        raise ValueError("int16 IQ stream must hold an even number of values")
    return raw.astype(np.float64).view(np.complex128) / 32768.0


def flip_chunk(x: np.ndarray) -> np.ndarray:
    """``np.flip`` of the chunk: the reference's "IQ inversion" (S:460, S:543,
    T:460) is a time reversal of the whole chunk."""
    return np.flip(x)


# --------------------------------------------------------------------------
# K2: software LO and mix   (S:2090-2094, T:1526-1530)
# --------------------------------------------------------------------------
def lo_mix(x: np.ndarray, fs: float, f_demod: float = 1.0) -> np.ndarray:
    """``x * 2**.5 * exp(-2j*pi*f_demod*t)``, ``t[k] = k*(1/fs)``.

    The reference builds ``t = np.arange(0, n/fs, 1/fs)`` (S:2091-2092,
    T:1527-1528); that is elementwise identical to ``k*(1/fs)`` whenever its
    length equals ``n`` (for ~7 % of arbitrary n it has n+1 points and the
    reference raises a broadcast error at T:1530 -- not reproduced).
    ``f_demod`` is hard-wired to 1.0 in the reference (S:2090, T:1526).
    """
    n = len(x)
    t = np.arange(n) * (1 / fs)
    lo = 2 ** .5 * np.exp(-2j * np.pi * f_demod * t)
    return x * lo


# --------------------------------------------------------------------------
# K3: one decimate-by-2 stage   (S:2098, T:1534 -> scipy:_signaltools.py:5206)
# --------------------------------------------------------------------------
def decimate2_ref(x: np.ndarray) -> np.ndarray:
    """Exactly the reference's call."""
    return scipy.signal.decimate(x, 2)


def odd_ext(x: np.ndarray, n: int) -> np.ndarray:
    """scipy:_arraytools.py odd_ext: 2*x[0]-x[n:0:-1] | x | 2*x[-1]-x[-2:-n-2:-1]."""
    left = 2 * x[0] - x[n:0:-1]
    right = 2 * x[-1] - x[-2:-(n + 2):-1]
    return np.concatenate((left, x, right))


def sos_zi(sos: np.ndarray) -> np.ndarray:
    """Steady-state DF2T states for a unit step (scipy sosfilt_zi), written out.

    Section k sees a constant input equal to the DC gain of sections < k; its
    transposed-direct-form-II states for constant input c are
        s2 = (b2 - a2*H0) * c,   s1 = (b1 - a1*H0) * c + s2,
    with H0 = (b0+b1+b2)/(1+a1+a2).
    """
    zi = np.zeros((sos.shape[0], 2))
    c = 1.0
    for k, (b0, b1, b2, _a0, a1, a2) in enumerate(sos):
        h0 = (b0 + b1 + b2) / (1 + a1 + a2)
        s2 = (b2 - a2 * h0) * c
        s1 = (b1 - a1 * h0) * c + s2
        zi[k] = (s1, s2)
        c *= h0
    return zi


def _sosfilt_explicit(sos, x, zi):
    """Cascade of biquads, transposed direct form II, one lfilter per section
    (the recurrence  y=b0 x+s1; s1=b1 x-a1 y+s2; s2=b2 x-a2 y)."""
    y = x
    for k in range(sos.shape[0]):
        b, a = sos[k, :3], sos[k, 3:]
        y, _ = scipy.signal.lfilter(b, a, y, zi=zi[k])
    return y


def decimate2_explicit(x: np.ndarray) -> np.ndarray:
    """sosfiltfilt + [::2] spelled out (scipy:_signaltools.py:5091-5204, 5367).

    1. ext = odd_ext(x, 27)
    2. forward cascade over ext, initial state = zi * ext[0]
    3. backward cascade over the forward output, initial state = zi * y_fwd[-1]
    4. drop the 27-sample pads, keep every 2nd sample starting at 0.
    """
    x = np.asarray(x)
    if x.shape[0] <= DECIM_PADLEN:
        raise ValueError("The length of the input vector x must be greater "
                         "than padlen, which is %d." % DECIM_PADLEN)
    sos = decim_sos()
    zi = sos_zi(sos)
    ext = odd_ext(x, DECIM_PADLEN)
    y = _sosfilt_explicit(sos, ext, zi * ext[0])
    y = _sosfilt_explicit(sos, y[::-1], zi * y[-1])[::-1]
    y = y[DECIM_PADLEN:-DECIM_PADLEN]
    return y[::2]


def zoom_mix_decimate(x, fs, ratio, f_demod=1.0, explicit=False):
    """zoomfft (S:2088-2100) / the zoom branch of PSD.update (T:1525-1534)."""
    x_mix = lo_mix(x, fs, f_demod)
    power2 = int(np.log2(ratio))                     # S:2096, T:1532
    dec = decimate2_explicit if explicit else decimate2_ref
    for _ in range(power2):
        x_mix = dec(x_mix)
    return x_mix


# --------------------------------------------------------------------------
# K4-K6: Welch PSD   (S:2111, T:1536/1538 -> scipy:_spectral_py.py:515-972)
# --------------------------------------------------------------------------
def welch_ref(x, fs, window, nfft):
    """Exactly the reference's call (returns natural FFT order, two-sided for
    complex input)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")              # nperseg > len(x) warning
        _f, p = scipy.signal.welch(x, fs, window=window, nperseg=nfft,
                                   nfft=nfft)
    return p


def welch_plan(n: int, nfft: int):
    """(nperseg, hop, nseg) scipy uses for ``welch(x[n], nperseg=nfft,
    nfft=nfft)``: nperseg = min(nfft, n) (scipy:_spectral_py.py:897-900),
    noverlap = nperseg//2 (:912), nseg = (n - noverlap)//hop."""
    nperseg = min(nfft, n)
    noverlap = nperseg // 2
    hop = nperseg - noverlap
    nseg = (n - noverlap) // hop
    return nperseg, hop, nseg


def welch_explicit(x, fs, window, nfft):
    """Welch spelled out: periodic window, 50 % overlap, per-segment complex
    mean removed before windowing, nfft-point FFT (zero padded if the segment
    is shorter), |X|^2 averaged over segments, / (fs * sum(w^2))."""
    x = np.asarray(x)
    n = len(x)
    nperseg, hop, nseg = welch_plan(n, nfft)
    win = scipy.signal.get_window(window, nperseg)   # fftbins=True: periodic
    scale = 1.0 / (fs * (win * win).sum())
    acc = np.zeros(nfft)
    for s in range(nseg):
        seg = x[s * hop:s * hop + nperseg]
        seg = seg - seg.mean()                       # detrend='constant'
        spec = np.fft.fft(seg * win, nfft)
        acc += spec.real ** 2 + spec.imag ** 2
    p = acc / nseg * scale
    if np.isrealobj(x):                              # one-sided for real input
        p = p[:nfft // 2 + 1].copy()
        if nfft % 2:
            p[1:] *= 2
        else:
            p[1:-1] *= 2
    return p


# --------------------------------------------------------------------------
# K7/K8: fftshift + centre crop, 20*log10(abs(.))
# --------------------------------------------------------------------------
def crop_width(fft_size: int, fft_ratio, crop) -> int:
    """Row width W.  crop='thread': T:1542  ``2*int(.5*N/R)``;
    crop=int: S:2114 ``N_WIN`` (1024 at start S:1712, int(N/R) after an FFT
    size change S:1757); crop=None: all N bins."""
    if crop is None:
        return int(fft_size)
    if isinstance(crop, str):
        if crop != "thread":
            raise ValueError("crop must be 'thread', an int or None")
        return 2 * int(.5 * fft_size / fft_ratio)
    return 2 * (int(crop) // 2)


def shift_crop(spec: np.ndarray, fft_size: int, width: int) -> np.ndarray:
    """``np.fft.fftshift(spec)[N//2-W//2 : N//2+W//2]`` (S:2114, T:1543)."""
    half = width // 2
    return np.fft.fftshift(spec)[fft_size // 2 - half:fft_size // 2 + half]


def to_db20(spec: np.ndarray) -> np.ndarray:
    """``20*np.log10(abs(spec))`` (S:2117-2119, T:1548) -- twice a power dB."""
    with np.errstate(divide="ignore"):
        return 20 * np.log10(abs(spec))


# --------------------------------------------------------------------------
# whole frame
# --------------------------------------------------------------------------
def zoom_psd_power(chunk, fs, fft_size, fft_ratio, window, *, f_demod=1.0,
                   crop="thread", flip=False, explicit=False):
    """Linear (pre-dB) cropped PSD row of one chunk.

    chunk: complex ndarray, uint8 interleaved I,Q (converted as pyrtlsdr) or
    int16 interleaved I,Q (SoapySDR CS16).
    Mirrors ApplicationDisplay.update (S:2102-2114) for ``crop=N_WIN`` and
    PSD.update (T:1513-1543) for ``crop='thread'``.
    """
    chunk = np.asarray(chunk)
    if chunk.dtype == np.uint8:
        chunk = rtlsdr_bytes_to_iq(chunk)
    elif chunk.dtype == np.int16:
        chunk = cs16_to_iq(chunk)
    if flip:
        chunk = flip_chunk(chunk)
    if fft_ratio > 1:                                # S:2108, T:1525
        chunk = zoom_mix_decimate(chunk, fs, fft_ratio, f_demod, explicit)
    welch = welch_explicit if explicit else welch_ref
    spec = welch(chunk, fs, window, fft_size)
    width = crop_width(fft_size, fft_ratio, crop)
    return shift_crop(spec, fft_size, width)


def zoom_psd(chunk, fs, fft_size, fft_ratio, window, *, f_demod=1.0,
             crop="thread", flip=False, explicit=False):
    """One dB20 waterfall row (float64)."""
    return to_db20(zoom_psd_power(chunk, fs, fft_size, fft_ratio, window,
                                  f_demod=f_demod, crop=crop, flip=flip,
                                  explicit=explicit))


def ema_rows_db20(power_rows: np.ndarray, alpha: float) -> np.ndarray:
    """Exponential averaging across rows (NOT in the reference -- SURVEY §0:
    builder-defined extension).  State is linear power; first row initialises:
        a_0 = p_0 ;  a_i = alpha*p_i + (1-alpha)*a_{i-1} ;  row_i = 20log10(a_i)
    """
    power_rows = np.asarray(power_rows, dtype=np.float64)
    out = np.empty_like(power_rows)
    a = None
    for i, p in enumerate(power_rows):
        a = p.copy() if a is None else alpha * p + (1 - alpha) * a
        out[i] = to_db20(a)
    return out


# --------------------------------------------------------------------------
# host-side containers (restated for tests of the product's mirrors)
# --------------------------------------------------------------------------
class DataOracle:
    """Fold-back sample buffer, T:1400-1483, without the QMutex / NewtRap /
    sleep (those are pacing, out of scope)."""

    def __init__(self, chunk_size=8196 * 2):                     # T:1402
        self.chunk_size = chunk_size
        self.max_size = self.chunk_size * 16                     # T:1406
        self.target_size = self.max_size * .9                    # T:1407

    def new_complex(self):                                       # T:1419-1423
        self.data = np.zeros(self.max_size) * (1 + 1j)
        self.real = False
        return self._new_common()

    def new_real(self):                                          # T:1413-1417
        self.data = np.zeros(self.max_size)
        self.real = True
        return self._new_common()

    def _new_common(self):                                       # T:1425-1431
        self.size = 0
        self.real_size = 0
        self.total_size = 0
        return self

    def add(self, chunk):                                        # T:1433-1457
        length = len(chunk)
        new_size = self.size + length
        if new_size > self.max_size:
            self.size = 0
            new_size = length
        self.target_size = np.clip(self.target_size, 8192, self.max_size)
        self.data[self.size:new_size] = chunk
        self.size = new_size
        self.real_size = max(self.real_size, self.size)
        self.total_size += length

    def take(self):
        """get_data_start .. get_data_end (T:1459-1468) as one call: returns a
        copy of data[:real_size] and resets the counters."""
        out = self.data[:self.real_size].copy()
        self.size = 0
        self.real_size = 0
        self.total_size = 0
        return out


def waterfall_init(fftwidth: int) -> np.ndarray:
    """Waterfall.init_image (S:1625-1635): (w//4, w) image at -500 with the
    two outer grid columns zeroed."""
    img = -500 * np.ones((fftwidth // 4, fftwidth))
    img[:, 0] = 0
    img[:, fftwidth - 1] = 0
    return img


def waterfall_update(img: np.ndarray, psd: np.ndarray, scroll: int = 1):
    """Waterfall.image_update (S:1638-1662) minus setImage: zero the three
    grid bins IN PLACE on psd, insert as last row, roll by -scroll, draw the
    tick marks.  Returns the new image (psd is mutated like the reference)."""
    w = np.size(psd)
    if img is None or img.shape[1] != w:
        img = waterfall_init(w)
    for x in (0, w // 2, w - 1):
        psd[x] = 0
    img[-1:] = psd
    img = np.roll(img, -scroll, 0)
    for i, x in enumerate(range(0, w - 1, (w // 10))):
        if i != 5 and i != 10:
            if scroll > 0:
                for y in range(5, 15):
                    img[y, x] = 0
            else:
                for y in range(-10, -2):
                    img[y, x] = 0
    return img


def waterfall_autolevel(img: np.ndarray):
    """Waterfall.autolevel (S:1676): the 2nd / 98th percentile of the image
    pixels below zero (the -500 fill counts, grid and tick zeros do not)."""
    lo, hi = np.percentile(img[img < 0], [2, 98])
    return float(lo), float(hi)


def waterfall_indices(img: np.ndarray, minlev: float = -220, maxlev: float = -120,
                      ncolors: int = 256) -> np.ndarray:
    """Colour-table index pyqtgraph's ImageItem computes for every pixel of
    the image the reference hands it (S:1664 setImage(..., autoLevels=False)
    after S:1594 setLevels([minlev, maxlev]) and S:1623 setLookupTable(256
    entries)).  PARITY UNPINNED: pyqtgraph is a third-party dependency of the
    reference (README.md:7 `pip3 install pyqtgraph`, no version), not vendored
    and not installed here; this restates its published algorithm
    (pyqtgraph/graphicsItems/ImageItem.py `_try_rescale_float` ->
    functions.rescaleData): scale = ncolors / (max - min); d = (img - min) *
    scale; clip to [0, ncolors-1]; truncate to an unsigned integer."""
    d = (np.asarray(img, dtype=np.float64) - float(minlev)) * (ncolors / (float(maxlev) - float(minlev)))
    d = np.clip(d, 0, ncolors - 1)
    return d.astype(np.uint8)


def waterfall_rgba(indices: np.ndarray, lut: np.ndarray) -> np.ndarray:
    """lut[index]: the table lookup of pyqtgraph's makeARGB / ImageItem
    render step (same caveat as waterfall_indices)."""
    return np.asarray(lut, dtype=np.uint8)[indices]


def colormap_lut(pos, color, npts: int = 256) -> np.ndarray:
    """ColorMap(pos, color).getLookupTable(0.0, 1.0, 256) as called at
    S:1622-1623 (pyqtgraph/colormap.py, published algorithm, unpinned): RGB(A)
    linearly interpolated between the stops over linspace(0, 1, npts), returned
    as uint8 with alpha 255.  Colour stops are taken modulo 256 like numpy < 2
    did for the reference's out-of-range 'Default' stop 2020 (S:1581; numpy >=
    2 raises OverflowError there)."""
    pos = np.asarray(pos, dtype=np.float64)
    col = (np.asarray(color, dtype=np.int64) % 256).astype(np.float64)
    x = np.linspace(0.0, 1.0, npts)
    out = np.empty((npts, 4), dtype=np.uint8)
    for c in range(4):
        out[:, c] = (np.interp(x, pos, col[:, c]) + 0.5).astype(np.uint8)
    return out


# --------------------------------------------------------------------------
# closed forms used as known-answer tests (SURVEY §8a "dB20 note")
# --------------------------------------------------------------------------
def tone_peak_db20(amplitude, fs, window, fft_size, zoomed: bool) -> float:
    """dB20 reading of a bin-centred complex tone: 20log10(g*A^2*(sum w)^2 /
    (fs*sum w^2)), g = 2 when zoomed (the sqrt(2) LO, S:2093) else 1."""
    w = scipy.signal.get_window(window, fft_size)
    g = 2.0 if zoomed else 1.0
    return 20 * math.log10(g * amplitude ** 2 * w.sum() ** 2
                           / (fs * (w * w).sum()))
