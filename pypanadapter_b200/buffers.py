"""Host-side mirrors of the reference's buffer objects around the hot path.

Same names, attributes and call order as the reference so that its front-ends
and Qt display code run untouched:

* ``Data``      -- pypanadapter_thread.py:1400-1483.  The fold-back sample
                   buffer; its storage is the engine's *pinned* sample ring and
                   every ``add`` is followed at once by an async H2D copy on
                   the engine's copy stream (C ABI ``zfb_samples_*``).
* ``PSD``       -- pypanadapter_thread.py:1485-1549.  ``update()`` turns
                   whatever is in ``Data`` into one dB20 row on the GPU and
                   publishes it as ``self.psd`` under ``self.lock``.
* ``Waterfall`` -- pypanadapter_spectrum.py:1625-1664.  Rows live in the
                   engine's device-resident ring; ``img_array`` is assembled
                   (grid, tick marks, scroll order) only when it is read.

Pacing (NewtRap delay/target control, T:1411,1457,1463,1496-1511) is host
scheduling and out of scope: ``delay_time`` stays 0 unless a controller
object is handed in.
"""
from __future__ import annotations

import threading
import time
import weakref

import numpy as np

from .engine import ZoomPSD, default_engine


class _Mutex:
    """QMutex-shaped lock (``lock()`` / ``unlock()``), T:1403, T:1492."""

    def __init__(self):
        self._l = threading.Lock()

    def lock(self):
        self._l.acquire()

    def unlock(self):
        if self._l.locked():
            self._l.release()


class Data:
    """Fold-back sample buffer in pinned host memory (T:1400-1483)."""

    def __init__(self, chunk_size=8196 * 2, *, engine: ZoomPSD | None = None, controller=None):
        self.engine = engine or default_engine()
        self.lock = _Mutex()                                   # T:1403
        self.chunk_size = chunk_size                           # T:1405
        self.max_size = self.chunk_size * 16                   # T:1406
        self.target_size = self.max_size * .9                  # T:1407
        self.delay_time = 0.                                   # T:1410
        self.NR = controller                                   # T:1411 (pacing: optional)
        self.wire = None
        self.real = False
        self.data = None
        self._detached = False           # True once another Data has taken over the engine's ring
        self.size = self.real_size = self.total_size = 0
        # event-driven hand-off for a GPU consumer: add() raises the flag when
        # `ready_size` samples have arrived since the last take, PSD.run(event_driven=True)
        # waits on it instead of sleeping out a fixed frame time (T:1498-1511)
        self._ready = threading.Condition()
        self.ready_size = 0

    # -- allocation (T:1413-1431) -------------------------------------------
    def _new(self, wire, real):
        # An engine owns ONE pinned sample ring; creating it again frees the previous one
        # (the reference re-creates Data when its main loop restarts ApplicationDisplay,
        # T:2509-2543).  The previous owner is detached first: its `data` becomes a private
        # array, so a late add() from an old reader thread never writes freed pinned memory.
        prev_ref = getattr(self.engine, "_samples_owner", None)
        prev = prev_ref() if prev_ref is not None else None
        if prev is not None and prev is not self:
            prev._detach()
        self.lock.lock()
        self.wire = wire
        self.real = real
        self.data = self.engine.samples_create(self.max_size, wire)   # zero-filled, pinned
        self._detached = False
        self.engine._samples_owner = weakref.ref(self)
        return self.new_common()

    def _detach(self):
        self.lock.lock()
        try:
            if self.data is not None and not self._detached:
                self.data = np.array(self.data)                # private copy, ordinary memory
            self._detached = True
        finally:
            self.lock.unlock()

    def new_real(self):
        """Real samples (AudioPan) are stored as complex64 with zero imaginary part."""
        return self._new("c64", True)

    def new_complex(self):
        return self._new("c64", False)

    def new_u8(self):
        """RTL-SDR wire format: interleaved uint8 I,Q (2*max_size bytes); the
        conversion (S:543) then happens on the device."""
        return self._new("u8", False)

    def new_cs16(self):
        """SoapySDR CS16 wire format: interleaved int16 I,Q (2*max_size values),
        widened to complex64 (/32768) on the device."""
        return self._new("cs16", False)

    def new_common(self):
        self.size = 0            # position of the next entry
        self.real_size = 0       # high-water mark since the last take
        self.total_size = 0      # samples received since the last take
        self.delay_time = .01 if self.NR is not None else 0.
        self.lock.unlock()
        return self

    # -- producer (T:1433-1457) ------------------------------------------------
    def add(self, chunk):
        chunk = np.asarray(chunk)
        per = 1 if self.wire == "c64" else 2
        if self.wire == "u8" and chunk.dtype != np.uint8:
            raise TypeError("this Data holds raw uint8 IQ")
        if self.wire == "cs16" and chunk.dtype != np.int16:
            raise TypeError("this Data holds raw int16 IQ")
        if len(chunk) % per:
            raise ValueError("interleaved IQ chunk must hold an even number of values")
        length = len(chunk) // per
        self.lock.lock()
        try:
            new_size = self.size + length
            if new_size > self.max_size:           # fold back on overflow
                self.size = 0
                new_size = length
            self.target_size = np.clip(self.target_size, 8192, self.max_size)
            if new_size > self.max_size:
                # the reference fails here too (broadcast error at T:1447)
                raise ValueError("could not broadcast input array from shape (%d,) into shape (%d,)"
                                 % (length, self.max_size - self.size))
            if not self._detached:
                self.engine.samples_begin_write(self.size, length)
            np.copyto(self.data[self.size * per:new_size * per], chunk, casting="unsafe")
            if not self._detached:
                self.engine.samples_commit(self.size, length)  # async H2D of this chunk
            self.size = new_size
            self.real_size = max(self.real_size, self.size)
            self.total_size += length
            ready = self.ready_size > 0 and self.real_size >= self.ready_size
        finally:
            self.lock.unlock()
        if ready:
            with self._ready:
                self._ready.notify_all()
        if self.delay_time:
            time.sleep(abs(self.delay_time))                   # T:1457

    # -- consumer (T:1459-1468) ------------------------------------------------
    def get_data_start(self):
        self.lock.lock()

    def get_data_end(self):
        if self.NR is not None:
            self.delay_time = self.NR.next(self.total_size)
        self.size = 0
        self.real_size = 0
        self.total_size = 0
        self.lock.unlock()

    def wait_ready(self, nsamples: int, timeout: float | None = None) -> bool:
        """Block until at least ``nsamples`` samples have arrived since the last
        take (or ``timeout`` seconds pass); True when they are there."""
        self.ready_size = int(nsamples)
        with self._ready:
            return self._ready.wait_for(lambda: self.real_size >= self.ready_size, timeout)

    @property
    def target(self):
        return self.target_size

    @target.setter
    def target(self, t):
        if t <= self.max_size and t >= getattr(self, "min_target", 0):
            self.target_size = t
            if self.NR is not None:
                self.NR.target = t

    @property
    def maxsize(self):
        return self.max_size


class PSD:
    """PSD worker (T:1485-1549): ``update()`` = one row from whatever is in
    ``Data``.  ``state`` is the reference's AppState (or any object with
    ``fft_size, fft_ratio, fft_tapering, panadapter.SampleRate``)."""

    FRAME_TIME = 0.1                                           # T:68-69

    def __init__(self, dataclass: Data, state, *, engine: ZoomPSD | None = None,
                 flip=False, ema_alpha=None):
        self.state = state
        self.dataclass = dataclass
        self.engine = engine or dataclass.engine
        self.psd = np.zeros(state.fft_size)                    # T:1490 default blank
        self.lock = _Mutex()
        self.loop = True
        self.flip = flip
        self.ema_alpha = ema_alpha

    def run(self, event_driven: bool = False, frame_samples: int | None = None):
        """T:1498-1511 without NewtRap: one row per .95 FRAME_TIME.  With
        ``event_driven`` the loop instead wakes when ``frame_samples`` (default
        ``fft_size * fft_ratio``-ish: one full Welch segment after decimation)
        new samples are in the ring -- the GPU needs microseconds per row, so
        the row rate is set by the sample source, not by a sleep."""
        while self.loop:
            if event_driven:
                st = self.state
                need = frame_samples or int(st.fft_size * max(1, st.fft_ratio))
                need = min(need, self.dataclass.max_size)
                if not self.dataclass.wait_ready(need, timeout=self.FRAME_TIME):
                    continue
                self.update()
                continue
            target = time.monotonic() + .95 * self.FRAME_TIME
            self.update()
            end = time.monotonic()
            if end < target:
                time.sleep(target - end)

    def update(self):
        st = self.state
        d = self.dataclass
        d.get_data_start()
        size = d.real_size
        row = None
        try:
            if size >= st.fft_size and not getattr(d, "_detached", False):   # T:1522
                # enqueue under the lock: the kernels read the device mirror the
                # producer has just finished filling; afterwards it moves on to
                # the other mirror, so nothing is overwritten under the kernels
                self.engine.configure(st.panadapter.SampleRate, st.fft_size, st.fft_ratio, size,
                                      st.fft_tapering, dtype=d.wire, flip=self.flip,
                                      crop="thread", ema_alpha=self.ema_alpha,
                                      onesided=d.real and not st.fft_ratio > 1)    # T:1538: welch of real samples
                row = self.engine.samples_process()
        finally:
            d.get_data_end()
        if row is None:
            return
        self.lock.lock()
        self.psd = row.astype(np.float64)                      # T:1548
        self.lock.unlock()


def tick_columns(fftwidth: int):
    """Columns the reference marks (S:1655-1656)."""
    return [x for i, x in enumerate(range(0, fftwidth - 1, (fftwidth // 10))) if i != 5 and i != 10]


class Waterfall:
    """Waterfall row store (S:1625-1664) on the engine's device-resident ring.

    ``image_update(psd)`` only zeroes the three grid bins in place on ``psd``
    (as the reference does, S:1647-1648) and appends the row to the device
    ring if the engine has not put it there already; the ``(w//4, w)``
    ``img_array`` with grid, tick marks and scroll order is read back and
    assembled when accessed (display time)."""

    def __init__(self, engine: ZoomPSD | None = None, *, scroll=1, on_image=None):
        self.engine = engine or default_engine()
        self.scroll = scroll
        self.fftwidth = 0
        self.rows_seen = 0               # rows since init_image
        self._pending_engine_rows = 0
        self.on_image = on_image         # e.g. pg.ImageItem.setImage wrapper

    def init_image(self):                                      # S:1625-1635
        self.rows_seen = 0
        self._pending_engine_rows = 0    # a resized ring starts empty: the row in hand is pushed
        # the ring is the waterfall's own, `fftwidth` wide whatever the engine computes next
        # (the reference takes rows of any width: the blank np.zeros(fft_size) of T:1490 before
        # the first real row, the stale width right after a zoom click)
        self.engine.ring_configure(max(4, self.fftwidth // 4), self.fftwidth)

    def note_engine_rows(self, n=1):
        """The engine has just appended ``n`` rows to its ring itself (PSD.update
        / zoom_psd on this engine): image_update will not push them again."""
        self._pending_engine_rows += n

    def image_update(self, psd):                               # S:1638-1664
        fftwidth = np.size(psd)
        if fftwidth != self.fftwidth or self.engine.ring_width != fftwidth:
            self.fftwidth = fftwidth                           # S:1641-1643
            self.init_image()
        for x in (0, fftwidth // 2, fftwidth - 1):             # grid, in place like the reference
            psd[x] = 0
        if self._pending_engine_rows > 0:
            self._pending_engine_rows -= 1
        else:
            self.engine.push_rows(psd)
        self.rows_seen += 1
        if self.on_image is not None:
            self.on_image(self.img_array.T)                    # S:1664

    # reference defaults (S:1592-1594)
    minlev = -220
    maxlev = -120

    @property
    def img_array(self) -> np.ndarray:
        """The reference's ``img_array`` after the same sequence of updates
        (-500 fill, grid columns, scroll order, tick marks: S:1625-1662),
        assembled on the device from the ring when it is displayed."""
        w = self.fftwidth
        return self.engine.ring_image(w // 4, self.scroll, self.rows_seen, "f32").astype(np.float64)

    def image_indices(self, levels=None) -> np.ndarray:
        """8-bit colour indices pyqtgraph's ImageItem would look up for
        ``img_array`` with ``setLevels(levels)`` and a 256-entry table
        (S:1594, S:1623, S:1664): the only thing read back for display."""
        lo, hi = levels if levels is not None else (self.minlev, self.maxlev)
        return self.engine.ring_image(self.fftwidth // 4, self.scroll, self.rows_seen, "u8", levels=(lo, hi))

    def image_rgba(self, lut, levels=None) -> np.ndarray:
        """``lut[image_indices]``: (h, w, 4) uint8, lut = the (256, 4) table of
        ``ColorMap.getLookupTable(0.0, 1.0, 256)`` (S:1623) with alpha."""
        lo, hi = levels if levels is not None else (self.minlev, self.maxlev)
        return self.engine.ring_image(self.fftwidth // 4, self.scroll, self.rows_seen, "rgba",
                                      levels=(lo, hi), lut=lut)

    def autolevel(self, fix: bool = False):
        """Waterfall.autolevel (S:1668-1680): the 2nd and 98th percentile of the
        image pixels below zero, selected on the device.  Like the reference
        the result lands in ``minlevel`` / ``maxlevel`` and the levels in use
        (``minlev`` / ``maxlev``) are returned unchanged -- the reference's
        attribute-name slip (S:1676-1677); ``fix=True`` applies them."""
        (lo, hi), _ = self.engine.ring_quantiles(self.fftwidth // 4, self.scroll, self.rows_seen, [0.02, 0.98])
        self.minlevel, self.maxlevel = lo, hi
        if fix:
            self.minlev, self.maxlev = lo, hi
        return self.minlev, self.maxlev

    def newlevel(self, low, high):                             # S:1682-1686
        self.minlev, self.maxlev = low, high
        return low, high
