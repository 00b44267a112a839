"""Put the GPU path behind an already imported pypanadapter module.

    import pypanadapter_thread as pan          # or pypanadapter_spectrum
    import pypanadapter_b200.dropin as dropin
    dropin.install(pan)                         # before pan.main()

Only the hot-path bodies are replaced; every Qt class, menu, front-end and
the AppState object stay the reference's own:

* ``ApplicationDisplay.zoomfft`` / ``.update(chunk)``  (S:2088-2130)
* ``PSD.update``                                      (T:1513-1549)
* ``Data``'s storage and ``add`` / ``get_data_*``      (T:1400-1483)
* ``FFTTaperingControl.ShowCurve``                    (S:1354-1379): taper table
  and preview spectrum from the device
* ``Waterfall.init_image`` / ``image_update`` / ``autolevel`` / ``newlevel``
  and ``img_array``                                   (S:1625-1686): rows stay
  in the device ring, ``setImage`` receives 8-bit colour indices

``install`` returns a dict of the originals so that ``uninstall`` can put
them back.
"""
from __future__ import annotations

import numpy as np

from . import buffers
from .engine import ZoomPSD, default_engine


def _spectrum_methods(module, engine_of):
    def zoomfft(self, x, ratio=1):                                  # S:2088-2100
        st = module.AppState
        eng = engine_of()
        x = np.asarray(x)
        # the reference's LO has fft_size*fft_avg points (S:2091-2092): same length check
        if len(x) != st.fft_size * st.fft_avg:
            raise ValueError("operands could not be broadcast together with shapes (%d,) (%d,)"
                             % (len(x), st.fft_size * st.fft_avg))
        eng.configure(st.panadapter.SampleRate, st.fft_size, ratio, len(x), st.fft_tapering,
                      crop=None)
        eng.process(x)
        return eng.read_decimated().astype(np.complex128)

    def update(self, chunk):                                        # S:2102-2130
        st = module.AppState
        eng = engine_of()
        fs = st.panadapter.SampleRate
        bw_hz = fs * self.N_WIN / st.fft_size
        self.win.setWindowTitle('PEPYSCOPE - IS0KYB - N_FFT: %d, BW: %.1f kHz'
                                % (st.fft_size, bw_hz / 1000. / st.fft_ratio))
        chunk = np.asarray(chunk)
        # raw wire formats from a source that skips the host conversion (replay.RtlTcpPan's byte
        # callback, a Soapy CS16 stream): converted on the device; RTL bytes are also flipped there,
        # as RTLSDRstream.read_callback flips what it emits (S:459-460)
        if chunk.dtype == np.uint8:
            dtype, n, flip = "u8", len(chunk) // 2, True
        elif chunk.dtype == np.int16:
            dtype, n, flip = "cs16", len(chunk) // 2, False
        else:
            dtype, n, flip = "c64", len(chunk), False
        # real chunks (AudioPan, S:712-714) without zoom: welch is one-sided (S:2111)
        onesided = dtype == "c64" and np.isrealobj(chunk) and not st.fft_ratio > 1
        eng.configure(fs, st.fft_size, st.fft_ratio, n, st.fft_tapering, dtype=dtype, flip=flip,
                      crop=self.N_WIN, onesided=onesided)
        psd = eng.process(chunk)[0].astype(np.float64)
        wf = getattr(self.waterfall, "__dict__", {}).get("_zfb_wf")
        if isinstance(wf, buffers.Waterfall) and wf.engine is eng:
            wf.note_engine_rows(1)                                  # the row is in the device ring already
        self.waterfall.image_update(psd)
        hz = fs / 4
        self.spectrum_plot.setData(np.linspace(-hz, hz, psd.shape[0]), psd, pen="g")

    return zoomfft, update


def _thread_psd_update(module, engine_of):
    def update(self):                                               # T:1513-1549
        st = module.AppState
        d = self.dataclass
        eng = engine_of()
        d.get_data_start()
        size = d.real_size
        row = None
        try:
            if size >= st.fft_size and not getattr(d, "_detached", False):
                if isinstance(d, buffers.Data):
                    eng.configure(st.panadapter.SampleRate, st.fft_size, st.fft_ratio, size,
                                  st.fft_tapering, dtype=d.wire, crop="thread",
                                  onesided=d.real and not st.fft_ratio > 1)
                    row = eng.samples_process()
                else:               # the reference's own Data: snapshot under the lock
                    real = np.isrealobj(d.data)                 # Data.new_real (T:1413-1417)
                    chunk = np.array(d.data[:size], dtype=np.complex64)
                    eng.configure(st.panadapter.SampleRate, st.fft_size, st.fft_ratio, size,
                                  st.fft_tapering, crop="thread", onesided=real and not st.fft_ratio > 1)
                    row = eng.process(chunk)[0]
        finally:
            d.get_data_end()
        if row is None:
            return
        self.lock.lock()
        self.psd = row.astype(np.float64)
        self.lock.unlock()

    return update


_ABSENT = object()


def _waterfall_methods(module, engine_of):
    """Waterfall.init_image / image_update / autolevel / newlevel (S:1625-1686)
    on the device ring: nothing is rolled or redrawn on the host, the item is
    handed the 8-bit colour indices (with levels [0, 256] the table lookup
    pyqtgraph then does is the identity on them)."""

    def mirror(self) -> buffers.Waterfall:
        wf = self.__dict__.get("_zfb_wf")             # (not getattr: Qt base classes may answer anything)
        if wf is None:
            wf = self.__dict__["_zfb_wf"] = buffers.Waterfall(engine_of(), scroll=module.AppState.scroll)
        wf.scroll = module.AppState.scroll
        return wf

    def levels(self):
        return self.__dict__.get("_zfb_levels", (self.minlev, self.maxlev))

    def init_image(self):                                           # S:1625-1635
        st = module.AppState
        bw_hz = st.panadapter.SampleRate / st.fft_size * self.fftwidth / 1.e6 / st.fft_ratio
        self.scale(bw_hz, 1)
        wf = mirror(self)
        wf.fftwidth = self.fftwidth
        wf.init_image()

    def image_update(self, psd):                                    # S:1638-1664
        wf = mirror(self)
        fftwidth = np.size(psd)
        if fftwidth != self.fftwidth:
            self.fftwidth = fftwidth
            self.init_image()
        wf.image_update(psd)                                        # grid bins zeroed in place, row -> ring
        self.setImage(wf.image_indices(levels(self)).T, autoLevels=False, levels=[0, 256],
                      opacity=1.0, autoDownsample=True)

    def autolevel(self):                                            # S:1668-1680
        wf = mirror(self)
        wf.minlev, wf.maxlev = self.minlev, self.maxlev
        wf.autolevel()
        self.minlevel, self.maxlevel = wf.minlevel, wf.maxlevel     # S:1676 (sic)
        self._zfb_levels = (self.minlev, self.maxlev)               # S:1677 setLevels([minlev, maxlev])
        return self.minlev, self.maxlev

    def newlevel(self, low, high):                                  # S:1682-1686
        self._zfb_levels = (low, high)
        return low, high

    img_array = property(lambda self: mirror(self).img_array)
    return dict(init_image=init_image, image_update=image_update, autolevel=autolevel,
                newlevel=newlevel, img_array=img_array)


def _taper_show_curve(module, engine_of):
    """FFTTaperingControl.ShowCurve (S:1354-1379): the dialog's widgets and the
    AppState hand-over stay as they are; the taper table and its preview
    spectrum come from the device (pypanadapter_b200.taper)."""
    from . import taper as _taper

    def ShowCurve(self):
        shape = type(self).taper_list[self.taper]             # [(label, default), ...]
        values = [self.P0val.value(), self.P1val.value()][:len(shape)]
        module.AppState.fft_tapering = (self.taper, *values) if values else self.taper
        cls = type(self)
        curve, spectrum = _taper.show_curve(module.AppState.fft_tapering, getattr(cls, "taper_size", 51),
                                            getattr(cls, "fft_size", 2048), engine=engine_of())
        pen = module.pg.mkPen(color='k', width=2) if hasattr(module, "pg") else None
        for attr, plot, data in (("taperplot", self.plot0, curve), ("fftplot", self.plot1, spectrum)):
            item = getattr(self, attr, None)
            if item:
                item.setData(data)
            else:
                setattr(self, attr, plot.plot(data, pen=pen))

    return ShowCurve


def install(module, *, engine: ZoomPSD | None = None, replace_data: bool = True,
            replace_waterfall: bool = True) -> dict:
    """Patch ``module`` (a loaded pypanadapter_spectrum / pypanadapter_thread)
    in place; returns the originals."""
    engine_of = (lambda: engine) if engine is not None else default_engine
    saved = {}
    app = getattr(module, "ApplicationDisplay", None)
    if app is not None and hasattr(app, "zoomfft"):
        saved["ApplicationDisplay.zoomfft"] = app.zoomfft
        saved["ApplicationDisplay.update"] = app.update
        app.zoomfft, app.update = _spectrum_methods(module, engine_of)
    psd = getattr(module, "PSD", None)
    if psd is not None:
        saved["PSD.update"] = psd.update
        psd.update = _thread_psd_update(module, engine_of)
        if replace_data and hasattr(module, "Data"):
            saved["Data"] = module.Data
            ref_data = module.Data

            class Data(buffers.Data):
                """buffers.Data with the reference's constructor signature
                (T:1402) and its NewtRap pacing controller when available."""

                def __init__(self, chunk_size=8196 * 2):
                    ctrl = None
                    try:
                        ctrl = ref_data(chunk_size).NR          # T:1411
                    except Exception:
                        pass
                    super().__init__(chunk_size, engine=engine_of(), controller=ctrl)

                @buffers.Data.target.setter
                def target(self, t):                            # T:1474-1479
                    if t >= module.AppState.fft_size and t <= self.max_size:
                        self.target_size = t
                        if self.NR is not None:
                            self.NR.target = t

            module.Data = Data
    if not saved:
        raise ValueError("module has neither ApplicationDisplay.zoomfft nor PSD: not a pypanadapter module")
    tap = getattr(module, "FFTTaperingControl", None)
    if tap is not None and hasattr(tap, "ShowCurve"):
        saved["FFTTaperingControl.ShowCurve"] = tap.ShowCurve
        tap.ShowCurve = _taper_show_curve(module, engine_of)
    wfc = getattr(module, "Waterfall", None)
    if replace_waterfall and wfc is not None:
        for name, fn in _waterfall_methods(module, engine_of).items():
            saved["Waterfall." + name] = wfc.__dict__.get(name, _ABSENT)
            setattr(wfc, name, fn)
        if psd is not None:
            # threaded variant: the GUI timer shows whatever row is current (T:2140-2148), not
            # every row PSD.update computes -- only displayed rows enter the ring
            engine_of().set_option("ring_append", 0)
            saved["option.ring_append"] = engine_of()
    return saved


def uninstall(module, saved: dict) -> None:
    for key, val in saved.items():
        if key == "option.ring_append":
            val.set_option("ring_append", 1)
        elif val is _ABSENT:
            cls, attr = key.split(".")
            delattr(getattr(module, cls), attr)
        elif "." in key:
            cls, attr = key.split(".")
            setattr(getattr(module, cls), attr, val)
        else:
            setattr(module, key, val)
