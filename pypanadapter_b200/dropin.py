"""Put the GPU path behind an already imported pypanadapter module.

    import pypanadapter_thread as pan          # or pypanadapter_spectrum
    import pypanadapter_b200.dropin as dropin
    dropin.install(pan)                         # before pan.main()

Only the hot-path bodies are replaced; every Qt class, menu, front-end and
the AppState object stay the reference's own:

* ``ApplicationDisplay.zoomfft`` / ``.update(chunk)``  (S:2088-2130)
* ``PSD.update``                                      (T:1513-1549)
* ``Data``'s storage and ``add`` / ``get_data_*``      (T:1400-1483)

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
        # real chunks (AudioPan, S:712-714) without zoom: welch is one-sided (S:2111)
        onesided = np.isrealobj(chunk) and chunk.dtype != np.uint8 and not st.fft_ratio > 1
        eng.configure(fs, st.fft_size, st.fft_ratio, len(chunk), st.fft_tapering, crop=self.N_WIN,
                      onesided=onesided)
        psd = eng.process(chunk)[0].astype(np.float64)
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
            if size >= st.fft_size:
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


def install(module, *, engine: ZoomPSD | None = None, replace_data: bool = True) -> dict:
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
    return saved


def uninstall(module, saved: dict) -> None:
    for key, val in saved.items():
        if "." in key:
            cls, attr = key.split(".")
            setattr(getattr(module, cls), attr, val)
        else:
            setattr(module, key, val)
