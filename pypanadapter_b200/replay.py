"""Replay of recorded RTL-SDR IQ (uint8 offset binary, the rtl_sdr / rtl_tcp
wire format) behind the reference's sample-source interface -- BASELINE
configs[1] without a dongle.

``ReplayPan`` has the surface of the reference's ``PanBlockClass`` sources
(pypanadapter_spectrum.py:231-305, ``RTLSDR`` S:493-547): ``Mode == 'Block'``,
``SampleRate``, ``name``, ``driver``, ``Read(size)``, ``SetFrequency(IF)``,
``Close()``.  ``Read`` hands back what ``RTLSDR.Read`` does -- complex128,
pyrtlsdr's ``u/127.5 - (1+1j)``, ``np.flip``-ped (S:541-543) -- so the
reference's ``DataReader`` / block timer work unchanged.  ``ReadRaw`` returns
the bytes themselves for ``buffers.Data.new_u8()`` (conversion + flip then
happen on the device).
"""
from __future__ import annotations

import numpy as np


class ReplayPan:
    Mode = "Block"
    _name = "Replay"

    def __init__(self, source, sample_rate=3.2e6, loop=True, name=None):
        """``source``: path of a .bin/.cu8 file of interleaved uint8 I,Q, or a
        uint8 ndarray."""
        if isinstance(source, (str, bytes)):
            self._raw = np.memmap(source, dtype=np.uint8, mode="r")
            self._name = "Replay %s" % source
        else:
            self._raw = np.ascontiguousarray(source, dtype=np.uint8)
        if name:
            self._name = name
        if len(self._raw) < 2 or len(self._raw) % 2:
            raise ValueError("uint8 IQ recording must hold an even, non-zero number of bytes")
        self.SampleRate = float(sample_rate)
        self.loop = loop
        self.driver = self            # the reference only tests it for truth / center_freq
        self.center_freq = 0.0
        self._pos = 0                 # in samples

    @property
    def name(self):
        return self._name

    # -- PanBlockClass surface ------------------------------------------------
    def SetFrequency(self, IF):
        self.center_freq = IF         # tuning is in the recording

    def ReadRaw(self, size) -> np.ndarray:
        """``size`` samples as 2*size interleaved uint8 (wire format)."""
        size = int(size)
        total = len(self._raw) // 2
        out = np.empty(2 * size, dtype=np.uint8)
        done = 0
        while done < size:
            if self._pos >= total:
                if not self.loop:
                    raise EOFError("end of recording")
                self._pos = 0
            n = min(size - done, total - self._pos)
            out[2 * done:2 * (done + n)] = self._raw[2 * self._pos:2 * (self._pos + n)]
            done += n
            self._pos += n
        return out

    def Read(self, size) -> np.ndarray:
        """What ``RTLSDR.Read`` returns (S:541-543): flipped complex128."""
        raw = self.ReadRaw(size)
        iq = raw.astype(np.float64).view(np.complex128)
        iq = iq / 127.5 - (1 + 1j)
        return np.flip(iq)

    def Close(self):
        self.driver = None
