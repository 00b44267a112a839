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


class RtlTcpPan:
    """rtl_tcp client behind the reference's ``PanStreamClass`` surface
    (pypanadapter_spectrum.py:271-297: ``Mode == 'Stream'``, ``Stream(cb,
    chunk_size)``, ``stream_open/stream_close``, ``SetFrequency``, ``Close``).

    rtl_tcp's wire protocol (osmocom rtl-sdr, ``rtl_tcp.c``): the server greets
    with 12 bytes -- magic ``RTL0``, tuner type and gain count as big-endian
    uint32 -- and then sends the dongle's interleaved uint8 I,Q without any
    framing; the client sends 5-byte commands (1 byte id, big-endian uint32
    argument: 0x01 centre frequency, 0x02 sample rate, 0x09 direct sampling).
    The callback receives RAW BYTES (``chunk_size`` samples = 2*chunk_size
    uint8), ready for ``buffers.Data.new_u8().add`` -- conversion (S:543) and
    flip happen on the device.  ``complex_callback=True`` instead hands over
    what ``RTLSDRstream.read_callback`` emits (S:459-460): flipped complex128.
    """
    Mode = "Stream"
    _name = "rtl_tcp"

    def __init__(self, host="127.0.0.1", port=1234, sample_rate=2.56e6, complex_callback=False, timeout=5.0):
        import socket
        self.SampleRate = float(sample_rate)          # S:407
        self.complex_callback = complex_callback
        self.stream = None
        self.chunk_size = 0
        self.update_signal = None
        self._sock = socket.create_connection((host, port), timeout=timeout)
        hdr = self._recv_exact(12)
        if hdr[:4] != b"RTL0":
            self._sock.close()
            raise ConnectionError("not an rtl_tcp server (greeting %r)" % hdr[:4])
        self.tuner_type = int.from_bytes(hdr[4:8], "big")
        self.gain_count = int.from_bytes(hdr[8:12], "big")
        self.driver = self
        self._name = "rtl_tcp @%s:%d" % (host, port)
        self._command(0x02, int(self.SampleRate))
        # the connect / greeting timeout must not apply to the sample stream: a source that
        # pauses longer than it would end the pump silently
        self._sock.settimeout(None)
        self.stream_error = None

    @property
    def name(self):
        return self._name

    def _recv_exact(self, n):
        buf = bytearray()
        while len(buf) < n:
            part = self._sock.recv(n - len(buf))
            if not part:
                raise EOFError("rtl_tcp server closed the connection")
            buf += part
        return bytes(buf)

    def _command(self, cmd, arg):
        self._sock.sendall(bytes([cmd]) + int(arg).to_bytes(4, "big"))

    # -- PanStreamClass surface ------------------------------------------------
    def SetFrequency(self, IF):                       # S:462-476
        self._command(0x09, 2 if IF < 30.e6 else 0)   # direct sampling below 30 MHz (S:531-537)
        self._command(0x01, int(IF))
        self.center_freq = IF

    def Stream(self, update_signal, chunk_size):      # S:290-297
        self.update_signal = update_signal
        self.chunk_size = int(chunk_size)
        if self.stream:
            self.stream_close()
        self.stream = self.stream_open()

    def stream_open(self):
        import threading
        stop = threading.Event()                      # one per pump: a late pump never restarts
        self._stop = stop
        emit = getattr(self.update_signal, "emit", self.update_signal)
        chunk_bytes = 2 * self.chunk_size

        def pump():
            try:
                while not stop.is_set():
                    raw = np.frombuffer(self._recv_exact(chunk_bytes), dtype=np.uint8)
                    if stop.is_set():
                        break
                    if self.complex_callback:
                        iq = raw.astype(np.float64).view(np.complex128) / 127.5 - (1 + 1j)
                        emit(np.flip(iq))
                    else:
                        emit(raw)
            except (EOFError, OSError) as exc:
                if not stop.is_set():
                    self.stream_error = exc           # the source went away: visible to the owner
        t = threading.Thread(target=pump, daemon=True)
        t.start()
        return t

    def stream_close(self):
        """Stops the pump for good: the socket's read side is shut down so that a
        pump blocked in recv returns at once -- two pumps must never read the same
        socket (a partial chunk taken by the old one would swap I and Q for the new)."""
        import socket
        stop = getattr(self, "_stop", None)
        if stop is not None:
            stop.set()
        t, self.stream = self.stream, None
        if t is not None and t.is_alive():
            try:
                self._sock.shutdown(socket.SHUT_RD)
            except OSError:
                pass
            t.join()

    def Close(self):
        self.stream_close()
        try:
            self._sock.close()
        except OSError:
            pass
        self.driver = None
