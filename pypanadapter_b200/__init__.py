"""pypanadapter_b200 -- B200-native zoom-FFT PSD path behind pypanadapter's call surface.

The arithmetic lives in ``libzoomfft_b200.so`` (hand-written sm_100a CUDA, C ABI in
``include/zoomfft_b200.h``); there is no CPU fallback.  Importing this package does not
load the library -- the first engine does, and fails loudly if it is missing.
"""
from .engine import (ZoomFFTError, ZoomPSD, crop_width, decim_sos, default_engine,  # noqa: F401
                     plan_geometry, window_table, zoom_psd)

__all__ = ["ZoomPSD", "ZoomFFTError", "zoom_psd", "default_engine", "plan_geometry",
           "crop_width", "window_table", "decim_sos"]
