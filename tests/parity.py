"""Shared parity helpers: run a golden case through an engine and compare with
the recorded reference row under BASELINE.json's tolerance:

* bin ordering / argmax peak bin: exact;
* dB20 values: |diff| <= 0.01 for every bin above -100 dBFS, where -100 dBFS is
  read in the strictest way: 200 dB20 below the reading of a
  full-scale bin-centred tone for that configuration.
"""
from __future__ import annotations

import os

import numpy as np

from oracle import golden_cases as gc
from oracle import zoompsd_oracle as zo

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DB_TOL = 0.01            # dB20, BASELINE.json north_star
FLOOR_BELOW_FS = 200.0   # dB20 == 100 true dB

_rows = None


def golden_rows():
    global _rows
    if _rows is None:
        _rows = dict(np.load(os.path.join(GOLDEN_DIR, "rows.npz")))
    return _rows


def floor_db20(fs, window, nperseg, zoomed):
    return zo.tone_peak_db20(1.0, fs, window, nperseg, zoomed) - FLOOR_BELOW_FS


def assert_row_parity(row, ref, floor, what=""):
    row = np.asarray(row, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert row.shape == ref.shape, "%s: row width %r != reference %r" % (what, row.shape, ref.shape)
    mask = ref > floor
    if mask.any():
        # peak bin bit-exact (compare on the reference's own argmax; -inf rows have no peak)
        assert int(np.argmax(row)) == int(np.argmax(ref)), "%s: argmax differs" % what
        d = np.abs(row[mask] - ref[mask])
        assert np.all(np.isfinite(row[mask])), "%s: non-finite value above the floor" % what
        assert d.max() <= DB_TOL, "%s: max |diff| %.5f dB20 > %.2f over %d bins" % (
            what, d.max(), DB_TOL, int(mask.sum()))
    # bins the reference reports as -inf (exact zero power) must not come back large
    neg_inf = np.isneginf(ref)
    if neg_inf.any():
        assert np.all(row[neg_inf] < floor), "%s: reference -inf bins came back above the floor" % what
    return float(np.abs(row[mask] - ref[mask]).max()) if mask.any() else 0.0


def case_config(case):
    crop = "thread" if case["path"] == "T" else case["n_win"]
    dtype = "u8" if case.get("u8") else "c64"
    return crop, dtype


def run_case(engine, case, x=None, mode="exact"):
    """One golden case through an engine's HOST path -> float64 row."""
    if x is None:
        x = gc.make_input(case)
    crop, dtype = case_config(case)
    n = len(x) // 2 if dtype == "u8" else len(x)
    # welch of real samples is one-sided (S:2111; T:1538 only when Data holds reals)
    onesided = dtype == "c64" and np.isrealobj(x) and case["R"] == 1 and \
        (case["path"] == "S" or bool(case.get("data_real")))
    engine.configure(case["fs"], case["N"], case["R"], n, case["window"], dtype=dtype,
                     flip=bool(case.get("flip")), crop=crop, mode=mode, onesided=onesided)
    return engine.process(x)[0].astype(np.float64)


def check_case(engine, name, mode="exact"):
    case = gc.case_by_name(name)
    row = run_case(engine, case, mode=mode)
    floor = floor_db20(case["fs"], case["window"], engine.geometry["nperseg"], case["R"] > 1)
    return assert_row_parity(row, golden_rows()[name], floor, name)


def oracle_row(case, x=None):
    if x is None:
        x = gc.make_input(case)
    crop, _ = case_config(case)
    if np.isrealobj(x) and x.dtype != np.uint8 and case["path"] == "T" and not case.get("data_real"):
        x = x.astype(np.complex128)              # the reference's Data buffer is complex (T:1807)
    return zo.zoom_psd(x, case["fs"], case["N"], case["R"], case["window"], crop=crop,
                       flip=bool(case.get("flip")))
