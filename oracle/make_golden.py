"""Record golden vectors from the REFERENCE's own methods -- TEST
INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden

Writes tests/golden/rows.npz (one float64 row per case in
oracle/golden_cases.py), tests/golden/zoomfft.npz (mixed+decimated chunks),
tests/golden/data_trace.npz (Data.add fold-back trace), tests/golden/
waterfall.npz (Waterfall.image_update image) and tests/golden/MANIFEST.json
(library versions).  Inputs are regenerated from seeds by the tests.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import scipy

from oracle import golden_cases as gc
from oracle import ref_harness as rh
from oracle import zoompsd_oracle as zo

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "tests", "golden")


DATA_TRACE_LENS = [16392] * 15 + [16392, 5000, 16392 * 3, 262272, 7]


def data_trace_chunks():
    """Deterministic, compressible chunks: chunk i = (i+1) + 1j*(k mod 97)."""
    return [np.full(n, i + 1.0) + 1j * (np.arange(n) % 97)
            for i, n in enumerate(DATA_TRACE_LENS)]


def reference_row(case):
    x = gc.make_input(case)
    if case.get("u8"):
        x = zo.rtlsdr_bytes_to_iq(x)      # pyrtlsdr restatement (unpinned)
    if case.get("flip"):
        x = np.flip(x)                    # what RTLSDR.Read hands on (T:460)
    if case["path"] == "T":
        return rh.thread_update(x, case["fs"], case["N"], case["R"],
                                case["window"], real=bool(case.get("data_real")))
    return rh.spectrum_update(x, case["fs"], case["N"], case["R"],
                              case["window"], case["n_win"])


def main():
    if not rh.available():
        sys.exit("reference tree not present; golden vectors can only be "
                 "regenerated in the build container")
    os.makedirs(OUT, exist_ok=True)
    rows = {}
    for case in gc.CASES:
        rows[case["name"]] = np.asarray(reference_row(case), dtype=np.float64)
        print("%-24s W=%d  max=%.4f" % (case["name"], len(rows[case["name"]]),
                                       np.max(rows[case["name"]])))
    np.savez_compressed(os.path.join(OUT, "rows.npz"), **rows)

    zf = {}
    for case in gc.ZOOMFFT_CASES:
        x = gc.make_input(case)
        zf[case["name"]] = rh.spectrum_zoomfft(x, case["fs"], case["N"],
                                               case["R"])
    np.savez_compressed(os.path.join(OUT, "zoomfft.npz"), **zf)

    # Data fold-back trace: chunk lengths chosen to wrap the 16-chunk buffer
    lens = DATA_TRACE_LENS
    chunks = data_trace_chunks()
    trace, tail, max_size = rh.thread_data_trace(chunks)
    np.savez_compressed(os.path.join(OUT, "data_trace.npz"), lens=np.array(lens),
                        trace=trace, tail=tail, max_size=max_size)

    # Waterfall image after 70 rows of width 256 (wraps the 64-row image)
    rows_wf = gc.waterfall_rows_ramp()
    img_pos = rh.waterfall_rows(rows_wf, scroll=1)
    img_neg = rh.waterfall_rows(rows_wf, scroll=-1)
    # the reference's own autolevel (S:1668-1680) on those images and on a
    # partly filled image of noise rows
    wf = dict(img_pos=img_pos, img_neg=img_neg)
    for key, rows_a, scroll in (("ramp_pos", rows_wf, 1), ("ramp_neg", rows_wf, -1),
                                ("noise_pos", gc.waterfall_rows_noise(), 1),
                                ("noise_neg", gc.waterfall_rows_noise(), -1)):
        levels, ret = rh.waterfall_autolevel(rows_a, scroll)
        wf["auto_" + key] = levels
        wf["autoret_" + key] = ret
    wf["img_noise_pos"] = rh.waterfall_rows(gc.waterfall_rows_noise(), scroll=1)
    np.savez_compressed(os.path.join(OUT, "waterfall.npz"), **wf)

    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"numpy": np.__version__, "scipy": scipy.__version__,
                   "generator": "python -m oracle.make_golden",
                   "source": "reference methods ApplicationDisplay.update/"
                             "zoomfft, PSD.update, Data.add, "
                             "Waterfall.image_update, Waterfall.autolevel run under Qt stubs "
                             "(oracle/ref_harness.py)",
                   "cases": [c["name"] for c in gc.CASES + gc.ZOOMFFT_CASES]},
                  f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
