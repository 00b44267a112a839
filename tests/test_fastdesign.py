"""Host-side FIR plan of mode fast (pypanadapter_b200/fastdesign.py): the design
meets its own specification for every zoom ratio the UI offers (S:2079-2086:
up to 512), in fp64, against the reference's cheby1 response -- no GPU needed."""
import ctypes as C

import numpy as np
import pytest

from oracle import zoompsd_oracle as zo
from pypanadapter_b200 import _lib, fastdesign as fd


@pytest.mark.parametrize("R", [4, 8, 16, 32, 64, 128, 256, 512])
def test_plan_meets_spec(R):
    sos = zo.decim_sos()
    plan = fd.design(R, sos)
    k = int(np.log2(R))
    assert len(plan["stages"]) == k - 1
    f_last = 2.0 / R
    B = fd.BETA * f_last
    total = np.ones(400)
    fg = np.linspace(0, B, 400)
    for s, a in enumerate(plan["stages"]):
        rate = 1.0 / 2 ** s
        b = B / rate
        assert len(a) - 1 <= fd.MAX_HALF
        assert abs(a[0] + 2 * a[1:].sum() - 1.0) < 1e-12                     # unit DC gain
        gp = fd.response(a, np.linspace(0, b, 300))
        gs = fd.response(a, np.linspace(0.5 - b, 0.5, 600))
        assert gp.min() > 0
        assert 20 * np.log10(np.abs(gs).max() / gp.min()) <= -fd.REJECT_DB + 1e-6   # alias rejection
        assert 20 * np.log10(gp.max() / gp.min()) <= fd.MAX_DROOP_DB + 1e-6
        total *= fd.response(a, fg / rate)
    # compensator restores the product of the reference's zero-phase gains on |f| <= B
    want = np.ones_like(fg)
    for s in range(k - 1):
        want *= fd.cheby_power_gain(sos, fg * 2 ** s)
    got = total * fd.response(plan["comp"], fg / f_last)
    assert np.abs(got / want - 1).max() < 2 * fd.FIT_TOL
    assert len(plan["comp"]) - 1 <= fd.MAX_COMP_HALF
    assert plan["strip"] >= 64 and plan["strip"] % 2 == 0


def test_last_stage_rejects_beyond_protected_band():
    """Why beta = 0.35: the exact last stage (cheby1 order 8, twice) attenuates
    everything beyond 0.35 of ITS input rate by >= 178 dB."""
    sos = zo.decim_sos()
    f = np.linspace(0.35, 0.5, 400)
    # cheby_power_gain = |H|^2 = the AMPLITUDE gain of the zero-phase (forward + backward) pass
    assert 20 * np.log10(fd.cheby_power_gain(sos, f).max()) <= -178.0
    # and the pass band is the reference's: <= 2 x 0.05 dB ripple up to 0.2 of the rate
    p = 20 * np.log10(fd.cheby_power_gain(sos, np.linspace(0, 0.2, 400)))
    assert p.max() <= 1e-9 and p.min() >= -0.1001


def test_cheby_power_gain_matches_scipy():
    import scipy.signal
    sos = zo.decim_sos()
    f = np.linspace(0, 0.5, 257)
    _, h = scipy.signal.sosfreqz(sos, worN=2 * np.pi * f)
    assert np.allclose(fd.cheby_power_gain(sos, f), np.abs(h) ** 2, rtol=1e-10, atol=1e-300)


def test_fp64_chain_matches_reference_interior():
    """The designed chain (numpy, fp64) against scipy's decimate cascade away
    from the chunk edges: the design residual, free of fp32 effects."""
    R, n = 16, 1 << 17
    rng = np.random.default_rng(3)
    k = np.arange(n)
    x = 0.5 * np.exp(2j * np.pi * 0.0007 * k) + 0.3 * np.exp(-2j * np.pi * 0.31 * k) \
        + 1e-2 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    plan = fd.design(R, zo.decim_sos())
    y = x
    for a in plan["stages"]:
        taps = np.concatenate([a[:0:-1], a])
        M = len(a) - 1
        y = np.convolve(y, taps)[M:M + len(y)][::2]
    c = np.concatenate([plan["comp"][:0:-1], plan["comp"]])
    Mc = len(plan["comp"]) - 1
    y = np.convolve(y, c)[Mc:Mc + len(y)]
    y = zo.decimate2_ref(y)
    ref = x
    for _ in range(4):
        ref = zo.decimate2_ref(ref)
    K = plan["strip"]
    err = np.abs(y - ref)[K:-K].max() / np.abs(ref).max()
    assert err < 2e-6


def test_engine_rejects_bad_plans():
    """zfb_set_fast_plan validates what it is handed (host-only call)."""
    import os
    import shutil
    from pypanadapter_b200 import build
    if not os.path.isfile(_lib.LIB_PATH):
        if shutil.which("nvcc") is None and not os.path.isfile("/usr/local/cuda/bin/nvcc"):
            pytest.skip("no nvcc and no prebuilt library")
        build.build()
    lib = _lib.product_library()
    # no engine without a GPU: only the NULL-engine path can be exercised here
    assert lib.zfb_set_fast_plan(None, None) == _lib.ZFB_EINVAL
    assert lib.zfb_fast_active(None) == _lib.ZFB_EINVAL
    assert C.sizeof(_lib.ZfbFastPlan) == 4 + 4 * 12 + 4 + 8 * 12 + 4 + 4 + 8 + 8   # with padding: 176
