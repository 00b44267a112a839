"""Host-side (fp64, numpy) design of the ZFB_MODE_FAST FIR plan.

For a zoom ratio R = 2^k the reference runs k zero-phase cheby1 stages
(pypanadapter_spectrum.py:2096-2098 -> scipy.signal.decimate).  FAST keeps the
last one exact and replaces the first k-1 in the chunk interior:

* stage s (input rate fs/2^s): symmetric FIR G_s with unit DC gain whose only
  hard job is alias rejection -- >= ``reject_db`` over the band
  [rate/2 - B, rate/2] that folds onto the protected band |f| <= B, with
  B = beta * 2fs/R (beta = 0.35: beyond it the exact last stage attenuates by
  >= 178 dB).  Minimum stop-band energy subject to G(0) = 1, with a soft pull
  of the pass band to 1 that keeps the droop below ``max_droop_db``.
* compensator C at rate 2fs/R: least-squares symmetric FIR with
  C(f) * prod G_s(f) = prod |H_s(f)|^2 on |f| <= B (H_s = the reference's
  cheby1(8, 0.05 dB, 0.4) at stage s's rate), to ``fit_tol`` relative.

Nothing here runs per frame; plans are cached per (R, parameters).
"""
from __future__ import annotations

import functools

import numpy as np

BETA = 0.35
REJECT_DB = 140.0
MAX_DROOP_DB = 6.0
FIT_TOL = 1e-5
STRIP = 128          # decimated samples recomputed exactly at either chunk end
MAX_HALF = 20        # zfb_firchain.cuh FIR_MAX_HALF
MAX_COMP_HALF = 24   # FIR_COMP_MAX_HALF


def _cos_basis(f, M):
    C = np.ones((len(f), M + 1))
    if M:
        C[:, 1:] = 2.0 * np.cos(2.0 * np.pi * np.outer(f, np.arange(1, M + 1)))
    return C


def response(half_taps, f):
    """Zero-phase response of the symmetric FIR given centre-first half taps."""
    half_taps = np.asarray(half_taps, dtype=np.float64)
    return _cos_basis(np.atleast_1d(np.asarray(f, dtype=np.float64)), len(half_taps) - 1) @ half_taps


def cheby_power_gain(sos, f):
    """|H(e^{j 2 pi f})|^2 of an SOS cascade = gain of one zero-phase pass."""
    z = np.exp(-2j * np.pi * np.atleast_1d(np.asarray(f, dtype=np.float64)))
    h = np.ones_like(z)
    for b0, b1, b2, a0, a1, a2 in np.asarray(sos, dtype=np.float64):
        h = h * (b0 + b1 * z + b2 * z * z) / (a0 + a1 * z + a2 * z * z)
    return np.abs(h) ** 2


def design_stage(b, reject_db=REJECT_DB, max_droop_db=MAX_DROOP_DB):
    """Shortest symmetric FIR (centre-first half taps) with unit DC gain,
    >= reject_db rejection on [0.5-b, 0.5] relative to its smallest gain on
    [0, b], and at most max_droop_db of pass-band droop."""
    fs_ = np.linspace(0.5 - b, 0.5, 1200)
    fp = np.linspace(0.0, b, 600)
    for M in range(1, MAX_HALF + 1):
        Cs, Cp = _cos_basis(fs_, M), _cos_basis(fp, M)
        Ps = Cs.T @ Cs / len(fs_)
        Pp = Cp.T @ Cp / len(fp)
        qp = Cp.T @ np.ones(len(fp)) / len(fp)
        c = np.ones(M + 1)
        c[1:] = 2.0
        for lam in (0.0, 1e-16, 1e-14, 1e-12, 1e-10):
            P = Ps + lam * Pp + 1e-20 * np.eye(M + 1)
            pq = np.linalg.solve(P, lam * qp)
            pc = np.linalg.solve(P, c)
            a = pq + (1.0 - c @ pq) / (c @ pc) * pc          # min a'Pa - 2 lam qp'a  s.t. c'a = 1
            gp, gs = Cp @ a, Cs @ a
            if gp.min() <= 0:
                continue
            rej = 20 * np.log10(np.abs(gs).max() / gp.min())
            droop = 20 * np.log10(gp.max() / gp.min())
            if rej <= -reject_db and droop <= max_droop_db:
                return a, rej, droop
    raise ValueError("no FIR of <= %d taps meets %g dB on b = %g" % (2 * MAX_HALF + 1, reject_db, b))


@functools.lru_cache(maxsize=64)
def _design(R, sos_key, beta, reject_db, max_droop_db, fit_tol):
    sos = np.array(sos_key).reshape(-1, 6)
    k = int(np.log2(R))
    ne = k - 1
    if ne < 1:
        raise ValueError("mode FAST needs fft_ratio >= 4")
    f_last = 2.0 / R                      # input rate of the last (exact) stage, in units of fs
    B = beta * f_last
    stages, report = [], []
    for s in range(ne):
        rate = 1.0 / 2 ** s
        a, rej, droop = design_stage(B / rate, reject_db, max_droop_db)
        stages.append(a)
        report.append(dict(stage=s, taps=2 * (len(a) - 1) + 1, reject_db=float(rej), droop_db=float(droop)))
    fg = np.linspace(0.0, B, 800)
    target = np.ones_like(fg)
    for s in range(ne):
        rate = 1.0 / 2 ** s
        target *= cheby_power_gain(sos, fg / rate) / response(stages[s], fg / rate)
    comp, err = None, None
    for M in range(1, MAX_COMP_HALF + 1):
        C = _cos_basis(fg / f_last, M)
        a, *_ = np.linalg.lstsq(C, target, rcond=None)
        err = float(np.abs(C @ a - target).max() / np.abs(target).min())
        if err < fit_tol:
            comp = a
            break
    if comp is None:
        raise ValueError("compensator does not reach %g (best %g)" % (fit_tol, err))
    return dict(R=R, stages=[np.ascontiguousarray(a) for a in stages], comp=np.ascontiguousarray(comp),
                strip=STRIP, report=report, comp_taps=2 * (len(comp) - 1) + 1, fit_err=err, band=B)


def design(R: int, sos, *, beta=BETA, reject_db=REJECT_DB, max_droop_db=MAX_DROOP_DB, fit_tol=FIT_TOL):
    """FIR plan for zoom ratio R given the reference's SOS (zfb_decim_sos)."""
    key = tuple(np.asarray(sos, dtype=np.float64).ravel().tolist())
    return _design(int(R), key, float(beta), float(reject_db), float(max_droop_db), float(fit_tol))
