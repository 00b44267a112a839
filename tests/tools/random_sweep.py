#!/usr/bin/env python
"""Randomised configuration sweep on the GPU beyond the seeds in the test
suite (tests/engine_suite.random_configs: frame length, N, R, window, wire
dtype, flip, crop, f_demod, decimator mode) against the oracle.

    python tests/tools/random_sweep.py
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pypanadapter_b200.engine import ZoomPSD
from tests import engine_suite as es
eng = ZoomPSD(0)
fails = 0
for seed in range(300, 340):
    try:
        es.random_configs(eng, seed, 20)
    except AssertionError as exc:
        fails += 1
        print("seed", seed, "FAIL", str(exc)[:300])
print("random sweep (late-mix build): %d configurations, failing seeds: %d" % (40 * 20, fails))
