#!/usr/bin/env python
"""Generates tests/golden/curv_golden.npz: outputs of the REFERENCE's own curvilinear-SR / GR
kernels and metric classes (compiled in place by oracle/Makefile -> oracle/_ref/
libref_curv_o*.so) on the seeded inputs of tests/curv_cases.py. Needs /root/reference at build
time; the tests only need the committed .npz.

usage: python tests/golden/make_curv_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import curv_cases as cc  # noqa: E402

be = cc.RefBackend()
setup = cc.RefSetup(be)
out = cc.run_all(be, setup)
# what the device backend cannot compute itself: time steps and GR momentum scales
for mname in list(cc.SR_METRICS) + list(cc.GR_METRICS):
    out[f"setup/dt/{mname}"] = np.float32(setup.dt(mname))
path = os.path.join(ROOT, "tests", "golden", "curv_golden.npz")
np.savez_compressed(path, **out)
print(f"{len(out)} arrays -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)")
