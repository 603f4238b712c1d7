#!/usr/bin/env python
"""Generates tests/golden/mink_golden.npz: outputs of the REFERENCE's own Minkowski kernels
(src/kernels/{pushers/sr,currents_deposit,digital_filter,faraday_mink,ampere_mink}.hpp compiled
in place -> oracle/_ref/libref_o*.so) on the seeded inputs of tests/mink_cases.py.

usage: python tests/golden/make_mink_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import mink_cases as mc  # noqa: E402
from oracle import orc  # noqa: E402

refs = {o: orc.reference(o) for o in range(4)}
assert all(r is not None for r in refs.values()), "oracle/_ref/libref_o*.so not built"
out = mc.run_all(lambda o: refs[o])
path = os.path.join(ROOT, "tests", "golden", "mink_golden.npz")
np.savez_compressed(path, **out)
print(f"{len(out)} arrays -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)")
