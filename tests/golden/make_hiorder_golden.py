#!/usr/bin/env python
"""Generates tests/golden/hiorder_golden.npz: the REFERENCE's pusher (field interpolation with
prtl_shape::order<.., O>) and Esirkepov deposit for SHAPE_ORDER = 4..11, compiled in place
(oracle/_ref/libref_o{4..11}.so: `make -C oracle ref`), on the seeded inputs of
tests/hiorder_cases.py.

usage: python tests/golden/make_hiorder_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import hiorder_cases as hc  # noqa: E402
from oracle import orc  # noqa: E402

refs = {o: orc.reference(o) for o in hc.ORDERS}
assert all(r is not None for r in refs.values()), "oracle/_ref/libref_o{4..11}.so not built"
out = hc.run_all(lambda o: refs[o])
path = os.path.join(ROOT, "tests", "golden", "hiorder_golden.npz")
np.savez_compressed(path, **out)
print(f"{len(out)} arrays -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)")
