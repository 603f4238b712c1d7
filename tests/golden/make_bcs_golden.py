#!/usr/bin/env python
"""Generates tests/golden/bcs_golden.npz: outputs of the REFERENCE's own
kernel::bc::MatchBoundaries_kernel (src/kernels/fields_bcs.hpp compiled in place ->
oracle/_ref/libref_bcs.so, oracle/ref_bcs_driver.cpp) on the seeded inputs of tests/bcs_cases.py.

usage: python tests/golden/make_bcs_golden.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bcs_cases as bc  # noqa: E402
from oracle import orc  # noqa: E402


def load():
    path = os.path.join(ROOT, "oracle", "_ref", "libref_bcs.so")
    assert os.path.exists(path), "oracle/_ref/libref_bcs.so not built (make -C oracle ref)"
    lib = C.CDLL(path)
    f32p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int)
    lib.ref_match_fields.argtypes = [C.POINTER(orc.Grid), f32p, f32p, C.c_int, C.c_float, f32p,
                                     C.c_int, C.c_float, C.c_float, C.c_int, i32p, i32p]
    lib.ref_match_fields.restype = None
    return lib


def run_all(lib):
    out = {}
    f32p = C.POINTER(C.c_float)
    te, tb = lib.ref_bc_tag_e(), lib.ref_bc_tag_b()
    for name, dim, o, sign, nds, tags, b_only in bc.cases():
        g, em, xg_edge, ds, rmin, rmax = bc.setup(dim, o, sign, nds)
        xmin = (C.c_float * 3)(*(list(bc.XMIN[dim]) + [0.0] * (3 - dim)))
        lo, hi = (C.c_int * 3)(*(rmin + [0] * (3 - dim))), (C.c_int * 3)(*(rmax + [1] * (3 - dim)))
        coef = np.ascontiguousarray(bc.COEF)
        rt = (te if tags & bc.BC_E else 0) | (tb if tags & bc.BC_B else 0)
        lib.ref_match_fields(C.byref(g), em.ctypes.data_as(f32p), coef.ctypes.data_as(f32p),
                             int(b_only), bc.DX, xmin, o, xg_edge, ds, rt, lo, hi)
        out[name] = em
    return out


if __name__ == "__main__":
    out = run_all(load())
    path = os.path.join(ROOT, "tests", "golden", "bcs_golden.npz")
    np.savez_compressed(path, **out)
    print(f"{len(out)} arrays -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)")
