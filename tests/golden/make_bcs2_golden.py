#!/usr/bin/env python
"""Generates tests/golden/bcs2_golden.npz: outputs of the REFERENCE's own
kernel::bc::ConductorBoundaries_kernel and AxisBoundaries_kernel (src/kernels/fields_bcs.hpp:
523-868, compiled in place -> oracle/_ref/libref_bcs.so) over the ranges
srpic::PerfectConductorFieldsIn / AxisFieldsIn build, on seeded random 2D fields.

usage: python tests/golden/make_bcs2_golden.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402

N = (19, 13)
NG = 2


def field(seed):
    g = orc.Grid.make(N, NG)
    return g, np.random.default_rng(seed).standard_normal(g.shape(6)).astype(np.float32)


def cases():
    for o in (0, 1):
        for sign in (-1, 1):
            for tags in (1, 2, 3):
                yield f"conductor_o{o}_s{sign}_t{tags}", "conductor", o, sign, tags
    for sign in (-1, 1):
        for tags in (1, 2, 3):
            yield f"axis_s{sign}_t{tags}", "axis", 1, sign, tags


if __name__ == "__main__":
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_bcs.so"))
    f32p = C.POINTER(C.c_float)
    te, tb = lib.ref_bc_tag_e(), lib.ref_bc_tag_b()
    out = {}
    for k, (name, kind, o, sign, tags) in enumerate(cases()):
        g, em = field(1000 + k)
        rt = (te if tags & 1 else 0) | (tb if tags & 2 else 0)
        if kind == "conductor":
            lib.ref_conductor_fields(C.byref(g), em.ctypes.data_as(f32p), o, sign, rt)
        else:
            lib.ref_axis_fields(C.byref(g), em.ctypes.data_as(f32p), sign, rt)
        out[name] = em
    path = os.path.join(ROOT, "tests", "golden", "bcs2_golden.npz")
    np.savez_compressed(path, **out)
    print(f"{len(out)} arrays -> {path}")
