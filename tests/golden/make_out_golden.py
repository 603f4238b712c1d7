#!/usr/bin/env python
"""Generates tests/golden/out_golden.npz: outputs of the REFERENCE's own FieldsToPhys_kernel and
PrtlToPhys_kernel (compiled in place -> oracle/_ref/libref_out.so, oracle/ref_out_driver.cpp) on
the seeded inputs of tests/out_cases.py, for Minkowski and the five curvilinear metrics (2D).

usage: python tests/golden/make_out_golden.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import out_cases as oc  # noqa: E402
from oracle import orc  # noqa: E402


class RefMetric(C.Structure):
    _fields_ = [("kind", C.c_int), ("n1", C.c_int), ("n2", C.c_int), ("x1min", C.c_float),
                ("x1max", C.c_float), ("x2min", C.c_float), ("x2max", C.c_float), ("r0", C.c_float),
                ("h", C.c_float), ("a", C.c_float)]


def metric(kind):
    m = RefMetric()
    m.kind, m.n1, m.n2 = kind, oc.N[0], oc.N[1]
    d = oc.METRICS[kind]
    if kind == 0:
        m.x1min, m.x2min = d["xmin"]
        m.x1max = m.x1min + np.float32(d["dx"]) * oc.N[0]
        m.x2max = m.x2min + np.float32(d["dx"]) * oc.N[1]
    else:
        m.x1min, m.x1max, m.x2min, m.x2max, m.r0, m.h, m.a = d["mp"]
    return m


if __name__ == "__main__":
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_out.so"))
    f32p = C.POINTER(C.c_float)
    interp_flag = {0: 0, 1: lib.refo_flag(0), 2: lib.refo_flag(1)}
    conv_flag = {0: 0, 1: lib.refo_flag(2), 2: lib.refo_flag(3), 3: lib.refo_flag(4)}
    out = {}
    g = orc.Grid.make(oc.N, oc.NG)
    for kind in oc.METRICS:
        m = metric(kind)
        for k, (interp, conv, cf, ct) in enumerate(oc.FIELD_CASES):
            src = oc.field(kind)
            dst = np.zeros_like(src)
            lib.refo_fields_to_phys(C.byref(m), C.byref(g), src.ctypes.data_as(f32p),
                                    dst.ctypes.data_as(f32p), (C.c_int * 3)(*cf), (C.c_int * 3)(*ct),
                                    interp_flag[interp] | conv_flag[conv])
            out[f"fld_m{kind}_c{k}"] = dst
        p = oc.particles(kind)
        ps = orc.ParticleSet(oc.NPART)
        for nm, v in p.items():
            getattr(ps, nm)[:] = v
        nout = (oc.NPART + oc.STRIDE - 1) // oc.STRIDE
        bufs = [np.zeros(nout, np.float32) for _ in range(7)]
        st = ps.struct()
        lib.refo_prtls_to_phys(C.byref(m), C.byref(st), oc.NPART, oc.STRIDE, nout,
                               *[b.ctypes.data_as(f32p) for b in bufs])
        out[f"prt_m{kind}"] = np.stack(bufs)
    path = os.path.join(ROOT, "tests", "golden", "out_golden.npz")
    np.savez_compressed(path, **out)
    print(f"{len(out)} arrays -> {path} ({os.path.getsize(path) / 1e3:.0f} kB)")
