#!/usr/bin/env python
"""Generates tests/golden/stats_golden.npz: outputs of the REFERENCE's own reduced-statistics
kernels (src/kernels/reduced_stats.hpp compiled in place -> oracle/_ref/libref_stats.so, driven
by oracle/ref_stats_driver.cpp in serial order with the reference's fp32 accumulator) on the
seeded inputs of tests/stats_cases.py.

usage: python tests/golden/make_stats_golden.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import stats_cases as sc  # noqa: E402
from oracle import orc  # noqa: E402


def load():
    path = os.path.join(ROOT, "oracle", "_ref", "libref_stats.so")
    assert os.path.exists(path), "oracle/_ref/libref_stats.so not built (make -C oracle ref)"
    lib = C.CDLL(path)
    f32p = C.POINTER(C.c_float)
    lib.ref_stats_fields.argtypes = [C.POINTER(orc.Grid), f32p, f32p, C.c_float, C.c_int, C.c_int]
    lib.ref_stats_fields.restype = C.c_float
    lib.ref_stats_particles.argtypes = [C.POINTER(orc.Grid), C.POINTER(orc.Prtls), C.c_uint32,
                                        C.c_float, C.c_float, C.c_int, C.c_float, C.c_int,
                                        C.c_int, C.c_int]
    lib.ref_stats_particles.restype = C.c_float
    lib.ref_stats_fields_curv.argtypes = [C.c_int, C.POINTER(orc.Grid), f32p, f32p, f32p, C.c_int, C.c_int]
    lib.ref_stats_fields_curv.restype = C.c_float
    lib.ref_stats_particles_curv.argtypes = [C.c_int, C.POINTER(orc.Grid), f32p, C.POINTER(orc.Prtls), C.c_uint32,
                                             C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.ref_stats_particles_curv.restype = C.c_float
    return lib


def run_all(lib):
    out = {}
    f32p = C.POINTER(C.c_float)
    for dim, name, what, comp in sc.field_cases():
        g, em, cur = sc.fields(dim)
        v = lib.ref_stats_fields(C.byref(g), em.ctypes.data_as(f32p), cur.ctypes.data_as(f32p),
                                 sc.DX, what, max(comp, 1))
        out[f"f_{dim}d_{name}_{comp}"] = np.float32(v)
    for dim, k, mass, charge, name, what, use_w, c1, c2 in sc.particle_cases():
        g, p, n = sc.particles(dim, k)
        s = p.struct()
        v = lib.ref_stats_particles(C.byref(g), C.byref(s), n, mass, charge, int(use_w), sc.DX,
                                    what, c1, c2)
        out[f"p_{dim}d_s{k}_{name}_w{int(use_w)}_{c1}{c2}"] = np.float32(v)
    # 2D curvilinear SRPIC meshes: the same kernels instantiated with the reference's own
    # metric::Spherical / metric::QSpherical
    for mname, (kind, ext) in sc.CURV.items():
        e = (C.c_float * 6)(*ext)
        g, em, cur = sc.fields(2)
        for name, what, comp in sc.curv_field_cases():
            v = lib.ref_stats_fields_curv(kind, C.byref(g), e, em.ctypes.data_as(f32p), cur.ctypes.data_as(f32p),
                                          what, max(comp, 1))
            out[f"fc_{mname}_{name}_{comp}"] = np.float32(v)
        for k, mass, charge, name, what, use_w, c1, c2 in sc.curv_particle_cases():
            g, p, n = sc.curv_particles(k)
            s = p.struct()
            v = lib.ref_stats_particles_curv(kind, C.byref(g), e, C.byref(s), n, mass, charge, int(use_w), what,
                                             c1, c2)
            out[f"pc_{mname}_s{k}_{name}_w{int(use_w)}_{c1}{c2}"] = np.float32(v)
    return out


if __name__ == "__main__":
    out = run_all(load())
    path = os.path.join(ROOT, "tests", "golden", "stats_golden.npz")
    np.savez_compressed(path, **out)
    print(f"{len(out)} values -> {path}")
