#!/usr/bin/env python
"""Generates tests/golden/match_layer_golden.npz: the MATCH layer geometry (edge, index range,
intersects or not) from the REFERENCE's own Mesh::Intersects / Mesh::ExtentToRange
(src/framework/domain/mesh.h compiled in place -> oracle/_ref/libref_mesh.so) for seeded domains.

usage: python tests/golden/make_match_layer_golden.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402


def cases(seed=4321, n=400):
    """rows: dim, n[3], dx, local xmin[3], global lo/hi of dimension o, o, sign, ds. A domain is
    a block of a decomposed global box, so that layers miss it, clip it or cover it."""
    rng = np.random.default_rng(seed)
    rows = []
    for _ in range(n):
        dim = int(rng.integers(1, 4))
        nn = [int(rng.integers(4, 40)) if a < dim else 1 for a in range(3)]
        dx = float(np.float32(rng.choice([0.25, 0.5, 1.0, 0.1, 0.3])))
        gmin = [float(np.float32(rng.uniform(-20, 20))) for _ in range(3)]
        nblk = [int(rng.integers(1, 4)) for _ in range(3)]
        blk = [int(rng.integers(0, nblk[a])) for a in range(3)]
        lmin = [float(np.float32(gmin[a]) + np.float32(dx) * np.float32(nn[a] * blk[a])) for a in range(3)]
        lmax = [float(np.float32(lmin[a]) + np.float32(dx) * np.float32(nn[a])) for a in range(3)]
        o = int(rng.integers(0, dim))
        ghi = float(np.float32(gmin[o]) + np.float32(dx) * np.float32(nn[o] * nblk[o]))
        sign = int(rng.choice([-1, 1]))
        ds = float(np.float32(rng.choice([dx * int(rng.integers(1, 30)), rng.uniform(0.1, 25.0)])))
        rows.append((dim, nn, dx, lmin, lmax, gmin[o], ghi, o, sign, ds))
    return rows


def load():
    path = os.path.join(ROOT, "oracle", "_ref", "libref_mesh.so")
    assert os.path.exists(path), "oracle/_ref/libref_mesh.so not built (make -C oracle ref)"
    lib = C.CDLL(path)
    f32p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int)
    lib.ref_match_layer.argtypes = [C.POINTER(orc.Grid), f32p, f32p, C.c_float, C.c_float, C.c_int,
                                    C.c_int, C.c_float, f32p, i32p, i32p]
    lib.ref_match_layer.restype = C.c_int
    return lib


def run_all(lib):
    out = []
    for dim, nn, dx, lmin, lmax, glo, ghi, o, sign, ds in cases():
        g = orc.Grid.make(nn[:dim], 2)
        lo, hi = (C.c_float * 3)(*lmin), (C.c_float * 3)(*lmax)
        edge = C.c_float(0)
        rmin, rmax = (C.c_int * 3)(0, 0, 0), (C.c_int * 3)(0, 0, 0)
        rc = lib.ref_match_layer(C.byref(g), lo, hi, glo, ghi, o, sign, ds, C.byref(edge), rmin, rmax)
        out.append([rc, edge.value] + list(rmin)[:3] + list(rmax)[:3])
    return np.array(out, dtype=np.float64)


if __name__ == "__main__":
    res = run_all(load())
    path = os.path.join(ROOT, "tests", "golden", "match_layer_golden.npz")
    np.savez_compressed(path, results=res)
    print(f"{len(res)} cases ({int(res[:, 0].sum())} intersecting) -> {path}")
