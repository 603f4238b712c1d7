"""Seeded inputs of the matching-boundary parity cases (tests/golden/bcs_golden.npz): shared by
the generator (the compiled reference, oracle/_ref/libref_bcs.so) and the tests. The
MatchFields functor of these cases is the polynomial field setter of oracle/ref_bcs_driver.cpp:
value_c(x) = a[c][0] + a[c][1] x1 + a[c][2] x2 + a[c][3] x3, evaluated left to right in fp32."""
import numpy as np

from helpers import smooth_fields
from oracle import orc

DX = 0.5
G = 2
GRIDS = {1: (64,), 2: (40, 32), 3: (14, 12, 10)}
XMIN = {1: (-3.0,), 2: (1.0, -8.0), 3: (0.5, 2.0, -1.5)}
COEF = np.array([[0.30, 0.10, -0.20, 0.05], [-0.10, 0.02, 0.07, -0.03], [0.20, -0.04, 0.01, 0.06],
                 [1.00, 0.03, -0.05, 0.02], [-0.50, 0.08, 0.04, -0.01], [0.70, -0.06, 0.09, 0.03]],
                dtype=np.float32)
BC_E, BC_B = 1, 2


def grid(dim):
    return orc.Grid.make(GRIDS[dim], G)


def cases():
    """(name, dim, o, sign, ncell_ds, tags, b_only)"""
    out = []
    for dim in (1, 2, 3):
        for o in range(dim):
            for sign in (-1, +1):
                out.append((f"{dim}d_o{o}_{'p' if sign > 0 else 'm'}_EB", dim, o, sign, 5, BC_E | BC_B, False))
    out.append(("2d_o1_p_E", 2, 1, +1, 7, BC_E, False))
    out.append(("2d_o1_m_B", 2, 1, -1, 7, BC_B, False))
    out.append(("2d_o0_p_bonly", 2, 0, +1, 6, BC_E | BC_B, True))
    out.append(("3d_o2_m_bonly", 3, 2, -1, 4, BC_E | BC_B, True))
    return out


def setup(dim, o, sign, ncell_ds):
    """fields, matching geometry and the index range srpic::MatchFieldsIn would pass
    (src/engines/srpic/fields_bcs.h:72-114): the layer [xg_min, xg_max] of thickness ds at the
    domain edge, ghosts included on the outer side and over the full transverse extent"""
    g = grid(dim)
    n = GRIDS[dim]
    em = smooth_fields(g, 700 + 10 * dim + o, amp=0.9)
    ds = np.float32(ncell_ds * DX)
    xmin_o = np.float32(XMIN[dim][o])
    xmax_o = np.float32(xmin_o + np.float32(DX) * np.float32(n[o]))
    rmin, rmax = [0] * dim, [n[a] + 2 * G for a in range(dim)]
    if sign > 0:
        xg_edge = xmax_o
        rmin[o] = G + n[o] - ncell_ds
    else:
        xg_edge = xmin_o
        rmax[o] = G + ncell_ds
    return g, em, float(xg_edge), float(ds), rmin, rmax


def target(dim, mask_b_only=False):
    """the polynomial setter on every component's own node (tetrad basis), fp32, in the
    operation order of the reference: x = (i_ + stag) * dx + xmin; v = a0 + a1 x1 + ..."""
    n = GRIDS[dim]
    ext = [n[a] + 2 * G for a in range(dim)]
    t = np.zeros((6,) + tuple(reversed(ext)), np.float32)
    for c in range(6):
        is_b, a = c >= 3, c % 3
        xs = []
        for d in range(dim):
            stag = (d != a) if is_b else (d == a)
            i_ = (np.arange(ext[d], dtype=np.float32) - np.float32(G))
            xi = (i_ + np.float32(0.5)) if stag else i_
            xs.append((xi * np.float32(DX) + np.float32(XMIN[dim][d])).astype(np.float32))
        v = np.full(tuple(reversed(ext)), COEF[c][0], np.float32)
        for d in range(dim):
            shape = [1] * dim
            shape[dim - 1 - d] = ext[d]
            v = (v + (COEF[c][1 + d] * xs[d]).astype(np.float32).reshape(shape)).astype(np.float32)
        t[c] = v
    return t
