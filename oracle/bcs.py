"""TEST INFRASTRUCTURE ONLY -- numpy restatement of kernel::bc::MatchBoundaries_kernel for an
SRPIC Minkowski domain (src/kernels/fields_bcs.hpp:42-560): every component defined by the
field setter becomes s F + (1 - s) transform<T->U>(target) with s = tanh(|x_o - xg_edge| 4 / ds)
at the component's staggered node. fp32 throughout, same operation order. Pinned against the
compiled reference: tests/golden/bcs_golden.npz (tests/test_bcs.py); numpy's tanh is not
glibc's tanhf bit for bit, hence the stated tolerance there."""
import ctypes
import ctypes.util

import numpy as np

F32 = np.float32
BC_E, BC_B = 1, 2

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.tanhf.restype = ctypes.c_float
_libm.tanhf.argtypes = [ctypes.c_float]


def tanhf(x):
    """glibc's tanhf element by element: what the reference's host build calls (math::tanh on
    fp32), so that the restatement is bit-exact against a reference built with the same libm
    (numpy's float32 tanh differs in the last ulp). Only ever applied to 1D profiles."""
    x = np.asarray(x, dtype=F32)
    out = np.empty(x.shape, F32)
    xf, of = x.ravel(), out.ravel()
    for k in range(xf.size):
        of[k] = _libm.tanhf(float(xf[k]))
    return out


def match_fields(g, em, target, o, dx, xmin_o, xg_edge, ds, tags, mask, rmin, rmax):
    D, G = g.dim, g.ng
    dx = F32(dx)
    sl = tuple(slice(rmin[d], rmax[d]) for d in reversed(range(D)))
    idx = np.arange(rmin[o], rmax[o], dtype=F32) - F32(G)
    for c in range(6):
        is_b, a = c >= 3, c % 3
        if not (mask >> c) & 1 or not (tags & (BC_B if is_b else BC_E)):
            continue
        stag = (o != a) if is_b else (o == a)
        xi = (idx + F32(0.5)) if stag else idx
        xph = (xi * dx + F32(xmin_o)).astype(F32)
        s = tanhf((np.abs(xph - F32(xg_edge)).astype(F32) * F32(4.0)).astype(F32) / F32(ds))
        shape = [1] * D
        shape[D - 1 - o] = s.size
        s = s.reshape(shape)
        t = target[c][sl]
        tu = (t / dx).astype(F32) if a < D else t
        f = em[c][sl]
        em[c][sl] = ((s * f).astype(F32) + ((F32(1.0) - s).astype(F32) * tu).astype(F32)).astype(F32)
