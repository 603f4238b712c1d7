"""TEST INFRASTRUCTURE ONLY -- the turbulence antenna's external current
(pgens/turbulence/pgen.hpp:223-279) and its addition in CurrentsAmpere_kernel<D, ExtCurrent>
(src/kernels/ampere_mink.hpp:134-215), restated in numpy fp32 with glibc's cosf / sinf (what the
reference's host build calls), mode by mode in the reference's order."""
from __future__ import annotations

import ctypes
import ctypes.util

import numpy as np

F32 = np.float32
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _f in ("cosf", "sinf"):
    getattr(_libm, _f).restype = ctypes.c_float
    getattr(_libm, _f).argtypes = [ctypes.c_float]


def _ew(fn, x):
    x = np.asarray(x, F32)
    out = np.empty(x.shape, F32)
    xf, of = x.ravel(), out.ravel()
    f = getattr(_libm, fn)
    for k in range(xf.size):
        of[k] = f(float(xf[k]))
    return out


def mode_table(dim, k, a_real, a_imag, a_real_inv, a_imag_inv):
    """eb200_ext_current_t contents for the antenna: per component c and mode m the prefactor that
    multiplies (a_real cos - a_imag sin), exactly as pgen.hpp:223-279 forms it in fp32
    (jx1 -= TWO k0 k2 (...), jx2 -= TWO k1 k2 (...), jx3 += TWO kperp^2 (...); the 2D jx3 adds the
    inverse-helicity amplitudes with the same prefactor)."""
    k = np.asarray(k, F32)
    nm = k.shape[1]
    pref = np.zeros((3, nm), F32)
    pref2 = np.zeros((3, nm), F32)
    two = F32(2.0)
    kperp2 = (k[0] * k[0] + k[1] * k[1]).astype(F32)
    if dim == 3:
        pref[0] = -((two * k[0]).astype(F32) * k[2]).astype(F32)
        pref[1] = -((two * k[1]).astype(F32) * k[2]).astype(F32)
        pref[2] = (two * kperp2).astype(F32)
    else:
        pref[2] = (two * kperp2).astype(F32)
        pref2[2] = pref[2]
    kk = np.zeros((3, nm), F32)
    kk[:k.shape[0]] = k
    return dict(nmodes=nm, k=kk, pref=pref, a_real=np.asarray(a_real, F32), a_imag=np.asarray(a_imag, F32),
                pref2=pref2, a_real2=np.asarray(a_real_inv, F32), a_imag2=np.asarray(a_imag_inv, F32))


def add_ext_current(g, cur, tab, ppc0, dx, xmin):
    """J_c += ppc0 * jx_c(x) on every ACTIVE cell, x = the component's own node."""
    D, G = g.dim, g.ng
    dx = F32(dx)
    ax = []
    for a in range(D):
        ax.append(np.arange(g.n[a], dtype=F32))
    for c in range(3):
        if not (np.any(tab["pref"][c] != 0) or np.any(tab["pref2"][c] != 0)):
            # a component without any mode still goes through "+= ppc0 * 0": a no-op
            continue
        xs = []
        for a in range(D):
            f = ax[a] + F32(0.5) if a == c else ax[a]
            xs.append((f * dx).astype(F32) + F32(xmin[a]))
        grids = np.meshgrid(*xs[::-1], indexing="ij")[::-1]  # grids[a] indexed [.., x2, x1]
        j = np.zeros(grids[0].shape, F32)
        for m in range(tab["nmodes"]):
            kr = (tab["k"][0][m] * grids[0]).astype(F32)
            kr = (kr + (tab["k"][1][m] * grids[1]).astype(F32)).astype(F32)
            if D == 3:
                kr = (kr + (tab["k"][2][m] * grids[2]).astype(F32)).astype(F32)
            cs, sn = _ew("cosf", kr), _ew("sinf", kr)
            t = ((tab["a_real"][m] * cs).astype(F32) - (tab["a_imag"][m] * sn).astype(F32)).astype(F32)
            j = (j + (tab["pref"][c][m] * t).astype(F32)).astype(F32)
            if tab["pref2"][c][m] != 0:
                t2 = ((tab["a_real2"][m] * cs).astype(F32) - (tab["a_imag2"][m] * sn).astype(F32)).astype(F32)
                j = (j + (tab["pref2"][c][m] * t2).astype(F32)).astype(F32)
        sl = (c,) + tuple(slice(G, G + g.n[a]) for a in reversed(range(D)))
        cur[sl] = (cur[sl] + (F32(ppc0) * j).astype(F32)).astype(F32)
