"""Seeded Minkowski-path cases shared by tests/golden/make_mink_golden.py (runs the REFERENCE's
kernels, oracle/_ref/libref_o*.so, and stores their outputs) and tests/test_golden_mink.py
(re-runs the oracle restatement -- and the compiled reference where present -- on the same
inputs and compares bit for bit)."""
from __future__ import annotations

import numpy as np

from helpers import random_fields, random_particles, smooth_fields
from oracle import orc
from test_oracle_vs_ref import DIMS, PUSH_CASES

STENCIL = np.array([0.01, 0.02, 0.03, 0.015, 0.012, 0.022, 0.011, 0.017, 0.019], np.float32)


def push_cases():
    out = []
    for dim in (1, 2, 3):
        for order in (0, 1, 2, 3):
            for case in (range(len(PUSH_CASES)) if order == 0 else (0, 2, 8)):
                out.append((dim, order, case))
    return out


def run_all(impl_for_order):
    """impl_for_order(order) -> an orc.Impl (oracle port or compiled reference)."""
    out = {}
    i0 = impl_for_order(0)
    for dim in (1, 2, 3):
        g = orc.Grid.make(DIMS[dim], 2)
        for sname, st in (("std", None), ("ext", STENCIL)):
            em = random_fields(g, 6, 1)
            i0.faraday(g, em, 0.21, 0.37, st)
            out[f"fld/{dim}d/{sname}/faraday"] = em.copy()
            i0.ampere(g, em, 0.45, 0.4)
            out[f"fld/{dim}d/{sname}/ampere"] = em.copy()
            j = random_fields(g, 3, 2)
            i0.currents_ampere(g, em, j, -0.013, 16.0)
            out[f"fld/{dim}d/{sname}/cur_ampere_E"] = em.copy()
            out[f"fld/{dim}d/{sname}/cur_ampere_J"] = j.copy()
        for fname, kind in (("periodic", orc.FBC_PERIODIC), ("conductor", orc.FBC_CONDUCTOR)):
            buff, a = random_fields(g, 3, 3), random_fields(g, 3, 4)
            i0.filter_pass(g, a, buff, [kind] * 6)
            out[f"fld/{dim}d/filter/{fname}"] = a
    for dim, order, case in push_cases():
        impl = impl_for_order(order)
        g = orc.Grid.make(DIMS[dim], orc.nghosts_for(order))
        kw = dict(PUSH_CASES[case])
        pbc = kw.pop("pbc", "periodic")
        code = dict(periodic=orc.PBC_PERIODIC, absorb=orc.PBC_ABSORB, reflect=orc.PBC_REFLECT,
                    none=orc.PBC_NONE)[pbc]
        dx = 0.5
        ctx = orc.make_pusher(dt=0.45 * dx, omegaB0=0.7, mass=1.0, charge=-1.0, dx=dx,
                              xmin=[0.1, 0.2, 0.3], pbc=[code] * 6, **kw)
        em = smooth_fields(g, 10 + dim, amp=0.6)
        n = 700
        p = random_particles(g, n, 100 + case, umag=2.0, dead_frac=0.05)
        j = np.zeros(g.shape(3), np.float32)
        for step in range(2):
            impl.push(g, order, ctx, p, n, em)
            if pbc == "none":
                o = np.zeros(n, bool)
                for a, nm in enumerate(["i1", "i2", "i3"][:dim]):
                    o |= (getattr(p, nm) < 0) | (getattr(p, nm) >= g.n[a])
                p.tag[o] = 0
            impl.deposit(g, order, p, n, -1.0, ctx.dt, dx, j)
        key = f"prtl/{dim}d/o{order}/case{case}"
        for nm in p.names():
            if nm == "phi" or (nm[-1] in "23" and int(nm[-1]) > dim and nm[0] in "id") or \
                    (nm.endswith("_prev") and int(nm[-6]) > dim):
                continue
            out[f"{key}/{nm}"] = getattr(p, nm).copy()
        out[f"{key}/J"] = j
    return out
