"""Seeded push + deposit cases for the particle shape orders 4..11, shared by
tests/golden/make_hiorder_golden.py (runs the REFERENCE's kernels compiled in place for each
SHAPE_ORDER: oracle/_ref/libref_o{4..11}.so) and the tests that compare with its output."""
from __future__ import annotations

import numpy as np

from helpers import random_particles, smooth_fields
from oracle import orc

DIMS = {1: (37,), 2: (23, 17), 3: (11, 9, 13)}
ORDERS = tuple(range(4, 12))
NPART = {1: 400, 2: 400, 3: 150}
DX = 0.5
STEPS = 2


def setup(dim, order):
    g = orc.Grid.make(DIMS[dim], orc.nghosts_for(order))
    ctx = orc.make_pusher(dt=0.45 * DX, omegaB0=0.7, mass=1.0, charge=-1.0, dx=DX,
                          xmin=[0.1, 0.2, 0.3], pbc=[orc.PBC_PERIODIC] * 6)
    em = smooth_fields(g, 30 + dim, amp=0.6)
    n = NPART[dim]
    p = random_particles(g, n, 500 + 10 * dim + order, umag=1.5, dead_frac=0.05)
    return g, ctx, em, p, n


def names(dim):
    out = ["ux1", "ux2", "ux3", "weight", "tag"]
    for a in range(1, dim + 1):
        out += [f"i{a}", f"dx{a}", f"i{a}_prev", f"dx{a}_prev"]
    return out


def run(impl, dim, order):
    """STEPS x (push, deposit) with `impl` (an orc.Impl); returns {name: array}"""
    g, ctx, em, p, n = setup(dim, order)
    j = np.zeros(g.shape(3), np.float32)
    for _ in range(STEPS):
        impl.push(g, order, ctx, p, n, em)
        impl.deposit(g, order, p, n, -1.0, ctx.dt, DX, j)
    out = {nm: getattr(p, nm).copy() for nm in names(dim)}
    out["J"] = j
    return out


def run_all(impl_for_order):
    out = {}
    for dim in (1, 2, 3):
        for order in ORDERS:
            for k, v in run(impl_for_order(order), dim, order).items():
                out[f"{dim}d/o{order}/{k}"] = v
    return out
