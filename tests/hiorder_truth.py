"""fp64 restatement of the 1D Esirkepov deposit with the closed-form cardinal B-spline
(test infrastructure): the yardstick that separates OUR rounding from the REFERENCE's in the
shape orders whose monomial polynomials cancel badly in fp32 (O >= 9)."""
from __future__ import annotations

import math

import numpy as np


def bspline64(order, x):
    """B_O(x) = 1/O! sum_k (-1)^k C(O+1, k) (x + (O+1)/2 - k)_+^O, evaluated piecewise in exact
    integer / rational form via Python ints where it matters: here float64 with the local
    expansion is enough (error ~1e-15)"""
    from fractions import Fraction
    x = Fraction(float(abs(x)))
    half = Fraction(order + 1, 2)
    tot = Fraction(0)
    for k in range(order + 2):
        y = x + half - k
        if y > 0:
            tot += (-1) ** k * math.comb(order + 1, k) * y ** order
    return float(tot / math.factorial(order))


def deposit_1d(order, ng, n1, i, dx, i_prev, dx_prev, ux, weight, tag, charge, dt, dxc):
    """J[3, n1 + 2 ng] in float64: currents_deposit.hpp's 1D branch (jx1 = running sum of
    -Q (S_fin - S_init), jx2/3 = Q v (S_fin + S_init) / 2) with exact shapes on a wide window"""
    J = np.zeros((3, n1 + 2 * ng))
    inv_dt = 1.0 / dt
    for p in range(len(i)):
        if tag[p] == 0:
            continue
        u = [float(ux[c][p]) for c in range(3)]
        gamma = math.sqrt(1.0 + u[0] ** 2 + u[1] ** 2 + u[2] ** 2)
        v2, v3 = u[1] / gamma, u[2] / gamma
        coeff = float(weight[p]) * charge
        Q = coeff * inv_dt
        x0 = float(i_prev[p]) + float(dx_prev[p])
        x1 = float(i[p]) + float(dx[p])
        lo = min(int(i_prev[p]), int(i[p])) - (order + 1) // 2 - 1
        hi = max(int(i_prev[p]), int(i[p])) + (order + 1) // 2 + 2
        acc = 0.0
        for n in range(lo, hi + 1):
            s0, s1 = bspline64(order, x0 - n), bspline64(order, x1 - n)
            acc -= Q * (s1 - s0)
            if not 0 <= n + ng < J.shape[1]:
                continue  # outside both supports: s0 = s1 = 0 and acc = 0 (or its 1e-17 residue)
            J[0, n + ng] += acc
            J[1, n + ng] += coeff * v2 * 0.5 * (s0 + s1)
            J[2, n + ng] += coeff * v3 * 0.5 * (s0 + s1)
    return J
