"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's reduced statistics for a
Minkowski domain: kernel::ReducedFields_kernel (src/kernels/reduced_stats.hpp:25-386) and
kernel::ReducedParticleMoments_kernel (:400-536), i.e. the local sums ReduceFields /
ComputeMoments obtain from Kokkos::parallel_reduce (src/framework/domain/metadomain_stats.cpp:
72-183). Per-term arithmetic in fp32 as the reference, accumulation in fp64 (the reference
accumulates in real_t in backend order: parity is to summation-order tolerance). Pinned against
the compiled reference: tests/golden/stats_golden.npz (tests/test_stats.py).

Every function returns (sum, sum of |terms|): the second value is the scale of the tolerance."""
import numpy as np

F32 = np.float32


def _active(a, g, shift):
    """view of field plane `a` (axes reversed: x1 last) over the active cells shifted by
    `shift[d]` in dimension d"""
    G = g.ng
    sl = []
    for d in reversed(range(g.dim)):
        sl.append(slice(G + shift[d], G + shift[d] + g.n[d]))
    return a[tuple(sl)]


def _centred(fld, g, c0, c, is_b):
    """component c of the triple starting at plane c0, averaged to the cell centre over the
    active dimensions in which it is not staggered (reduced_stats.hpp:81-128, 186-251, 296-386);
    same order of additions, HALF / INV_4 of the sum"""
    avg = [a for a in range(g.dim) if not ((a != c) if is_b else (a == c))]
    s = None
    for m in range(1 << len(avg)):
        shift = [0, 0, 0]
        for q, a in enumerate(avg):
            shift[a] = (m >> q) & 1
        v = _active(fld[c0 + c], g, shift)
        s = v.copy() if s is None else (s + v).astype(F32)
    return (F32([1.0, 0.5, 0.25, 0.125][len(avg)]) * s).astype(F32)


def fields(g, em, cur, dx, what, comp):
    """what: 0 B2, 1 E2, 2 ExB, 3 JdotE; comp 1..3"""
    D = g.dim
    dx = F32(dx)
    sdh = F32(dx ** D) if D > 1 else dx
    sdh = {1: dx, 2: F32(dx * dx), 3: F32(F32(dx * dx) * dx)}[D]
    fT = lambda a: dx if a < D else F32(1.0)
    fD = lambda a: F32(dx * dx) if a < D else F32(1.0)
    c = comp - 1
    if what in (0, 1):
        u = _active(em[(3 if what == 0 else 0) + c], g, [0, 0, 0]).astype(F32)
        t = (u * (u * fD(c)).astype(F32)).astype(F32) * sdh
    elif what == 2:
        a, b = (c + 1) % 3, (c + 2) % 3
        ea = (_centred(em, g, 0, a, False) * fT(a)).astype(F32)
        eb = (_centred(em, g, 0, b, False) * fT(b)).astype(F32)
        ba = (_centred(em, g, 3, a, True) * fT(a)).astype(F32)
        bb = (_centred(em, g, 3, b, True) * fT(b)).astype(F32)
        t = ((ea * bb).astype(F32) - (eb * ba).astype(F32)).astype(F32) * sdh
    else:
        s = None
        for a in range(3):
            e = (_centred(em, g, 0, a, False) * fT(a)).astype(F32)
            j = (_centred(cur, g, 0, a, False) * fT(a)).astype(F32)
            s = (e * j).astype(F32) if s is None else (s + (e * j).astype(F32)).astype(F32)
        t = s * sdh
    t = t.astype(F32).astype(np.float64)
    return float(t.sum()), float(np.abs(t).sum())


def particles(g, p, n, mass, charge, dx, what, c1=0, c2=0, use_weights=False):
    """what: 0 Npart, 1 N, 2 Rho, 3 Charge, 4 T^{c1 c2}"""
    alive = p.tag[:n] == 1
    dx = F32(dx)
    dV = {1: dx, 2: F32(dx * dx), 3: F32(F32(dx * dx) * dx)}[g.dim]
    if what == 0:
        t = alive.astype(np.float64)
    elif what in (1, 2, 3):
        contrib = F32({1: 1.0, 2: mass, 3: charge}[what])
        w = p.weight[:n].astype(F32) if use_weights else np.full(n, contrib, F32)
        t = np.where(alive, (dV * w).astype(F32), F32(0)).astype(np.float64)
    else:
        u = [p.ux1[:n].astype(F32), p.ux2[:n].astype(F32), p.ux3[:n].astype(F32)]
        usq = ((u[0] * u[0] + u[1] * u[1]).astype(F32) + u[2] * u[2]).astype(F32)
        if mass == 0.0:
            energy = np.sqrt(usq).astype(F32)
        else:
            energy = (F32(mass) * np.sqrt((F32(1.0) + usq).astype(F32)).astype(F32)).astype(F32)
        coeff = np.ones(n, F32)
        for cc in (c1, c2):
            coeff = (coeff * (energy if cc == 0 else u[cc - 1])).astype(F32)
        with np.errstate(divide="ignore", invalid="ignore"):
            t = ((dV * coeff).astype(F32) / energy).astype(F32)
        t = np.where(alive, t, F32(0)).astype(np.float64)
    return float(t.sum()), float(np.abs(t).sum())
