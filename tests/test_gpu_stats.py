"""eb200_stats_fields / eb200_stats_particles (C ABI) against the golden values of the
reference's reduced_stats.hpp kernels (tests/golden/stats_golden.npz)."""
import os

import numpy as np
import pytest

import stats_cases as sc
from helpers import to_device
from oracle import stats as ostats

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 2e-4  # of the sum of |terms| (see tests/test_stats.py)


@pytest.fixture(scope="module")
def eb():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import entity_b200
    return entity_b200


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "stats_golden.npz"))


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_field_stats(eb, gold, dim):
    import torch
    g, em, cur = sc.fields(dim)
    ctx = eb.Context(sc.GRIDS[dim], order=0, dx=sc.DX)
    d_em, d_cur = torch.from_numpy(em).cuda(), torch.from_numpy(cur).cuda()
    n0 = ctx.launch_count
    for d, name, what, comp in sc.field_cases():
        if d != dim:
            continue
        v = ctx.stats_fields(d_em, d_cur, what, max(comp, 1))
        vo, scale = ostats.fields(g, em, cur, sc.DX, what, max(comp, 1))
        ref = float(gold[f"f_{dim}d_{name}_{comp}"])
        assert abs(v - ref) <= RTOL * scale, (name, comp, v, ref)
        # same fp32 terms, fp64 sums on both sides: the restatement agrees much closer
        assert abs(v - vo) <= 2e-6 * scale, (name, comp, v, vo)
    assert ctx.launch_count > n0
    with pytest.raises(eb.EB200Error):
        ctx.stats_fields(d_em, None, eb.lib.STATS_JDOTE)
    with pytest.raises(eb.EB200Error):
        ctx.stats_fields(d_em, d_cur, eb.lib.STATS_E2, 4)
    ctx.close()


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_particle_stats(eb, gold, dim):
    ctx = eb.Context(sc.GRIDS[dim], order=0, dx=sc.DX)
    cache = {}
    for d, k, mass, charge, name, what, use_w, c1, c2 in sc.particle_cases():
        if d != dim:
            continue
        if k not in cache:
            g, p, n = sc.particles(dim, k)
            cache[k] = (g, p, n, to_device(p))
        g, p, n, arr = cache[k]
        v = ctx.stats_particles(arr, n, mass, charge, what, c1, c2, use_w)
        vo, scale = ostats.particles(g, p, n, mass, charge, sc.DX, what, c1, c2, use_w)
        ref = float(gold[f"p_{dim}d_s{k}_{name}_w{int(use_w)}_{c1}{c2}"])
        assert abs(v - ref) <= RTOL * scale + 1e-12, (name, k, c1, c2, v, ref)
        assert abs(v - vo) <= 2e-6 * scale + 1e-12, (name, k, c1, c2, v, vo)
        if name == "Npart":
            assert v == ref
    with pytest.raises(eb.EB200Error, match="massless"):
        g, p, n, arr = cache[0]
        ctx.stats_particles(arr, n, 0.0, 1.0, eb.lib.STATS_RHO)
    ctx.close()


@pytest.mark.parametrize("mname", ["spherical", "qspherical"])
def test_curvilinear_stats(eb, gold, mname):
    """2D (q)spherical SRPIC meshes against the reference's kernels instantiated with its own
    metric classes (golden keys fc_* / pc_*). No numpy restatement here: the compiled reference is
    the checker, and the tolerance is 2e-4 of a positive bound of the sum of |terms| -- the
    statistic itself where every term has one sign (E^2, B^2, N, Rho, Charge, T^00), (E^2 + B^2) / 2
    for E x B, (E^2 + J^2) / 2 for J.E, T^00 for the other T components."""
    import torch
    from entity_b200 import lib as L
    kind, ext = sc.CURV[mname]
    metric = L.METRIC_SPHERICAL if mname == "spherical" else L.METRIC_QSPHERICAL
    ctx = eb.Context(sc.GRIDS[2], order=0, metric=metric, metric_params=ext + [0.0])
    g, em, cur = sc.fields(2)
    d_em, d_cur = torch.from_numpy(em).cuda(), torch.from_numpy(cur).cuda()
    e2 = sum(ctx.stats_fields(d_em, d_cur, sc.FIELD_STATS["E2"], c) for c in (1, 2, 3))
    b2 = sum(ctx.stats_fields(d_em, d_cur, sc.FIELD_STATS["B2"], c) for c in (1, 2, 3))
    j6 = torch.zeros_like(d_em)
    j6[:3] = d_cur
    j2 = sum(ctx.stats_fields(j6, None, sc.FIELD_STATS["E2"], c) for c in (1, 2, 3))
    assert e2 > 0 and b2 > 0 and j2 > 0
    for name, what, comp in sc.curv_field_cases():
        v = ctx.stats_fields(d_em, d_cur, what, max(comp, 1))
        ref = float(gold[f"fc_{mname}_{name}_{comp}"])
        scale = {"B2": abs(ref), "E2": abs(ref), "ExB": 0.5 * (e2 + b2), "JdotE": 0.5 * (e2 + j2)}[name]
        assert abs(v - ref) <= RTOL * scale, (mname, name, comp, v, ref)
    cache = {}
    for k, mass, charge, name, what, use_w, c1, c2 in sc.curv_particle_cases():
        if k not in cache:
            gp, p, n = sc.curv_particles(k)
            cache[k] = (p, n, to_device(p))
            cache[k][2]["phi"] = torch.from_numpy(p.phi.copy()).cuda()
        p, n, arr = cache[k]
        v = ctx.stats_particles(arr, n, mass, charge, what, c1, c2, use_w)
        ref = float(gold[f"pc_{mname}_s{k}_{name}_w{int(use_w)}_{c1}{c2}"])
        if name == "T":
            scale = abs(float(gold[f"pc_{mname}_s{k}_T_w0_00"]))
        else:
            scale = abs(ref)
        assert abs(v - ref) <= RTOL * scale + 1e-12, (mname, name, k, c1, c2, v, ref)
        if name == "Npart":
            assert v == ref
    ctx.close()
