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
