"""eb200_match_fields (C ABI) against the golden arrays of the reference's
MatchBoundaries_kernel (tests/golden/bcs_golden.npz)."""
import os

import numpy as np
import pytest

import bcs_cases as bc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-6  # of max|field|: tanhf differs from glibc's in the last ulp (tests/test_bcs.py)


@pytest.fixture(scope="module")
def eb():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import entity_b200
    return entity_b200


@pytest.mark.parametrize("case", bc.cases(), ids=[c[0] for c in bc.cases()])
def test_match_fields(eb, case):
    import torch
    gold = np.load(os.path.join(ROOT, "tests", "golden", "bcs_golden.npz"))
    name, dim, o, sign, nds, tags, b_only = case
    g, em, xg_edge, ds, rmin, rmax = bc.setup(dim, o, sign, nds)
    ctx = eb.Context(bc.GRIDS[dim], order=0, dx=bc.DX,
                     xmin=tuple(bc.XMIN[dim]) + (0.0,) * (3 - dim))
    d_em = torch.from_numpy(em.copy()).cuda()
    d_t = torch.from_numpy(bc.target(dim)).cuda()
    n0 = ctx.launch_count
    ctx.match_fields(d_em, d_t, o, xg_edge, ds, tags, 0b111000 if b_only else 63, rmin, rmax)
    out = d_em.cpu().numpy()
    ref = gold[name]
    assert ctx.launch_count == n0 + 1
    assert np.abs(out - ref).max() <= RTOL * np.abs(ref).max()
    inside = np.zeros(em.shape[1:], bool)
    inside[tuple(slice(rmin[d], rmax[d]) for d in reversed(range(dim)))] = True
    assert np.array_equal(out[:, ~inside], em[:, ~inside])
    with pytest.raises(eb.EB200Error):
        ctx.match_fields(d_em, d_t, dim, xg_edge, ds, tags, 63, rmin, rmax)  # direction >= dim
    with pytest.raises(eb.EB200Error):
        ctx.match_fields(d_em, d_t, o, xg_edge, 0.0, tags, 63, rmin, rmax)  # ds <= 0
    ctx.close()


def _bcs2_cases():
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "make_bcs2_golden", os.path.join(ROOT, "tests", "golden", "make_bcs2_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m, list(enumerate(m.cases()))


@pytest.mark.parametrize("k", range(18))
def test_conductor_and_axis_fields(eb, k):
    """eb200_conductor_fields (Minkowski 2D) and eb200_axis_fields (spherical 2D) against the
    reference's ConductorBoundaries_kernel / AxisBoundaries_kernel compiled in place
    (tests/golden/bcs2_golden.npz): copies and sign flips only, bit for bit."""
    import ctypes as C

    import torch
    from entity_b200 import lib as L
    m, cs = _bcs2_cases()
    _, (name, kind, o, sign, tags) = cs[k]
    gold = np.load(os.path.join(ROOT, "tests", "golden", "bcs2_golden.npz"))[name]
    _, em = m.field(1000 + k)
    d = torch.from_numpy(em.copy()).cuda()
    st = L.Context._stream(None)
    if kind == "conductor":
        ctx = eb.Context(m.N, order=0, dx=0.5)
        ctx._check(ctx.lib.eb200_conductor_fields(ctx.handle, C.c_void_p(d.data_ptr()), o, sign, tags, st))
    else:
        ctx = eb.Context(m.N, order=0, metric=L.METRIC_SPHERICAL,
                         metric_params=[1.0, 10.0, 0.0, float(np.float32(np.pi))])
        ctx._check(ctx.lib.eb200_axis_fields(ctx.handle, C.c_void_p(d.data_ptr()), sign, tags, st))
    out = d.cpu().numpy()
    assert np.array_equal(out.view(np.uint32) & 0x7fffffff, gold.view(np.uint32) & 0x7fffffff)
    assert np.array_equal(out == 0, gold == 0) and np.array_equal(np.signbit(out)[out != 0], np.signbit(gold)[gold != 0])
    ctx.close()
