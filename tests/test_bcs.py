"""Matching field boundaries: the numpy restatement (oracle/bcs.py)
against the golden arrays produced by the reference's own MatchBoundaries_kernel, and the golden
arrays re-derived from the compiled reference where /root/reference exists."""
import os

import numpy as np
import pytest

import bcs_cases as bc
from oracle import bcs as obcs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "bcs_golden.npz"))
# tanh of numpy / CUDA vs glibc's tanhf: last-ulp differences in s; everything else is the same
# fp32 operation sequence. Stated tolerance: 1e-6 of max|field|.
RTOL = 1e-6


@pytest.mark.parametrize("case", bc.cases(), ids=[c[0] for c in bc.cases()])
def test_match_restatement(case):
    name, dim, o, sign, nds, tags, b_only = case
    g, em, xg_edge, ds, rmin, rmax = bc.setup(dim, o, sign, nds)
    em0 = em.copy()
    obcs.match_fields(g, em, bc.target(dim), o, bc.DX, bc.XMIN[dim][o], xg_edge, ds, tags,
                      0b111000 if b_only else 63, rmin, rmax)
    ref = GOLD[name]
    assert np.abs(em - ref).max() <= RTOL * np.abs(ref).max()
    # cells outside the matching range are left alone (by the reference too)
    inside = np.zeros(em.shape[1:], bool)
    inside[tuple(slice(rmin[d], rmax[d]) for d in reversed(range(dim)))] = True
    assert np.array_equal(em[:, ~inside], em0[:, ~inside])
    assert np.array_equal(ref[:, ~inside], em0[:, ~inside])
    assert (ref != em0).any()
    if b_only or not tags & bc.BC_E:
        assert np.array_equal(ref[:3], em0[:3])
    if not tags & bc.BC_B:
        assert np.array_equal(ref[3:], em0[3:])


def test_golden_rederived_from_reference():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_bcs.so")):
        pytest.skip("oracle/_ref/libref_bcs.so not built (no reference tree here)")
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "make_bcs_golden", os.path.join(ROOT, "tests", "golden", "make_bcs_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = mod.run_all(mod.load())
    assert set(out) == set(GOLD.files)
    for k, v in out.items():
        assert np.array_equal(v, GOLD[k]), k
