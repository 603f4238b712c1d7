"""CPU-side checks of the curvilinear / GR rows (no GPU needed):
 * the metric functions the kernels use (same source compiled for the host, eb200_metric_eval)
   against the committed golden vectors of the reference's metric classes -- bit-exact;
 * the golden vectors themselves against the compiled reference wherever oracle/_ref was built
   (pins the fixtures to the reference, like tests/test_oracle_vs_ref.py does for Minkowski);
 * the C ABI exports every curvilinear / GR symbol include/entity_b200.h declares."""
import os
import re

import numpy as np
import pytest

import curv_cases as cc
from entity_b200 import lib as L
from oracle import refcurv as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "curv_golden.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.mark.parametrize("mname", list(cc.SR_METRICS) + list(cc.GR_METRICS))
def test_metric_functions_bit_exact(golden, mname):
    m = cc.metric(mname)
    rng = np.random.default_rng(7)
    # same draw order as curv_cases.run_all
    for nm in list(cc.SR_METRICS) + list(cc.GR_METRICS):
        x1 = rng.uniform(-1, cc.N[0] + 1, 200).astype(np.float32)
        x2 = rng.uniform(0.02, cc.N[1] - 0.02, 200).astype(np.float32)
        if nm == mname:
            break
    got = L.metric_eval(m.kind, cc.N, m.params8(), x1, x2)
    want = golden[f"metric/{mname}"]
    assert got.shape == want.shape
    bad = ~((got == want) | (np.isnan(got) & np.isnan(want)))
    assert not bad.any(), f"{mname}: {bad.sum()} metric values differ, columns {np.unique(np.nonzero(bad)[1])}"


def test_golden_reproduces_from_reference(golden):
    if R.reference(0) is None:
        pytest.skip("oracle/_ref/libref_curv_o*.so not built (no reference tree at build time)")
    be = cc.RefBackend()
    out = cc.run_all(be, cc.RefSetup(be))
    for k, v in out.items():
        w = golden[k]
        assert np.array_equal(v, w, equal_nan=True), f"golden entry {k} is stale"


def test_golden_is_sane(golden):
    keys = list(golden.keys())
    assert len(keys) > 400
    # every family is present and non-trivial
    for fam in ("metric/", "sr_push/", "sr_fld/", "gr_push/", "gr_fld/"):
        ks = [k for k in keys if k.startswith(fam)]
        assert ks, fam
    j = [golden[k] for k in keys if k.endswith("/J")]
    assert all(np.isfinite(a).all() and np.abs(a).max() > 0 for a in j)
    # particles left through the absorbing faces and were reflected at the axis
    tags = [golden[k] for k in keys if k.startswith("sr_push/") and k.endswith("/tag")]
    assert any((t == 0).sum() > 30 for t in tags)


def test_c_abi_exports_curvilinear_symbols():
    lib = L.load()
    hdr = open(os.path.join(ROOT, "include", "entity_b200.h")).read()
    names = set(re.findall(r"\b(eb200_[a-z0-9_]+)\s*\(", hdr))
    want = {"eb200_faraday_sr", "eb200_ampere_sr", "eb200_currents_ampere_sr", "eb200_push_gr",
            "eb200_gr_aux_e", "eb200_gr_aux_h", "eb200_faraday_gr", "eb200_ampere_gr",
            "eb200_currents_ampere_gr", "eb200_time_average", "eb200_metric_eval"}
    assert want <= names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
