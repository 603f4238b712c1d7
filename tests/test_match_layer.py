"""eb200_match_layer (pure host code: srpic::MatchFieldsIn's box + Mesh::Intersects /
Mesh::ExtentToRange) against the reference's own Mesh class on 400 seeded domains: the same
intersect decision, edge and index range, exactly."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gen():
    spec = importlib.util.spec_from_file_location(
        "make_match_layer_golden", os.path.join(ROOT, "tests", "golden", "make_match_layer_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_match_layer_matches_reference_mesh():
    from entity_b200 import lib as L
    mod = _gen()
    gold = np.load(os.path.join(ROOT, "tests", "golden", "match_layer_golden.npz"))["results"]
    rows = mod.cases()
    assert len(rows) == len(gold)
    hit = 0
    for row, ref in zip(rows, gold):
        dim, nn, dx, lmin, lmax, glo, ghi, o, sign, ds = row
        g = L.Grid.make(nn[:dim], 2)
        face = L.match_layer(g, dx, lmin, lmax, glo, ghi, o, sign, ds)
        if ref[0] == 0:
            assert face is None, row
            continue
        hit += 1
        assert face is not None, row
        fo, edge, fds, rmin, rmax = face
        assert fo == o and np.float32(edge) == np.float32(ref[1]) and np.float32(fds) == np.float32(ds)
        assert rmin == [int(v) for v in ref[2:2 + dim]], (row, rmin, ref)
        assert rmax == [int(v) for v in ref[5:5 + dim]], (row, rmax, ref)
    assert 50 < hit < len(rows) - 50  # both outcomes are exercised


def test_match_layer_bad_arguments():
    from entity_b200 import lib as L
    g = L.Grid.make((8, 8), 2)
    with pytest.raises(L.EB200Error):
        L.match_layer(g, 0.5, [0, 0], [4, 4], 0.0, 4.0, 2, 1, 1.0)  # direction outside the mesh
    with pytest.raises(L.EB200Error):
        L.match_layer(g, 0.5, [0, 0], [4, 4], 0.0, 4.0, 0, 0, 1.0)  # sign 0


def test_golden_rederived_from_reference():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_mesh.so")):
        pytest.skip("oracle/_ref/libref_mesh.so not built (no reference tree here)")
    mod = _gen()
    gold = np.load(os.path.join(ROOT, "tests", "golden", "match_layer_golden.npz"))["results"]
    assert np.array_equal(mod.run_all(mod.load()), gold)
