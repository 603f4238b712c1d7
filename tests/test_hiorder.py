"""Shape orders 4..11, CPU side: the generated B-spline tables and the golden
vectors of the compiled reference."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gen():
    spec = importlib.util.spec_from_file_location("gen_bspline", os.path.join(ROOT, "scripts", "gen_bspline.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_tables_are_the_generators_output():
    g = _gen()
    with open(g.PATH) as f:
        assert f.read() == g.render()


@pytest.mark.parametrize("order", range(4, 12))
def test_bspline_properties(order):
    """fp32 Horner evaluation of the tables as the kernel does: partition of unity to 2e-6, the
    fp64 closed form to 3e-7, support (O + 1) / 2, and the known central values of the reference's comments
    (particle_shapes.hpp: S4(0) = 115/192, S5(0) = 11/20, S6(0) = 5887/11520, S7(0) = 151/315)"""
    g = _gen()
    ps = [[np.float32(float(c)) for c in piece] for piece in g.local_pieces(order)]
    first = 0.5 if order % 2 == 0 else 1.0

    def S(x):
        x = np.float32(abs(x))
        p = 0 if x < first else int(x - np.float32(first)) + 1
        if p >= len(ps):
            return np.float32(0)
        t = x if p == 0 else np.float32(x - np.float32(first + (p - 1)))
        r = ps[p][order]
        for k in range(order - 1, -1, -1):
            r = np.float32(np.float32(r * t) + ps[p][k])
        return r

    rng = np.random.default_rng(order)
    for d in rng.random(50):
        xs = [d + n for n in range(-order, order + 1)]
        tot = sum(float(S(x)) for x in xs)
        assert abs(tot - 1.0) < 2e-6, (order, d, tot)
        for x in xs[::3]:
            assert abs(float(S(x)) - float(g.evaluate_exact(order, g.Fraction(float(np.float32(abs(x))))))) < 3e-7
    assert S((order + 1) / 2 + 1e-3) == 0
    central = {4: 115 / 192, 5: 11 / 20, 6: 5887 / 11520, 7: 151 / 315}
    if order in central:
        assert abs(float(S(0.0)) - central[order]) < 1e-7


def test_golden_rederived_where_the_reference_is_built():
    import hiorder_cases as hc
    from oracle import orc
    refs = {o: orc.reference(o) for o in hc.ORDERS}
    if any(r is None for r in refs.values()):
        pytest.skip("oracle/_ref/libref_o{4..11}.so not built (no reference tree)")
    z = np.load(os.path.join(ROOT, "tests", "golden", "hiorder_golden.npz"))
    for dim in (1, 2, 3):
        for order in (4, 7, 10, 11):
            got = hc.run(refs[order], dim, order)
            for k, v in got.items():
                assert np.array_equal(v, z[f"{dim}d/o{order}/{k}"], equal_nan=True), (dim, order, k)
