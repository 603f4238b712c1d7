"""Seeded cases for the output-staging kernels (eb200_fields_to_phys / eb200_prtls_to_phys):
shared by tests/golden/make_out_golden.py (the reference's kernels compiled in place) and
tests/test_gpu_output.py."""
import numpy as np

N = (20, 14)
NG = 2
PI = float(np.float32(np.pi))
# kind -> metric params [x1min, x1max, x2min, x2max, r0, h, a] (Minkowski: dx, xmin)
METRICS = {
    0: dict(dx=0.25, xmin=(-1.0, 0.5)),
    1: dict(mp=[1.0, 12.0, 0.0, PI, 0.0, 0.0, 0.0]),
    2: dict(mp=[1.0, 12.0, 0.0, PI, 0.2, 0.3, 0.0]),
    3: dict(mp=[1.2, 12.0, 0.0, PI, 0.0, 0.0, 0.9]),
    4: dict(mp=[1.2, 12.0, 0.0, PI, 0.1, 0.25, 0.9]),
    5: dict(mp=[1.2, 12.0, 0.0, PI, 0.0, 0.0, 0.0]),
}
# (interp, convert, comps_from, comps_to)
FIELD_CASES = [(1, 1, (0, 1, 2), (3, 4, 5)), (2, 1, (3, 4, 5), (0, 1, 2)), (1, 2, (0, 1, 2), (0, 1, 2)),
               (2, 3, (3, 4, 5), (3, 4, 5)), (0, 0, (2, 0, 1), (5, 3, 4)), (0, 1, (0, 1, 2), (0, 1, 2))]
NPART, STRIDE = 257, 3


def field(kind):
    shape = (6, N[1] + 2 * NG, N[0] + 2 * NG)
    return np.random.default_rng(300 + kind).standard_normal(shape).astype(np.float32)


def particles(kind):
    rng = np.random.default_rng(400 + kind)
    p = dict(i1=rng.integers(0, N[0], NPART).astype(np.int32), i2=rng.integers(0, N[1], NPART).astype(np.int32),
             dx1=rng.random(NPART, dtype=np.float32), dx2=rng.random(NPART, dtype=np.float32),
             ux1=rng.standard_normal(NPART).astype(np.float32), ux2=rng.standard_normal(NPART).astype(np.float32),
             ux3=rng.standard_normal(NPART).astype(np.float32),
             weight=rng.uniform(0.5, 2.0, NPART).astype(np.float32),
             phi=rng.uniform(0.0, 2 * np.pi, NPART).astype(np.float32))
    return p
