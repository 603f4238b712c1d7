"""eb200_fields_to_phys / eb200_prtls_to_phys (C ABI) against the reference's own
FieldsToPhys_kernel / PrtlToPhys_kernel compiled in place (tests/golden/out_golden.npz, made by
tests/golden/make_out_golden.py), Minkowski and the five curvilinear metrics.

Interpolation and the Minkowski conversions are bit-exact; the curvilinear conversions go
through expf / sinf / cosf / sqrtf of the metric (CUDA vs glibc, last ulp): 5e-6 relative."""
import ctypes as C
import os

import numpy as np
import pytest

import out_cases as oc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def eb():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import entity_b200
    return entity_b200


def make_ctx(eb, kind):
    from entity_b200 import lib as L
    d = oc.METRICS[kind]
    if kind == 0:
        return eb.Context(oc.N, order=0, dx=d["dx"], xmin=tuple(d["xmin"]) + (0.0,))
    return eb.Context(oc.N, order=0, metric=kind, metric_params=d["mp"])


@pytest.mark.parametrize("kind", sorted(oc.METRICS))
def test_fields_to_phys(eb, kind):
    import torch
    from entity_b200 import lib as L
    gold = np.load(os.path.join(ROOT, "tests", "golden", "out_golden.npz"))
    ctx = make_ctx(eb, kind)
    st = L.Context._stream(None)
    G = oc.NG
    act = (slice(None), slice(G, G + oc.N[1]), slice(G, G + oc.N[0]))
    for k, (interp, conv, cf, ct) in enumerate(oc.FIELD_CASES):
        src = torch.from_numpy(oc.field(kind)).cuda()
        dst = torch.zeros_like(src)
        ctx._check(ctx.lib.eb200_fields_to_phys(
            ctx.handle, C.c_void_p(src.data_ptr()), 6, C.c_void_p(dst.data_ptr()), 6,
            (C.c_int * 3)(*cf), (C.c_int * 3)(*ct), interp, conv, st))
        out, ref = dst.cpu().numpy(), gold[f"fld_m{kind}_c{k}"]
        assert np.array_equal(out == 0, ref == 0), f"case {k}: cells outside the active range touched"
        if kind == 0 or conv == 0:
            assert np.array_equal(out[act], ref[act]), f"metric {kind} case {k}"
        else:
            np.testing.assert_allclose(out[act], ref[act], rtol=5e-6, atol=1e-30,
                                       err_msg=f"metric {kind} case {k}")
    ctx.close()


@pytest.mark.parametrize("kind", sorted(oc.METRICS))
def test_prtls_to_phys(eb, kind):
    import torch
    from entity_b200 import lib as L
    gold = np.load(os.path.join(ROOT, "tests", "golden", "out_golden.npz"))[f"prt_m{kind}"]
    ctx = make_ctx(eb, kind)
    p = oc.particles(kind)
    arrays = {nm: torch.from_numpy(v).cuda() for nm, v in p.items()}
    arrays["tag"] = torch.ones(oc.NPART, dtype=torch.int16, device="cuda")
    for nm in ("i1_prev", "i2_prev"):
        arrays[nm] = arrays[nm[:2]].clone()
    for nm in ("dx1_prev", "dx2_prev"):
        arrays[nm] = arrays[nm[:3]].clone()
    s = L.Context.prtls_struct(arrays)
    nout = (oc.NPART + oc.STRIDE - 1) // oc.STRIDE
    bufs = [torch.zeros(nout, dtype=torch.float32, device="cuda") for _ in range(7)]
    ctx._check(ctx.lib.eb200_prtls_to_phys(ctx.handle, C.byref(s), oc.NPART, oc.STRIDE,
                                           *[C.c_void_p(b.data_ptr()) for b in bufs],
                                           L.Context._stream(None)))
    out = np.stack([b.cpu().numpy() for b in bufs])
    rows = [0, 1, 3, 4, 5, 6] if kind == 0 else range(7)  # x3 is not written for 2D Minkowski
    for r in rows:
        if kind == 0:
            assert np.array_equal(out[r], gold[r]), f"row {r}"
        else:
            np.testing.assert_allclose(out[r], gold[r], rtol=3e-6, atol=2e-6, err_msg=f"row {r}")
    ctx.close()
