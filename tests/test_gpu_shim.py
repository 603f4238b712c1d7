"""The drop-in itself: the reference's entity.xc (Kokkos-CUDA sm_100 build, its own engine, pgen,
boundaries, exchanges and time loop) with integration/eb200_shim.hpp patched into the SRPIC
Minkowski dispatchers -- Faraday, Ampere, CurrentsAmpere, ParticlePush + CurrentsDeposit (fused),
SortSpatially call libentity_b200.so -- against the unpatched build of the same tree on the same
input (pgens/streaming, 2D, four species, 10 steps). Both binaries are built in the container by
oracle/build_entity_xc.sh {cuda, cuda_shim} and travel under baseline/_ref/; skipped where absent.

Same Kokkos random pool, same particle order: every particle in the same cell, offsets and momenta
to 2e-6, fields and currents to 2e-6 of their maximum after 10 steps (the two sides contract FMAs
differently and sum J with atomics)."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_entity_xc_with_and_without_the_shim():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    for fl in ("cuda", "cuda_shim"):
        if not os.path.exists(os.path.join(ROOT, "baseline", "_ref", fl, "entity_streaming.xc")):
            pytest.skip(f"baseline/_ref/{fl}/entity_streaming.xc not built (oracle/build_entity_xc.sh {fl})")
    sys.path.insert(0, os.path.join(ROOT, "integration"))
    import run_shim_check
    res = run_shim_check.parity()
    assert set(res) == {"step0", "step5", "step10"}
    for step, rec in res.items():
        moved, total = rec["particles_in_another_cell"]
        assert total > 20000 and moved == 0, (step, rec)
        assert rec["max_abs_diff_dx_u"] <= 2e-6, (step, rec)
        assert rec["em_rel_err"] <= 2e-6 and rec["cur_rel_err"] <= 2e-6, (step, rec)
        for k, v in rec.items():
            if k.endswith("_npart"):
                assert v[0] == v[1], (step, k, v)
