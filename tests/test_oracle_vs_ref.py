"""Pins the oracle restatement (oracle/oracle_kernels.cpp) against the reference's own
kernel headers compiled in place (oracle/_ref/libref_o*.so, built by `make -C oracle ref`
from /root/reference). Bit-exact, seeded random inputs. Skipped on machines where the
reference tree -- and therefore oracle/_ref -- is absent and was not shipped."""
import numpy as np
import pytest

from helpers import (assert_bits_equal, assert_prtls_equal, random_fields, random_particles,
                     smooth_fields)

DIMS = {1: (37,), 2: (23, 17), 3: (11, 9, 13)}


def _ref_or_skip(orc_mod, order):
    ref = orc_mod.reference(order)
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    return ref


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("stencil", [None, "ext"])
def test_faraday_ampere(orc_mod, dim, stencil):
    orc, ref = orc_mod.oracle(), _ref_or_skip(orc_mod, 0)
    g = orc_mod.Grid.make(DIMS[dim], 2)
    em0 = random_fields(g, 6, 1)
    st = None
    if stencil:
        st = np.array([0.01, 0.02, 0.03, 0.015, 0.012, 0.022, 0.011, 0.017, 0.019], np.float32)
    a, b = em0.copy(), em0.copy()
    orc.faraday(g, a, 0.21, 0.37, st)
    ref.faraday(g, b, 0.21, 0.37, st)
    assert_bits_equal(a, b, "faraday")
    assert not np.array_equal(a, em0)
    orc.ampere(g, a, 0.45, 0.4)
    ref.ampere(g, b, 0.45, 0.4)
    assert_bits_equal(a, b, "ampere")
    ja, jb = random_fields(g, 3, 2), random_fields(g, 3, 2)
    orc.currents_ampere(g, a, ja, -0.013, 16.0)
    ref.currents_ampere(g, b, jb, -0.013, 16.0)
    assert_bits_equal(a, b, "currents_ampere E")
    assert_bits_equal(ja, jb, "currents_ampere J")


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("fbc", ["periodic", "conductor"])
def test_filter(orc_mod, dim, fbc):
    orc, ref = orc_mod.oracle(), _ref_or_skip(orc_mod, 0)
    g = orc_mod.Grid.make(DIMS[dim], 2)
    kind = orc_mod.FBC_PERIODIC if fbc == "periodic" else orc_mod.FBC_CONDUCTOR
    bc = [kind] * 6
    buff = random_fields(g, 3, 3)
    a, b = random_fields(g, 3, 4), random_fields(g, 3, 4)
    orc.filter_pass(g, a, buff, bc)
    ref.filter_pass(g, b, buff, bc)
    assert_bits_equal(a, b, "filter")


PUSH_CASES = [
    dict(pusher_flags=2),                                   # Boris
    dict(pusher_flags=4),                                   # Vay
    dict(pusher_flags=2 | 8, gca_larmor_max=5.0, gca_e_ovr_b_sqr_max=0.81),  # Boris+GCA
    dict(pusher_flags=1),                                   # photon
    dict(pusher_flags=2, drag_flags=3, sync_coeff=0.01, compton_coeff=0.02),
    dict(pusher_flags=2, pbc="absorb"),
    dict(pusher_flags=4, pbc="reflect"),
    dict(pusher_flags=2, pbc="none", tag_outgoing=0),
    dict(pusher_flags=2, has_atmosphere=1, atm_gx1=-0.3, atm_x_surf=0.4, atm_ds=0.7),
]


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("order", [0, 1, 2, 3])
@pytest.mark.parametrize("case", range(len(PUSH_CASES)))
def test_push_and_deposit(orc_mod, dim, order, case):
    orc, ref = orc_mod.oracle(), _ref_or_skip(orc_mod, order)
    ng = orc_mod.nghosts_for(order)
    g = orc_mod.Grid.make(DIMS[dim], ng)
    kw = dict(PUSH_CASES[case])
    pbc = kw.pop("pbc", "periodic")
    code = dict(periodic=orc_mod.PBC_PERIODIC, absorb=orc_mod.PBC_ABSORB,
                reflect=orc_mod.PBC_REFLECT, none=orc_mod.PBC_NONE)[pbc]
    dx = 0.5
    ctx = orc_mod.make_pusher(dt=0.45 * dx, omegaB0=0.7, mass=1.0, charge=-1.0, dx=dx,
                              xmin=[0.1, 0.2, 0.3], pbc=[code] * 6, **kw)
    em = smooth_fields(g, 10 + dim, amp=0.6)
    n = 4000
    pa = random_particles(g, n, 100 + case, umag=2.0, dead_frac=0.05)
    pb = pa.copy()
    ja = np.zeros(g.shape(3), np.float32)
    jb = np.zeros(g.shape(3), np.float32)
    for step in range(3):
        orc.push(g, order, ctx, pa, n, em)
        ref.push(g, order, ctx, pb, n, em)
        assert_prtls_equal(pa, pb, what=f"push step {step}")
        if pbc == "none":
            # particles outside the box would deposit out of bounds: mark them dead, as
            # migration does in the reference before the next deposit
            for p in (pa, pb):
                out = np.zeros(n, bool)
                for a, nm in enumerate(["i1", "i2", "i3"][:dim]):
                    out |= (getattr(p, nm) < 0) | (getattr(p, nm) >= g.n[a])
                p.tag[out] = 0
        orc.deposit(g, order, pa, n, -1.0, ctx.dt, dx, ja)
        ref.deposit(g, order, pb, n, -1.0, ctx.dt, dx, jb)
        assert_bits_equal(ja, jb, f"deposit step {step}")
    assert np.abs(ja).max() > 0


def test_send_tags(orc_mod):
    """mpi::SendTag values (src/global/arch/mpi_tags.h:45-233) for leaving particles."""
    orc = orc_mod.oracle()
    g = orc_mod.Grid.make((4, 4), 2)
    ctx = orc_mod.make_pusher(dt=0.5, dx=1.0, pbc=[orc_mod.PBC_NONE] * 6, tag_outgoing=1,
                              pusher_flags=2, omegaB0=1.0)
    em = np.zeros(g.shape(6), np.float32)
    p = orc_mod.ParticleSet(9)
    p.tag[:] = 1
    p.weight[:] = 1
    # one particle per direction incl. staying; speed ~c along the direction
    dirs = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 0), (0, 1), (1, -1), (1, 0), (1, 1)]
    for k, (a, b) in enumerate(dirs):
        p.i1[k] = 0 if a < 0 else 3
        p.i2[k] = 0 if b < 0 else 3
        p.dx1[k] = 0.01 if a < 0 else (0.99 if a > 0 else 0.5)
        p.dx2[k] = 0.01 if b < 0 else (0.99 if b > 0 else 0.5)
        if a == 0:
            p.i1[k] = 2
        if b == 0:
            p.i2[k] = 2
        p.ux1[k] = 10.0 * a
        p.ux2[k] = 10.0 * b
    orc.push(g, 0, ctx, p, 9, em)
    assert p.tag.tolist() == [2, 3, 4, 5, 1, 6, 7, 8, 9]
