"""Whole-step parity of eb200_srpic_step against the oracle stepping the same imported state
(never regenerate, always import the initial state)."""
import numpy as np
import pytest

from helpers import assert_values_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import entity_b200
    from entity_b200 import workloads
    from oracle import orc, pic
    orc.build()
    return entity_b200, workloads, orc, pic


def _compare_particles(sim, osim, exact):
    for sp, osp in zip(sim.species, osim.species):
        assert sp.npart == osp["npart"]
        n = sp.npart
        for nm in osp["prtls"].names():
            if nm not in sp.arrays:
                continue
            a = sp.arrays[nm][:n].cpu().numpy()
            b = getattr(osp["prtls"], nm)[:n]
            if exact:
                assert np.array_equal(a, b), f"{nm}: {(a != b).sum()} of {n} differ"
            elif a.dtype.kind == "f":
                np.testing.assert_allclose(a, b, rtol=2e-3, atol=2e-4, err_msg=nm)


@pytest.mark.parametrize("case", ["two_stream_2d", "two_stream_1d", "turbulence_3d_o3",
                                  "reconnection_2d", "two_stream_2d_o2"])
def test_step_strict_exact(mods, case):
    """strict-fp + ordered deposit: particles, J and E/B bit-identical to the serial oracle
    over a window of steps."""
    eb, wl, orc, pic = mods
    kw = dict(strict=True, deposit_mode=eb.DEPOSIT_ORDERED)
    if case == "two_stream_2d":
        sim = wl.two_stream((48, 32), ppc0=16, **kw)
    elif case == "two_stream_1d":
        sim = wl.two_stream((96,), ppc0=16, **kw)
    elif case == "two_stream_2d_o2":
        sim = wl.two_stream((32, 32), ppc0=8, **kw)
        # same plasma, different shape order: rebuild with order 2
        from entity_b200.srpic import Scales, Simulation
        s2 = Simulation((32, 32), 2, Scales(2, sim.ctx.dx, 100.0, 10.0, 8), nfilter=2, **kw)
        for sp in sim.species:
            s2.add_species(sp.mass, sp.charge, sp.arrays, sp.npart, sp.pusher)
        sim = s2
    elif case == "turbulence_3d_o3":
        sim = wl.turbulence((12, 10, 8), ppc0=4, order=3, **kw)
    else:
        sim = wl.reconnection((64, 64), ppc0=8, nfilter=3, **kw)
    osim = pic.from_device_sim(sim)
    nsteps = 6
    for step in range(nsteps):
        sim.step()
        osim.step()
        assert_values_equal(sim.cur.cpu().numpy(), osim.cur, f"{case}: J at step {step}")
        assert_values_equal(sim.em.cpu().numpy(), osim.em, f"{case}: EM at step {step}")
    _compare_particles(sim, osim, exact=True)
    assert np.abs(osim.cur).max() > 0


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("agg", [False, True])
def test_step_fast_tolerance(mods, fused, agg):
    """fast-fp (FMA contraction, atomic deposit, optionally fused push+deposit, sorting on):
    particle counts identical, fields/currents/energy within fp32 tolerance over the window."""
    eb, wl, orc, pic = mods
    sim = wl.two_stream((64, 48), ppc0=16, fused=fused, sort_interval=2,
                        deposit_mode=eb.DEPOSIT_AGGREGATED if agg else eb.DEPOSIT_ATOMIC)
    osim = pic.from_device_sim(sim)
    for step in range(8):
        sim.step()
        osim.step()
    assert sim.n_pushed() == osim.n_pushed()
    j, jo = sim.cur.cpu().numpy(), osim.cur
    e, eo = sim.em.cpu().numpy(), osim.em
    # stated tolerances: 5e-4 of the field maximum (fp32 sums of ~1e2 contributions per node,
    # reordered by atomics, accumulated over 8 steps)
    assert np.abs(j - jo).max() <= 5e-4 * np.abs(jo).max()
    assert np.abs(e - eo).max() <= 5e-4 * max(np.abs(eo).max(), 1e-30)
    g = osim.grid
    sl = (slice(None),) + tuple(slice(g.ng, g.ng + g.n[a]) for a in range(g.dim))[::-1]
    en = float((e[sl].astype(np.float64) ** 2).sum())
    eno = float((eo[sl].astype(np.float64) ** 2).sum())
    assert abs(en - eno) <= 1e-3 * eno


def _energies(em, grid, species):
    """(field, particle) energy in code units: sum over active cells of (E^2 + B^2) / 2 and
    sum over alive particles of w (gamma - 1), both accumulated in float64."""
    g = grid
    sl = (slice(None),) + tuple(slice(g.ng, g.ng + g.n[a]) for a in range(g.dim))[::-1]
    fld = 0.5 * float((em[sl].astype(np.float64) ** 2).sum())
    kin = 0.0
    for ux, uy, uz, w, tag in species:
        alive = tag == 1
        u2 = (ux[alive].astype(np.float64) ** 2 + uy[alive].astype(np.float64) ** 2 +
              uz[alive].astype(np.float64) ** 2)
        kin += float((w[alive].astype(np.float64) * (np.sqrt(1.0 + u2) - 1.0)).sum())
    return fld, kin


@pytest.mark.parametrize("which", [0, 2, 5, 6, 7])
def test_total_energy_window(mods, which):
    """north_star: 'fields, currents and total energy stay within a stated fp32 tolerance over a
    fixed step window'. Production path (fast build, fused push+deposit, warp-aggregated deposit,
    periodic re-sorting) on the reconnection workload, 20 steps, against the serial oracle on the
    imported state. Stated tolerances: field energy 1e-3, particle kinetic energy 1e-5, total
    energy 1e-4 (relative); particle counts identical at every step."""
    eb, wl, orc, pic = mods
    sim = wl.reconnection((128, 64), ppc0=8, nfilter=4, fused=True, sort_interval=5,
                          deposit_mode=eb.DEPOSIT_AGGREGATED)
    sim.ctx.set_pd_kernel(which)
    osim = pic.from_device_sim(sim)

    def dev_energy():
        sp = [tuple(s.arrays[k][:s.npart].cpu().numpy() for k in ("ux1", "ux2", "ux3", "weight", "tag"))
              for s in sim.species]
        return _energies(sim.em.cpu().numpy(), osim.grid, sp)

    def orc_energy():
        sp = [tuple(getattr(s["prtls"], k)[:s["npart"]] for k in ("ux1", "ux2", "ux3", "weight", "tag"))
              for s in osim.species]
        return _energies(osim.em, osim.grid, sp)

    f0, k0 = orc_energy()
    for step in range(20):
        sim.step()
        osim.step()
        assert sim.n_pushed() == osim.n_pushed(), f"particle count differs at step {step}"
    (fd, kd), (fo, ko) = dev_energy(), orc_energy()
    assert abs(fd - fo) <= 1e-3 * fo
    assert abs(kd - ko) <= 1e-5 * ko
    assert abs((fd + kd) - (fo + ko)) <= 1e-4 * (fo + ko)
    # and the window did something: energy moved between fields and particles
    assert abs(ko - k0) > 1e-6 * k0 or abs(fo - f0) > 1e-6 * f0


@pytest.mark.parametrize("chunk", ["1024", None])
def test_step_host_streamed_matches_device(mods, chunk, monkeypatch):
    """eb200_srpic_step_host (host buffers; particle chunks streamed up / pushed / streamed down
    on three streams, plain whole-array path on sort steps) against eb200_srpic_step on device
    arrays from the same state. The pusher is independent of the deposit order, so the first
    step's particles are bit-identical; J differs by the order of fp32 atomic additions (5e-4
    of max|J|), which feeds back into later steps (fp32 tolerance there)."""
    eb, wl, orc, pic = mods
    if chunk:
        monkeypatch.setenv("EB200_HOST_CHUNK", chunk)
    else:
        monkeypatch.delenv("EB200_HOST_CHUNK", raising=False)
    sim = wl.reconnection((64, 64), ppc0=8, nfilter=3, strict=True, fused=True, sort_interval=3,
                          deposit_mode=eb.DEPOSIT_AGGREGATED)
    hs = sim.host_state()
    # poison what the streamed path promises not to read: stale i_prev / dx_prev on the host
    for spec in hs["species"]:
        for nm in ("i1_prev", "i2_prev", "dx1_prev", "dx2_prev"):
            spec[nm].fill_(7)
    names = ["i1", "i2", "dx1", "dx2", "ux1", "ux2", "ux3", "weight", "i1_prev", "i2_prev",
             "dx1_prev", "dx2_prev", "tag"]
    for step in range(6):
        sim.step()
        up, down = sim.step_host(hs)
        assert up > 0 and down > 0
        j, jh = sim.cur.cpu().numpy(), hs["cur"].numpy()
        assert np.abs(j - jh).max() <= 5e-4 * np.abs(j).max(), f"J at step {step}"
        e, eh = sim.em.cpu().numpy(), hs["em"].numpy()
        assert np.abs(e - eh).max() <= 1e-4 * np.abs(e).max(), f"EM at step {step}"
        for sp, spec, c in zip(sim.species, hs["species"], hs["c"]):
            assert sp.npart == c.npart
            n = sp.npart
            if step == 0:
                for nm in names:
                    a, b = sp.arrays[nm][:n].cpu().numpy(), spec[nm][:n].numpy()
                    assert np.array_equal(a, b), f"{nm}: {(a != b).sum()} of {n} differ"
    # after the window (two sort steps inside): same multiset of particles within fp32 tolerance
    for sp, spec in zip(sim.species, hs["species"]):
        n = sp.npart
        assert np.array_equal(np.sort(sp.arrays["weight"][:n].cpu().numpy()), np.sort(spec["weight"][:n].numpy()))
        ua, ub = sp.arrays["ux1"][:n].cpu().numpy(), spec["ux1"][:n].numpy()
        assert abs(float(ua.astype(np.float64).sum()) - float(ub.astype(np.float64).sum())) <= 1e-3 * np.abs(ua).sum()


@pytest.mark.parametrize("host", [False, True])
def test_lean_prev(mods, host):
    """eb200_set_lean_prev: the fused kernel does not store i*_prev / dx*_prev, the sort leaves
    them out, the host step moves them in neither direction -- and nothing else changes. Two
    fast-build simulations from the same state (stable sort in both, so that the arrays can be
    compared element by element): particles bit-identical after the first step, fields / currents
    within the fp32 order-of-additions tolerance over the window; the lean run's prev arrays keep
    the poison they were given."""
    eb, wl, orc, pic = mods
    kw = dict(ppc0=8, nfilter=2, strict=False, fused=True, sort_interval=2, deposit_mode=eb.DEPOSIT_AGGREGATED)
    ref, lean = wl.reconnection((128, 64), **kw), wl.reconnection((128, 64), **kw)
    for sim in (ref, lean):
        sim.ctx.set_sort_mode(0)
    lean.ctx.set_lean_prev(True)
    prev = ("i1_prev", "i2_prev", "dx1_prev", "dx2_prev")
    hs = lean.host_state() if host else None
    for sp in lean.species:
        for nm in prev:
            sp.arrays[nm].fill_(7)
    if host:
        for spec in hs["species"]:
            for nm in prev:
                spec[nm].fill_(7)
    names = ["i1", "i2", "dx1", "dx2", "ux1", "ux2", "ux3", "weight", "tag"]
    for step in range(5):
        ref.step()
        if host:
            lean.step_host(hs)
            cur, em = hs["cur"].numpy(), hs["em"].numpy()
            arrays = [(c.npart, spec) for c, spec in zip(hs["c"], hs["species"])]
        else:
            lean.step()
            cur, em = lean.cur.cpu().numpy(), lean.em.cpu().numpy()
            arrays = [(sp.npart, sp.arrays) for sp in lean.species]
        j = ref.cur.cpu().numpy()
        assert np.abs(j - cur).max() <= 5e-4 * np.abs(j).max(), f"J at step {step}"
        e = ref.em.cpu().numpy()
        assert np.abs(e - em).max() <= 1e-4 * np.abs(e).max(), f"EM at step {step}"
        for sp, (n, arr) in zip(ref.species, arrays):
            assert sp.npart == n
            if step == 0:
                for nm in names:
                    a, b = sp.arrays[nm][:n].cpu().numpy(), arr[nm][:n].cpu().numpy()
                    assert np.array_equal(a, b), f"{nm}: {(a != b).sum()} of {n} differ"
    # groups of four particles go through the fused kernel; a tail of < 4 takes the generic one
    for n, arr in arrays:
        m = n // 4 * 4
        for nm in prev:
            assert float((arr[nm][:m].float() - 7).abs().max()) == 0.0, f"{nm} was written"
    for sp in ref.species:
        assert float((sp.arrays["i1_prev"][:sp.npart].float() - 7).abs().min()) >= 0  # (reference run: real values)
        assert not bool((sp.arrays["dx1_prev"][:sp.npart] == 7).all())


def test_step_with_match_boundaries(mods):
    """Whole steps with the reconnection configuration's boundaries in x2 (fields MATCH,
    particles ABSORB; pgens/reconnection/reconnection.toml) against the oracle stepper with the
    same MATCH layers (oracle/bcs.py, pinned to the reference's MatchBoundaries_kernel): strict
    build + ordered deposit. Everything but tanh is the same fp32 operation sequence: fields and
    currents within 2e-6 of their maxima, particle counts (absorption) identical every step."""
    eb, wl, orc, pic = mods
    import torch
    from entity_b200.srpic import Scales, Simulation
    n, G, dx = (48, 40), 2, 0.25
    fbc = [eb.FBC_PERIODIC, eb.FBC_PERIODIC, eb.FBC_NONE, eb.FBC_NONE, 0, 0]
    pbc = [eb.PBC_PERIODIC, eb.PBC_PERIODIC, eb.PBC_ABSORB, eb.PBC_ABSORB, 0, 0]
    base = wl.two_stream(n, ppc0=8, strict=True, deposit_mode=eb.DEPOSIT_ORDERED)
    sim = Simulation(n, 0, Scales(2, base.ctx.dx, 100.0, 10.0, 8), nfilter=2, strict=True,
                     deposit_mode=eb.DEPOSIT_ORDERED, fbc=fbc, pbc=pbc, xmin=(0.0, -5.0, 0.0))
    for sp in base.species:
        sp.arrays["ux2"].mul_(30.0)  # fast enough along x2 for some to reach the absorbing walls
        sim.add_species(sp.mass, sp.charge, sp.arrays, sp.npart, sp.pusher)
    dx = sim.ctx.dx
    # target: uniform guide field B_x1 = 0.2 (tetrad), nothing else; both x2 faces, 6 cells thick
    target = torch.zeros_like(sim.em)
    target[3] = 0.2
    sim.em[3] = 0.2 / dx
    nds, ext = 6, [n[0] + 2 * G, n[1] + 2 * G]
    ymin, ymax = -5.0, float(np.float32(-5.0) + np.float32(dx) * np.float32(n[1]))
    faces = [(1, ymin, nds * dx, [0, 0], [ext[0], G + nds]),
             (1, ymax, nds * dx, [0, G + n[1] - nds], [ext[0], ext[1]])]
    sim.set_match(faces, target, 63)
    osim = pic.from_device_sim(sim)
    n0 = sum(s.npart for s in sim.species)
    for step in range(8):
        sim.step()
        osim.step()
        alive = sum(int((s.arrays["tag"][:s.npart] == 1).sum()) for s in sim.species)
        oalive = sum(int((s["prtls"].tag[:s["npart"]] == 1).sum()) for s in osim.species)
        assert alive == oalive, f"alive particles differ at step {step}"
        e, eo = sim.em.cpu().numpy(), osim.em
        j, jo = sim.cur.cpu().numpy(), osim.cur
        assert np.abs(e - eo).max() <= 2e-6 * np.abs(eo).max(), f"EM at step {step}"
        assert np.abs(j - jo).max() <= 2e-6 * np.abs(jo).max() + 1e-12, f"J at step {step}"
    assert alive < n0, "no particle reached the absorbing walls: the case does not test them"
