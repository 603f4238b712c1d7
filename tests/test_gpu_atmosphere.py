"""srpic::ParticleInjector for an ATMOSPHERE face (eb200_atmosphere_particles, curvilinear
eb200_inject_nonuniform / eb200_particle_moment) against the RUNNING reference
(entity.xc, pgens/magnetosphere with the dump wrapper: tests/golden/run_magnetosphere_small.npz).

The reference draws from a Kokkos random pool, so the injected particles themselves cannot
match; what the dump pins:
 * the rule: with OUR density moment of the dumped pre-injection state, the reference's injected
   particles sit only in cells where ppc_real = nmax ppc0 / 2 * Replenish(AtmosphereDensityProfile)
   is positive, never more than floor(ppc_real) + 1 per cell, and their total over the window is
   the expectation sum(ppc_real) within 4 sigma of the stochastic rounding;
 * our per-cell counts are EXACTLY floor(ppc_real) + [u < frac] with the documented Philox
   stream (oracle/philox.py);
 * weights: the reference's injected weights equal ours for the same cell (sqrt_det_h(centre) /
   V0) to 2e-5;
 * velocities: ours, taken back to the tetrad basis at the cell centre, are a Maxwellian of the
   atmosphere's temperature (<v_i^2> = T to 3 %), as are the reference's (to 25 %: 293 pairs);
   phi = 0, both species of a pair at the same position;
 * the step with the injector registered keeps the particle count on the reference's curve."""
import numpy as np
import pytest

import run_cases as rc
from test_gpu_curv_step import build, mods  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu
f32 = np.float32
CASE = "magnetosphere_small"


def atm_dict(z, ng):
    _, atm = rc.sph_geometry(CASE, ng)
    scl = lambda k: float(z[f"meta/{k}"][0])
    return dict(dim=0, sign=-1, x_surf=atm["x_surf"], ds=scl("grid.boundaries.atmosphere.ds"),
                height=scl("grid.boundaries.atmosphere.height"),
                temperature=scl("grid.boundaries.atmosphere.temperature"),
                density=scl("grid.boundaries.atmosphere.density"), species=(0, 1))


def geometry():
    c = rc.SPH_CASES[CASE]
    n1, n2 = c["n"]
    r0 = f32(c["r0"])
    chi_min = np.log(f32(c["extent"][0]) - r0)
    dchi = (np.log(f32(c["extent"][1]) - r0) - chi_min) / f32(n1)
    deta = f32(np.pi) / f32(n2)
    return n1, n2, r0, chi_min, dchi, deta


def expected_ppc(dens, a, ppc0):
    """ppc_real per active cell [n2, n1] (fp32): injectors.hpp:616-624 with spatial_dist.h:56-80 and
    particle_injector.h:171-180"""
    n1, n2, r0, chi_min, dchi, _ = geometry()
    r = (r0 + np.exp((np.arange(n1, dtype=f32) + f32(0.5)) * dchi + chi_min)).astype(f32)
    xs, h, nmax = f32(a["x_surf"]), f32(a["height"]), f32(a["density"])
    tgt = np.where((r < xs) | (r >= xs + f32(a["ds"])), f32(0),
                   nmax * np.exp(-(xs / h) * (f32(1) - xs / r))).astype(f32)[None, :]
    sd = np.where(f32(0.9) * tgt > dens, (tgt - dens) / nmax, f32(0)).astype(f32)
    return (nmax * f32(ppc0) * f32(0.5) * sd).astype(f32)


def load_pre(mods_, z, s):
    """state after the reference's step s without the particles its injector appended"""
    sim = build(mods_, CASE, z, s)
    for k, sp in enumerate(sim.species):
        sp.npart = int(z[f"s{s}/sp{k}_npart"][0])
    sim._species_c = None
    return sim


def cell_counts(i1, i2, n1, n2):
    cnt = np.zeros((n2, n1), np.int64)
    np.add.at(cnt, (i2, i1), 1)
    return cnt


def test_rule_counts_weights_against_running_reference(mods):
    """the last dumped step holds the full arrays: [0, npre) is the state the reference's injector saw"""
    torch = mods[0]
    from oracle import philox
    z = rc.load(CASE)
    s = int(z["meta/steps"][1])
    n1, n2 = rc.SPH_CASES[CASE]["n"]
    ppc0 = float(z["meta/particles.ppc0"][0])
    sim = load_pre(mods, z, s)
    g = sim.grid
    a = atm_dict(z, g.ng)
    pre = [sp.npart for sp in sim.species]
    sim.step_index = s
    n_inj = sim.atmosphere_particles(a)
    dens = sim.buff[0, g.ng:g.ng + n2, g.ng:g.ng + n1].cpu().numpy()
    ppc_real = expected_ppc(dens, a, ppc0)
    assert ppc_real.max() > 0
    # ours: exact rule with the documented stream (cell index = position in the active range)
    t = np.arange(n1 * n2, dtype=np.uint32)
    u = philox.first_uniform(0x123456789abcdef0, s, 0x41544d, t).reshape(n2, n1)
    want = ppc_real.astype(np.uint32) + (u < ppc_real - ppc_real.astype(np.uint32).astype(f32))
    sp0, sp1 = sim.species
    ours = cell_counts(sp0.arrays["i1"][pre[0]:sp0.npart].cpu().numpy(),
                       sp0.arrays["i2"][pre[0]:sp0.npart].cpu().numpy(), n1, n2)
    assert n_inj == int(want.sum()) == sp0.npart - pre[0] == sp1.npart - pre[1]
    assert np.array_equal(ours, want)
    for nm in ("i1", "i2", "dx1", "dx2", "weight", "phi"):
        assert torch.equal(sp0.arrays[nm][pre[0]:sp0.npart], sp1.arrays[nm][pre[1]:sp1.npart]), nm
    assert float(sp0.arrays["phi"][pre[0]:sp0.npart].abs().max()) == 0.0
    # the reference's injected tail under the same rule
    npre, n = (int(v) for v in z[f"s{s}/sp0_npart"])
    assert npre == pre[0] and n > npre
    ri1, ri2 = z[f"s{s}/sp0_i1"][npre:n], z[f"s{s}/sp0_i2"][npre:n]
    ref = cell_counts(ri1, ri2, n1, n2)
    near = np.abs(ppc_real - np.round(ppc_real)) < 1e-3  # density rounding next to an integer
    assert (ref[(ppc_real <= 0)] == 0).all(), "reference injected where the rule gives nothing"
    assert (ref <= np.floor(ppc_real) + 1 + near).all() and (ref >= np.floor(ppc_real) - near).all()
    frac = ppc_real - np.floor(ppc_real)
    sigma = np.sqrt(float((frac * (1 - frac)).sum()))
    assert abs((n - npre) - float(ppc_real.sum())) <= 4.0 * sigma + 1.0, (n - npre, ppc_real.sum(), sigma)
    # weights per cell: the reference's vs ours
    wmap = {}
    oi1 = sp0.arrays["i1"][pre[0]:sp0.npart].cpu().numpy()
    oi2 = sp0.arrays["i2"][pre[0]:sp0.npart].cpu().numpy()
    ow = sp0.arrays["weight"][pre[0]:sp0.npart].cpu().numpy()
    for i, j, w in zip(oi1, oi2, ow):
        wmap[(int(i), int(j))] = float(w)
    rw = z[f"s{s}/sp0_weight"][npre:n]
    hits = 0
    for i, j, w in zip(ri1, ri2, rw):
        if (int(i), int(j)) in wmap:
            hits += 1
            assert abs(wmap[(int(i), int(j))] - float(w)) <= 2e-5 * abs(float(w))
    assert hits > 0


def test_injected_totals_over_the_window(mods):
    """every step of the window: the expectation sum(ppc_real) from OUR density of the evolving
    state (the reference's injected particles imported from the dump) against the number the
    reference injected; the rounding noise of one step is sqrt(sum f (1 - f))"""
    from entity_b200 import lib as L
    from test_gpu_curv_step import import_injected
    z = rc.load(CASE)
    s0, s1 = (int(v) for v in z["meta/steps"])
    n1, n2 = rc.SPH_CASES[CASE]["n"]
    ppc0 = float(z["meta/particles.ppc0"][0])
    sim = build(mods, CASE, z, s0)
    g = sim.grid
    a = atm_dict(z, g.ng)
    tot_ref, tot_exp, var = 0, 0.0, 0.0
    for s in range(s0 + 1, s1 + 1):
        sim.step()
        dens = sim.particle_moment(L.STATS_RHO, [0, 1], comp=0, use_weights=True)
        ppc_real = expected_ppc(dens[0, g.ng:g.ng + n2, g.ng:g.ng + n1].cpu().numpy(), a, ppc0)
        frac = ppc_real - np.floor(ppc_real)
        npre, n = (int(v) for v in z[f"s{s}/sp0_npart"])
        tot_ref += n - npre
        tot_exp += float(ppc_real.sum())
        var += float((frac * (1 - frac)).sum())
        import_injected(mods, sim, z, s, s1)
    assert tot_ref > 200
    assert abs(tot_ref - tot_exp) <= 4.0 * np.sqrt(var) + 1.0, (tot_ref, tot_exp, var)


def tetrad(u, i2, n2):
    """XYZ -> tetrad at the cell centre, phi = 0 (qspherical.h:412-424 with h = 0)"""
    th = (i2.astype(f32) + f32(0.5)) * (f32(np.pi) / f32(n2))
    st, ct = np.sin(th), np.cos(th)
    return np.stack([u[0] * st + u[2] * ct, u[0] * ct - u[2] * st, u[1]])


def test_velocities_are_maxwellian_in_the_tetrad_basis(mods):
    torch, eb, L, _ = mods
    z = rc.load(CASE)
    s0, s1 = (int(v) for v in z["meta/steps"])
    n1, n2 = rc.SPH_CASES[CASE]["n"]
    sim = load_pre(mods, z, s0 + 1)
    for sp in sim.species:
        sp.npart = 0
    sim._species_c = None
    T = 0.1
    sim2 = sim
    # capacity 4096 per species: 1 pair per cell on 64 x 48 cells = 3072
    n = sim2.inject_nonuniform((0, 1), 1.0, L.SDIST_UNIFORM, temperatures=(T, T), call=7)
    assert n == n1 * n2
    for sp in sim2.species:
        u = np.stack([sp.arrays[k][:n].cpu().numpy() for k in ("ux1", "ux2", "ux3")])
        vt = tetrad(u, sp.arrays["i2"][:n].cpu().numpy(), n2)
        for c in range(3):
            assert abs((vt[c] ** 2).mean() / T - 1.0) < 0.08, (c, (vt[c] ** 2).mean())
            assert abs(vt[c].mean()) < 4 * np.sqrt(T / n)
    # weights = sqrt_det_h(centre) / V0 (qspherical.h sqrt_det_h with h = 0)
    _, _, r0, chi_min, dchi, deta = geometry()
    sp = sim2.species[0]
    i1 = sp.arrays["i1"][:n].cpu().numpy().astype(f32)
    i2 = sp.arrays["i2"][:n].cpu().numpy().astype(f32)
    e = np.exp((i1 + f32(0.5)) * dchi + chi_min)
    w = dchi * deta * e * (r0 + e) ** 2 * np.sin((i2 + f32(0.5)) * deta) / f32(sim2.scales["V0"])
    assert np.allclose(sp.arrays["weight"][:n].cpu().numpy(), w, rtol=2e-5)
    # the reference's injected particles: same distribution (small sample)
    acc = []
    for s in range(s0 + 1, s1 + 1):
        for k in range(2):
            npre, m = (int(v) for v in z[f"s{s}/sp{k}_npart"])
            # injected at the end of step s: not pushed yet (intermediate steps hold only the tail)
            get = (lambda c: z[f"s{s}/sp{k}_{c}"][npre:m]) if s == s1 else (lambda c: z[f"s{s}/sp{k}_{c}_inj"])
            u = np.stack([get(c) for c in ("ux1", "ux2", "ux3")])
            acc.append(tetrad(u, get("i2"), n2))
    vt = np.concatenate(acc, axis=1)
    assert abs((vt ** 2).mean() / T - 1.0) < 0.25, (vt ** 2).mean()


def test_density_moment_curvilinear(mods):
    """Rho of the dumped state: sum over particles of m w inv_n0 / sqrt_det_h(cell centre)"""
    z = rc.load(CASE)
    s1 = int(z["meta/steps"][1])
    n1, n2 = rc.SPH_CASES[CASE]["n"]
    sim = build(mods, CASE, z, s1)
    g = sim.grid
    from entity_b200 import lib as L
    sim.particle_moment(L.STATS_RHO, [0, 1], comp=0, use_weights=True)
    got = sim.buff[0, g.ng:g.ng + n2, g.ng:g.ng + n1].cpu().numpy().astype(np.float64)
    _, _, r0, chi_min, dchi, deta = geometry()
    want = np.zeros((n2, n1))
    for k in range(2):
        n = int(z[f"s{s1}/sp{k}_npart"][1])
        alive = z[f"s{s1}/sp{k}_tag"][:n] != 0
        i1, i2 = z[f"s{s1}/sp{k}_i1"][:n][alive], z[f"s{s1}/sp{k}_i2"][:n][alive]
        e = np.exp((i1 + 0.5) * float(dchi) + float(chi_min))
        sdh = float(dchi) * float(deta) * e * (float(r0) + e) ** 2 * np.sin((i2 + 0.5) * float(deta))
        np.add.at(want, (i2, i1), z[f"s{s1}/sp{k}_weight"][:n][alive] / sdh / float(z["meta/scales.n0"][0]))
    assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max()


def test_step_with_registered_injector_follows_the_reference_count(mods):
    z = rc.load(CASE)
    s0, s1 = (int(v) for v in z["meta/steps"])
    sim = build(mods, CASE, z, s0)
    sim.set_atmosphere_injector(atm_dict(z, sim.grid.ng))
    sim.step(s1 - s0)
    want = int(z[f"s{s1}/sp0_npart"][1])
    got = [sp.npart for sp in sim.species]
    assert got[0] == got[1]
    # 293 pairs injected by the reference over the window; the rounding noise is ~ sqrt(293)
    assert abs(got[0] - want) <= 4 * np.sqrt(want - int(z[f"s{s0}/sp0_npart"][1])) + 2, (got, want)
    sim.set_atmosphere_injector(None)
