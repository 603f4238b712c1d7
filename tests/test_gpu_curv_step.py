"""Curvilinear SRPIC whole-step parity against the RUNNING reference (entity.xc, pgens/
magnetosphere with the dump wrapper; tests/golden/run_magnetosphere_small.npz): eb200_srpic_step
on a qspherical context started from the state after the reference's step s0, over a window of
12 steps.

Covers, in the order of SRPICEngine::step_forward: curvilinear Faraday / Ampere /
CurrentsAmpere, the Boris + GCA pusher with the atmosphere's gravity and the axis / absorbing
particle boundaries, the curvilinear deposit with weights, the spherical filter, and
srpic::FieldBoundaries with its ATMOSPHERE (EnforcedBoundaries towards pgen.AtmFields), AXIS and
MATCH (towards pgen.MatchFields) faces. The particles srpic::ParticleInjector appends every
step come from the dump (the injector is the host's).

Metric functions go through expf / logf / sinf / cosf (CUDA vs glibc, last ulp): E, B within
2e-5 and J within 1e-3 of max|F| (fp32 atomics vs the serial order) at every step, particle counts
exact; particle offsets / momenta within 1e-3 after the 12 steps."""
import numpy as np
import pytest

import run_cases as rc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import entity_b200 as eb
    from entity_b200 import lib as L
    from entity_b200.srpic import Simulation
    return torch, eb, L, Simulation


def build(mods, case, z, s0):
    torch, eb, L, Simulation = mods
    c = rc.SPH_CASES[case]
    f32 = np.float32
    scl = lambda k: float(z[f"meta/{k}"][0])
    mp = [c["extent"][0], c["extent"][1], 0.0, float(f32(np.pi)), c["r0"], c["h"], 0.0]
    scales = dict(dt=scl("algorithms.timestep.dt"), omegaB0=scl("scales.omegaB0"), q0=scl("scales.q0"),
                  B0=scl("scales.B0"), V0=scl("scales.V0"), n0=scl("scales.n0"),
                  ppc0=scl("particles.ppc0"), correction=scl("algorithms.timestep.correction"))
    fbc = [L.FBC_ATMOSPHERE, L.FBC_MATCH, L.FBC_AXIS, L.FBC_AXIS, 0, 0]
    # inside the pusher a PrtlBC::ATMOSPHERE face absorbs (context.h:140-143: is_absorb covers it)
    pbc = [L.PBC_ABSORB, L.PBC_ABSORB, L.PBC_AXIS, L.PBC_AXIS, 0, 0]
    sim = Simulation(c["n"], 0, scales, nfilter=c["nfilter"], strict=False, fused=False,
                     deposit_mode=eb.DEPOSIT_ATOMIC, fbc=fbc, pbc=pbc, metric=L.METRIC_QSPHERICAL,
                     metric_params=mp)
    sim.em.copy_(torch.from_numpy(z[f"s{s0}/em"]))
    sim.cur.copy_(torch.from_numpy(z[f"s{s0}/cur"]))
    match, atm = rc.sph_geometry(case, sim.grid.ng)
    tm = torch.from_numpy(z["meta/target_match"]).to(sim.device)
    ta = torch.from_numpy(z["meta/target_atm"]).to(sim.device)
    sim.set_field_bcs([
        dict(kind=L.FBC_ATMOSPHERE, o=0, sign=-1, target=ta, mask=int(z["meta/target_atm_mask"][0]),
             range_min=atm["range_min"], range_max=atm["range_max"], i_edge=atm["i_edge"]),
        dict(kind=L.FBC_MATCH, o=0, sign=+1, target=tm, mask=int(z["meta/target_match_mask"][0]),
             range_min=match["range_min"], range_max=match["range_max"], xg_edge=match["xg_edge"],
             ds=c["match_ds"]),
    ])
    sim.set_gca(scl("algorithms.gca.larmor_max"), scl("algorithms.gca.e_ovr_b_max"))
    sim.set_atmosphere((-scl("grid.boundaries.atmosphere.g"), 0.0, 0.0), atm["x_surf"],
                       scl("grid.boundaries.atmosphere.ds"))
    for k, pusher in enumerate(c["pushers"]):
        n = int(z[f"s{s0}/sp{k}_npart"][1])
        m, q = z[f"meta/sp{k}_mass_charge"]
        sp = sim.alloc_species(float(m), float(q), c["cap"], pusher)
        for a in rc.PRTL + ["phi"]:
            key = f"s{s0}/sp{k}_{a}"
            if key in z.files and a in sp.arrays:
                sp.arrays[a][:n] = torch.from_numpy(z[key]).to(sim.device)
        sp.npart = n
    sim._species_c = None
    sim.step_index = s0 + 1
    sim.time = float(z[f"s{s0}/time"][0]) + scales["dt"]
    return sim


def import_injected(mods, sim, z, s, s1):
    torch = mods[0]
    for k, sp in enumerate(sim.species):
        npre, n = (int(v) for v in z[f"s{s}/sp{k}_npart"])
        assert sp.npart == npre, f"step {s}: species {k} npart {sp.npart} != {npre}"
        if n > npre:
            for a in rc.PRTL + ["phi"]:
                key = f"s{s}/sp{k}_{a}_inj" if s != s1 else f"s{s}/sp{k}_{a}"
                if key in z.files and a in sp.arrays:
                    src = z[key] if s != s1 else z[key][npre:n]
                    sp.arrays[a][npre:n] = torch.from_numpy(src).to(sim.device)
            sp.npart = n
    sim._species_c = None


def test_magnetosphere_window(mods):
    case = "magnetosphere_small"
    z = rc.load(case)
    s0, s1 = (int(v) for v in z["meta/steps"])
    sim = build(mods, case, z, s0)
    for s in range(s0 + 1, s1 + 1):
        sim.step()
        for nm, tol_rel in (("em", 2e-5), ("cur", 1e-3)):
            a, b = getattr(sim, nm).cpu().numpy(), z[f"s{s}/{nm}"]
            m = np.isfinite(b)
            assert np.array_equal(np.isfinite(a), m), f"step {s}: {nm} non-finite pattern"
            err, tol = np.abs(a[m] - b[m]).max(), tol_rel * np.abs(b[m]).max()
            assert err <= tol, f"step {s}: {nm} off by {err:.3e} > {tol:.3e}"
        import_injected(mods, sim, z, s, s1)
    for k, sp in enumerate(sim.species):
        npre = int(z[f"s{s1}/sp{k}_npart"][0])
        assert np.array_equal(sp.arrays["tag"][:npre].cpu().numpy(), z[f"s{s1}/sp{k}_tag"][:npre])
        moved = np.zeros(npre, bool)
        for a in ("i1", "i2"):
            moved |= sp.arrays[a][:npre].cpu().numpy() != z[f"s{s1}/sp{k}_{a}"][:npre]
        assert moved.mean() <= 1e-2, f"{moved.sum()} particles ended in another cell"
        for a in ("dx1", "dx2", "ux1", "ux2", "ux3", "phi"):
            v, r = sp.arrays[a][:npre].cpu().numpy(), z[f"s{s1}/sp{k}_{a}"][:npre]
            err = np.abs(v - r)[~moved]
            if a == "phi":
                # an angle: 0 and 2 pi are the same place (atan2 decides the branch by rounding)
                err = np.minimum(err, np.abs(np.float32(2 * np.pi) - err))
            # 12 steps of a strongly magnetised plasma (larmor0 = 2e-5): the last-ulp differences of
            # expf / logf / sincosf between CUDA and glibc grow with the window (5 steps: 2e-4)
            assert err.max() <= 1e-3 * max(1.0, np.abs(r).max()), f"sp{k}.{a}: {err.max():.3e}"
