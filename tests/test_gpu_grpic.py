"""GRPIC whole-step parity against the RUNNING reference (entity.xc, pgens/wald and
pgens/accretion built with the dump wrapper; tests/golden/run_*.npz): eb200_grpic_step started
from the state after the reference's step s0 and compared with its dumps step by step.

The Kerr-Schild metric functions go through expf / logf / sinf / cosf / sqrtf (CUDA vs glibc:
last-ulp differences) and every step evaluates them on every cell: fields are compared to
2e-5 of max|F| per array over the window (stated tolerance; no bit-exactness is claimed for
curvilinear metrics)."""
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
    from entity_b200.grpic import GRSimulation
    return torch, eb, L, GRSimulation


def build(mods, case, z, s0):
    torch, eb, L, GRSimulation = mods
    c = rc.GR_CASES[case]
    f32 = np.float32
    metric = {"qkerr_schild": L.METRIC_QKERR_SCHILD, "kerr_schild": L.METRIC_KERR_SCHILD}[c["metric"]]
    mp = [c["extent"][0], c["extent"][1], 0.0, float(f32(np.pi)), c["r0"], c["h"], c["a"]]
    # the scalars the reference's host passes (dumped from its SimulationParams)
    scl = lambda k: float(z[f"meta/{k}"][0])
    dt = scl("algorithms.timestep.dt")
    sim = GRSimulation(c["n"], metric, mp, dt=dt, omegaB0=scl("scales.omegaB0"), q0=scl("scales.q0"),
                       B0=scl("scales.B0"), correction=scl("algorithms.timestep.correction"),
                       nfilter=c["nfilter"], deposit=c["deposit"],
                       pusher_niter=c.get("niter", 10), pusher_eps=c.get("eps", 1e-6))
    for nm in ("em", "em0", "cur", "cur0", "aux"):
        getattr(sim, nm).copy_(torch.from_numpy(z[f"s{s0}/{nm}"]))
    tgt = torch.from_numpy(z["meta/target_init_flds"]).to(sim.device)
    rmin, rmax = rc.gr_match_range(case, sim.grid.ng)
    sim.set_match(tgt, int(z["meta/target_init_flds_mask"][0]), float(f32(c["extent"][1])),
                  c["match_ds"], rmin, rmax)
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
    sim.time = float(z[f"s{s0}/time"][0]) + dt
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


def test_wald_vacuum_window(mods):
    """pgens/wald (BASELINE configs[4]): vacuum Wald solution, qkerr_schild a = 0.95; the GRPIC
    field path (time averages, aux E / H, both Faraday / Ampere sub-steps, SwapFields) with its
    HORIZON, MATCH (towards init_flds) and AXIS boundaries, 12 steps."""
    case = "wald_small"
    z = rc.load(case)
    s0, s1 = (int(v) for v in z["meta/steps"])
    sim = build(mods, case, z, s0)
    for s in range(s0 + 1, s1 + 1):
        sim.step()
        for nm in ("em", "em0", "aux"):
            a, b = getattr(sim, nm).cpu().numpy(), z[f"s{s}/{nm}"]
            assert np.array_equal(np.isfinite(a), np.isfinite(b)), f"step {s}: {nm} non-finite pattern"
            m = np.isfinite(b)
            tol = 2e-5 * np.abs(b[m]).max()
            err = np.abs(a[m] - b[m]).max()
            assert err <= tol, f"step {s}: {nm} off by {err:.3e} > {tol:.3e}"


def test_accretion_window(mods):
    """pgens/accretion: the same black hole with two Boris species injected every step (their
    injector's output imported from the dump): GR pusher (niter = 10), GR deposit into cur0,
    AbsorbCurrents in the MATCH layer, 4 spherical filter passes, TimeAverageJ and both
    AmpereCurrents sub-steps. Particle counts exact every step; D, B (em, em0) within 1e-4 and
    J (cur, cur0) within 1e-3 of max|F| (fp32 atomics sum in a different order than the serial
    reference, and the pusher's metric functions differ from glibc's in the last ulp); final
    particle momenta / offsets within 1e-4 with <= 0.5 % of particles in another cell."""
    case = "accretion_small"
    z = rc.load(case)
    s0, s1 = (int(v) for v in z["meta/steps"])
    sim = build(mods, case, z, s0)
    for s in range(s0 + 1, s1 + 1):
        sim.step()
        for nm, tol_rel in (("em", 1e-4), ("em0", 1e-4), ("cur", 1e-3), ("cur0", 1e-3)):
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
        assert moved.mean() <= 5e-3, f"{moved.sum()} particles ended in another cell"
        for a in ("dx1", "dx2", "ux1", "ux2", "ux3"):
            v, r = sp.arrays[a][:npre].cpu().numpy(), z[f"s{s1}/sp{k}_{a}"][:npre]
            err = np.abs(v - r)[~moved]
            assert err.max() <= 1e-4 * max(1.0, np.abs(r).max()), f"sp{k}.{a}: {err.max():.3e}"
