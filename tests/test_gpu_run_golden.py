"""GPU parity against the RUNNING reference: eb200_srpic_step (strict build, ORDERED deposit =
the serial program order of a one-thread Kokkos-OpenMP run) started from a state dumped by the
reference's own entity.xc and compared with its dumps step by step (tests/golden/run_*.npz).

* stream2d: every step bit for bit (E, B, J, every particle array).
* reconnection_small (MATCH + ABSORB x2 walls, injector output imported from the dump): particle
  counts exact every step; E, B, J within 3e-6 of max|F| and particle arrays within 1e-5 (cell
  indices equal for all but <= 0.1 % face-crossers): the MATCH profile s = tanh(...) is evaluated
  on the device (CUDA tanhf vs glibc tanhf, last ulp)."""
import numpy as np
import pytest

import run_cases as rc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import torch
    import entity_b200 as eb
    from entity_b200 import lib as L
    from entity_b200.srpic import PRTL_DTYPES, Simulation
    return torch, eb, L, PRTL_DTYPES, Simulation


def build_sim(mods, case, z, s0, strict=True, deposit=None, fused=False):
    torch, eb, L, PRTL_DTYPES, Simulation = mods
    c = rc.CASES[case]
    walls = c["walls"]
    fbc = [L.FBC_PERIODIC] * 6
    pbc = [L.PBC_PERIODIC] * 6
    if walls:
        fbc[2] = fbc[3] = L.FBC_NONE
        pbc[2] = pbc[3] = L.PBC_ABSORB
    sim = Simulation(c["n"], c.get("order", 0), rc.scales(case), nfilter=c["nfilter"], strict=strict,
                     fused=fused, deposit_mode=eb.DEPOSIT_ORDERED if deposit is None else deposit,
                     fbc=fbc, pbc=pbc, xmin=(tuple(c["xmin"]) + (0.0,))[:3])
    sim.em.copy_(torch.from_numpy(z[f"s{s0}/em"]))
    sim.cur.copy_(torch.from_numpy(z[f"s{s0}/cur"]))
    for k, pusher in enumerate(c["pushers"]):
        n = int(z[f"s{s0}/sp{k}_npart"][1])
        m, q = z[f"meta/sp{k}_mass_charge"]
        sp = sim.alloc_species(float(m), float(q), c["cap"], pusher)
        for a in rc.PRTL:
            key = f"s{s0}/sp{k}_{a}"
            if key in z.files and a in sp.arrays:
                sp.arrays[a][:n] = torch.from_numpy(z[key]).to(sim.device)
        sp.npart = n
    sim._species_c = None
    sim.step_index = s0 + 1
    sim.time = float(z[f"s{s0}/time"][0]) + float(np.float32(sim.dt))
    if walls:
        tgt = torch.from_numpy(rc.match_target(case, sim.grid)).to(sim.device)
        sim.set_match(rc.match_faces(case, sim.grid), tgt, 63)
    return sim


def import_injected(mods, sim, z, s, s1):
    torch = mods[0]
    for k, sp in enumerate(sim.species):
        npre, n = (int(v) for v in z[f"s{s}/sp{k}_npart"])
        assert sp.npart == npre, f"step {s}: species {k} npart {sp.npart} != {npre}"
        if n > npre:
            for a in rc.PRTL:
                key = f"s{s}/sp{k}_{a}_inj" if s != s1 else f"s{s}/sp{k}_{a}"
                if key in z.files and a in sp.arrays:
                    src = z[key] if s != s1 else z[key][npre:n]
                    sp.arrays[a][npre:n] = torch.from_numpy(src).to(sim.device)
            sp.npart = n
    sim._species_c = None


def test_stream2d_bit_exact(mods):
    z = rc.load("stream2d")
    s0, s1 = (int(v) for v in z["meta/steps"])
    sim = build_sim(mods, "stream2d", z, s0)
    for s in range(s0 + 1, s1 + 1):
        sim.step()
        em, cur = sim.em.cpu().numpy(), sim.cur.cpu().numpy()
        assert np.array_equal(em.view(np.uint32), z[f"s{s}/em"].view(np.uint32)), f"step {s}: E/B"
        assert np.array_equal(cur.view(np.uint32), z[f"s{s}/cur"].view(np.uint32)), f"step {s}: J"
        for k, sp in enumerate(sim.species):
            if rc.CASES["stream2d"]["pushers"][k] == 0:
                continue
            n = sp.npart
            for a in rc.PRTL:
                if a not in sp.arrays:
                    continue
                v = sp.arrays[a][:n].cpu().numpy()
                if f"s{s}/sp{k}_{a}" in z.files:
                    assert np.array_equal(v, z[f"s{s}/sp{k}_{a}"]), f"step {s}: sp{k}.{a}"
                elif f"s{s}/sp{k}_{a}_sum" in z.files:
                    assert rc.checksum(v) == z[f"s{s}/sp{k}_{a}_sum"][0], f"step {s}: sp{k}.{a} checksum"


@pytest.mark.parametrize("mode", ["strict_ordered", "fast_fused"])
def test_reconnection_small_window(mods, mode):
    torch, eb = mods[0], mods[1]
    case = "reconnection_small"
    z = rc.load(case)
    s0, s1 = (int(v) for v in z["meta/steps"])
    strict = mode == "strict_ordered"
    sim = build_sim(mods, case, z, s0, strict=strict,
                    deposit=None if strict else eb.DEPOSIT_AGGREGATED, fused=not strict)
    ftol = 3e-6 if strict else 2e-4  # of max|F|; the fast build contracts to FMA and sums J unordered
    for s in range(s0 + 1, s1 + 1):
        sim.step()
        for nm, a, b in (("E/B", sim.em.cpu().numpy(), z[f"s{s}/em"]),
                         ("J", sim.cur.cpu().numpy(), z[f"s{s}/cur"])):
            tol = ftol * np.abs(b).max()
            assert np.abs(a - b).max() <= tol, f"step {s}: {nm} off by {np.abs(a - b).max():.3e} > {tol:.3e}"
        import_injected(mods, sim, z, s, s1)  # asserts the exact particle counts
    for k, sp in enumerate(sim.species):
        npre = int(z[f"s{s1}/sp{k}_npart"][0])
        tag = sp.arrays["tag"][:npre].cpu().numpy()
        assert np.array_equal(tag, z[f"s{s1}/sp{k}_tag"][:npre]), "absorbed particles differ"
        moved = np.zeros(npre, bool)
        for a in ("i1", "i2"):
            moved |= sp.arrays[a][:npre].cpu().numpy() != z[f"s{s1}/sp{k}_{a}"][:npre]
        assert moved.mean() <= 1e-3, f"{moved.sum()} particles ended in another cell"
        for a in ("dx1", "dx2", "ux1", "ux2", "ux3"):
            v, r = sp.arrays[a][:npre].cpu().numpy(), z[f"s{s1}/sp{k}_{a}"][:npre]
            err = np.abs(v - r)[~moved]
            assert err.max() <= (1e-5 if strict else 2e-4) * max(1.0, np.abs(r).max()), f"sp{k}.{a}: {err.max():.3e}"


@pytest.mark.parametrize("case", ["turbulence2d", "turbulence3d"])
@pytest.mark.parametrize("mode", ["strict_ordered", "fast_fused"])
def test_turbulence_window(mods, case, mode):
    """pgens/turbulence with the antenna's ext_current (eb200_ext_current_t refilled every step
    from the dumped amplitudes), 2D zig-zag and 3D third-order Esirkepov (the fast fused 3D run
    goes through the shared-memory J tile). The antenna term goes through cosf / sinf (CUDA vs
    glibc, last ulp): E, B, J within 3e-6 of max|F| in the strict build (2e-4 fast), particle
    counts exact, final particle arrays within 1e-5 (2e-4 fast)."""
    torch, eb = mods[0], mods[1]
    z = rc.load(case)
    s0, s1 = (int(v) for v in z["meta/steps"])
    strict = mode == "strict_ordered"
    sim = build_sim(mods, case, z, s0, strict=strict,
                    deposit=None if strict else eb.DEPOSIT_AGGREGATED, fused=not strict)
    ftol = 3e-6 if strict else 2e-4
    for s in range(s0 + 1, s1 + 1):
        sim.set_ext_current(rc.antenna_table(case, z, s - 1))
        sim.step()
        for nm, a, b in (("E/B", sim.em.cpu().numpy(), z[f"s{s}/em"]),
                         ("J", sim.cur.cpu().numpy(), z[f"s{s}/cur"])):
            tol = ftol * np.abs(b).max()
            assert np.abs(a - b).max() <= tol, f"step {s}: {nm} off by {np.abs(a - b).max():.3e} > {tol:.3e}"
        for k, sp in enumerate(sim.species):
            assert sp.npart == int(z[f"s{s}/sp{k}_npart"][1])
    dim = len(rc.CASES[case]["n"])
    for k, sp in enumerate(sim.species):
        n = sp.npart
        moved = np.zeros(n, bool)
        for a in ("i1", "i2", "i3")[:dim]:
            moved |= sp.arrays[a][:n].cpu().numpy() != z[f"s{s1}/sp{k}_{a}"][:n]
        assert moved.mean() <= 2e-3, f"{moved.sum()} particles ended in another cell"
        for a in ("dx1", "dx2", "dx3")[:dim] + ("ux1", "ux2", "ux3"):
            v, r = sp.arrays[a][:n].cpu().numpy(), z[f"s{s1}/sp{k}_{a}"][:n]
            err = np.abs(v - r)[~moved]
            assert err.max() <= (1e-5 if strict else 2e-4) * max(1.0, np.abs(r).max()), f"sp{k}.{a}: {err.max():.3e}"
