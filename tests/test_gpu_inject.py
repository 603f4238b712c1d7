"""eb200_inject_nonuniform / eb200_particle_moment on the GPU.

The reference's injector draws from a Kokkos random pool (a different stream per backend and
thread count), so there is no bit-level oracle; what is checked instead:
 * the per-cell pair counts are EXACTLY what the reference's rule gives with this library's
   documented stream (ppc = floor(ppc0 sd) + [u < frac], u = the first Philox4x32-10 draw of the
   cell's stream, oracle/philox.py: pinned to the published known answers in tests/test_inject.py);
 * particles sit in their cells in cell order, both species of a pair at the same position,
   dx in [0, 1), tags alive, weights 1, i_prev / dx_prev = i / dx;
 * velocity statistics of arch::energy_dist::Maxwellian: <u_i^2> = T (non-relativistic),
   isotropy and <gamma> for the relativistic branch, the drift's mean four-velocity;
 * ReplenishUniform: cells below 0.9 target get (target - n) / target * ppc, the others nothing;
 * the density moment against numpy's histogram of the same particles;
 * determinism: same arguments -> identical arrays; another `call` id -> different particles;
 * capacity overflow is an error and injects nothing."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import entity_b200 as eb
    from entity_b200 import lib as L
    from entity_b200.srpic import Scales, Simulation
    return torch, eb, L, Scales, Simulation


def make(mods, n=(64, 48), cap=400000, ppc0=8.0):
    torch, eb, L, Scales, Simulation = mods
    sim = Simulation(n, 0, Scales(len(n), 0.5, larmor0=1.0, skindepth0=1.0, ppc0=ppc0))
    sim.scales["n0"] = sim.scales["ppc0"] / sim.scales["V0"]
    sim.alloc_species(1.0, -1.0, cap)
    sim.alloc_species(1.0, +1.0, cap)
    return sim


def arrays(sim, k):
    sp = sim.species[k]
    return {nm: v[:sp.npart].cpu().numpy() for nm, v in sp.arrays.items()}


def test_counts_positions_and_determinism(mods):
    torch, eb, L, _, _ = mods
    from oracle import philox
    sim = make(mods)
    g = sim.grid
    table = torch.zeros(g.shape(3), device=sim.device)
    jj, ii = np.meshgrid(np.arange(g.n[1]), np.arange(g.n[0]), indexing="ij")
    dens = (0.25 + 0.75 * np.sin(0.2 * ii) ** 2 * np.cos(0.15 * jj) ** 2).astype(np.float32)
    table[1, g.ng:g.ng + g.n[1], g.ng:g.ng + g.n[0]] = torch.from_numpy(dens).to(sim.device)
    ppc = 3.7
    n_inj = sim.inject_nonuniform((0, 1), ppc, L.SDIST_TABLE, table, comp=1,
                                  temperatures=(0.01, 0.02), seed=0xabcdef12345, call=3)
    # exact per-cell counts
    ppc_real = (np.float32(ppc) * dens).astype(np.float32)
    base = ppc_real.astype(np.uint32)
    u = philox.first_uniform(0xabcdef12345, sim.step_index, 3, np.arange(dens.size)).reshape(dens.shape)
    expect = base + (u < (ppc_real - base.astype(np.float32)))
    a, b = arrays(sim, 0), arrays(sim, 1)
    assert n_inj == int(expect.sum()) == a["i1"].size == b["i1"].size
    counts = np.zeros(dens.shape, np.int64)
    np.add.at(counts, (a["i2"], a["i1"]), 1)
    assert np.array_equal(counts, expect)
    key = a["i1"].astype(np.int64) + g.n[0] * a["i2"].astype(np.int64)
    assert (np.diff(key) >= 0).all(), "not in cell order"
    for nm in ("i1", "i2", "dx1", "dx2", "i1_prev", "i2_prev", "dx1_prev", "dx2_prev"):
        assert np.array_equal(a[nm], b[nm]), f"pair members differ in {nm}"
    assert np.array_equal(a["i1"], a["i1_prev"]) and np.array_equal(a["dx2"], a["dx2_prev"])
    for nm in ("dx1", "dx2"):
        assert a[nm].min() >= 0.0 and a[nm].max() < 1.0
        assert abs(a[nm].mean() - 0.5) < 0.01
    assert (a["tag"] == 1).all() and (a["weight"] == 1.0).all()
    assert not np.array_equal(a["ux1"], b["ux1"])
    for arr, T in ((a, 0.01), (b, 0.02)):
        for nm in ("ux1", "ux2", "ux3"):
            assert abs(arr[nm].var() / T - 1.0) < 0.03 and abs(arr[nm].mean()) < 4 * np.sqrt(T / n_inj) + 1e-4
    # determinism / call id
    sim2 = make(mods)
    sim2.inject_nonuniform((0, 1), ppc, L.SDIST_TABLE, table, comp=1, temperatures=(0.01, 0.02),
                           seed=0xabcdef12345, call=3)
    a2 = arrays(sim2, 0)
    assert all(np.array_equal(a[nm], a2[nm]) for nm in a)
    sim3 = make(mods)
    sim3.inject_nonuniform((0, 1), ppc, L.SDIST_TABLE, table, comp=1, temperatures=(0.01, 0.02),
                           seed=0xabcdef12345, call=4)
    assert not np.array_equal(arrays(sim3, 0)["dx1"][:1000], a["dx1"][:1000])


def test_relativistic_and_drifting_maxwellians(mods):
    torch, eb, L, _, _ = mods
    sim = make(mods, n=(128, 96), cap=700000)
    n = sim.inject_nonuniform((0, 1), 20.0, temperatures=(2.0, 1e-3), drifts=((0, 0, 0), (0.0, -3.0, 0.0)))
    a, b = arrays(sim, 0), arrays(sim, 1)
    gam = np.sqrt(1 + a["ux1"].astype(np.float64) ** 2 + a["ux2"] ** 2 + a["ux3"] ** 2)
    # Maxwell-Juttner: <gamma> = 3 T + K1(1/T) / K2(1/T); T = 2: K1(0.5) / K2(0.5) = 0.21939 -> 6.2194
    assert abs(gam.mean() - 6.2194) < 0.02
    for nm in ("ux1", "ux2", "ux3"):
        assert abs(a[nm].mean()) < 0.05  # isotropic
    v = [a[nm].astype(np.float64).var() for nm in ("ux1", "ux2", "ux3")]
    assert max(v) / min(v) < 1.03
    # cold drifting beam along -x2: u2 ~ -3, transverse spread ~ sqrt(T)
    assert abs(b["ux2"].mean() + 3.0) < 0.01 and abs(b["ux1"].mean()) < 1e-3
    assert abs(b["ux1"].var() / 1e-3 - 1.0) < 0.03
    assert n == a["i1"].size


def test_replenish_and_density_moment(mods):
    torch, eb, L, _, _ = mods
    sim = make(mods, ppc0=8.0)
    g = sim.grid
    # fill the left half at full density, the right half at 1/4
    half = g.n[0] // 2
    sim.inject_nonuniform((0, 1), 4.0, range_min=[g.ng, g.ng], range_max=[g.ng + half, g.ng + g.n[1]], call=0)
    sim.inject_nonuniform((0, 1), 1.0, range_min=[g.ng + half, g.ng], range_max=[g.ng + g.n[0], g.ng + g.n[1]], call=1)
    buff = sim.particle_moment(eb.STATS_N, [0, 1], comp=0)
    dens = buff[0, g.ng:g.ng + g.n[1], g.ng:g.ng + g.n[0]].cpu().numpy()
    # against numpy: N = sum over particles of inv_n0 / dx^2 per cell
    a, b = arrays(sim, 0), arrays(sim, 1)
    hist = np.zeros(dens.shape, np.float64)
    for arr in (a, b):
        np.add.at(hist, (arr["i2"], arr["i1"]), 1.0)
    inv = np.float32(1.0 / sim.scales["n0"]) / np.float32(0.25)
    np.testing.assert_allclose(dens, hist * float(inv), rtol=2e-6)
    assert abs(dens[:, :half].mean() - 1.0) < 1e-6 and abs(dens[:, half:].mean() - 0.25) < 1e-6
    n_before = sim.species[0].npart
    n_inj = sim.inject_nonuniform((0, 1), 4.0, L.SDIST_REPLENISH, buff, comp=0, target=1.0, call=2)
    a2 = arrays(sim, 0)
    new_i1 = a2["i1"][n_before:]
    assert (new_i1 >= half).all(), "replenished a cell that was above 0.9 of the target"
    expect = 4.0 * 0.75 * (g.n[0] - half) * g.n[1]
    assert abs(n_inj - expect) < 4 * np.sqrt(expect)
    buff = sim.particle_moment(eb.STATS_N, [0, 1], comp=0)
    assert abs(buff[0, g.ng:g.ng + g.n[1], g.ng + half:g.ng + g.n[0]].mean().item() - 1.0) < 0.02


def test_capacity_overflow(mods):
    torch, eb, L, _, _ = mods
    sim = make(mods, cap=1000)
    with pytest.raises(eb.EB200Error, match="maxnpart"):
        sim.inject_nonuniform((0, 1), 4.0)
    assert sim.species[0].npart == 0 and sim.species[1].npart == 0
