"""BASELINE configs[1] at its FULL size (reconnection 4096 x 2048 cells, 32 ppc, 3.3e8 particles)
through the C ABI, checked by size-independent properties (the oracle takes minutes there):

 * discrete charge continuity of the fused zig-zag push+deposit: div J + (rho1 - rho0) / dt = 0
   at every node, rho from the bilinear weights (what Sum div J = 0 of the reference's deposit
   test states globally, src/kernels/tests/deposit.cpp), J summed by fp32 atomics: <= 2e-4 of
   max|d rho / dt|;
 * eb200_set_lean_prev changes nothing but the stores of i_prev / dx_prev: particles bit-identical,
   J within the order-of-additions tolerance;
 * the fused kernel against the unfused pair (push kernel, aggregated deposit) validated
   bit-for-bit / to tolerance at oracle sizes in test_gpu_parity.py: same cells for all but
   <= 1e-5 of the particles, momenta to 1e-5, J to 2e-4 of max|J|;
 * one push + deposit of a whole species against the CPU oracle (the compiled reference where it
   travelled): strict build bit for bit, fast fused kernel to the tolerances of the small cases;
 * particle count conserved, nothing non-finite."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = (4096, 2048)
PPC = 32


@pytest.fixture(scope="module")
def state():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    free, _ = torch.cuda.mem_get_info()
    if free < 90e9:
        pytest.skip("needs ~80 GB of device memory")
    import entity_b200 as eb
    from entity_b200 import workloads
    sim = workloads.reconnection(N, ppc0=PPC, nfilter=0, fused=True, deposit_mode=eb.DEPOSIT_AGGREGATED,
                                 sort_interval=0, seed=0x5678)
    for sp in sim.species:
        sim.ctx.sort_particles(sp.arrays, sp.npart, remove_dead=0)
    return torch, eb, sim


def rho_nodes(torch, sim, sp, arrays=None):
    """charge density at the nodes of the active mesh (periodic), float64: sum of w q S(x)"""
    a = arrays or sp.arrays
    g = sim.grid
    n1, n2 = g.n[0], g.n[1]
    rho = torch.zeros(n1 * n2, dtype=torch.float64, device=sim.device)
    step = 1 << 25
    for lo in range(0, sp.npart, step):
        hi = min(sp.npart, lo + step)
        alive = a["tag"][lo:hi] == 1
        i, j = a["i1"][lo:hi].long(), a["i2"][lo:hi].long()
        dx, dy = a["dx1"][lo:hi].double(), a["dx2"][lo:hi].double()
        q = (a["weight"][lo:hi].double() * sp.charge) * alive
        for di, wx in ((0, 1.0 - dx), (1, dx)):
            for dj, wy in ((0, 1.0 - dy), (1, dy)):
                idx = ((j + dj) % n2) * n1 + ((i + di) % n1)
                rho.index_add_(0, idx, q * wx * wy)
    return rho.view(n2, n1)


def pusher_for(sim, sp):
    return sim.ctx.make_pusher(dt=sim.dt, omegaB0=sim.scales["omegaB0"], mass=sp.mass, charge=sp.charge,
                               dx=sim.ctx.dx, xmin=list(sim.ctx.xmin), pbc=list(sim.params.pbc))


def folded(torch, sim, cur):
    """the deposit's ghost contributions added to their periodic images; active part [3, n2, n1]"""
    buff = torch.zeros_like(cur)
    sim.ctx.sync_currents(cur, buff, list(sim.params.fbc))
    g = sim.grid
    return cur[:, g.ng:g.ng + g.n[1], g.ng:g.ng + g.n[0]].double()


def test_full_size_continuity_lean_and_unfused(state):
    torch, eb, sim = state
    sp = sim.species[0]
    n = sp.npart
    assert n > 1.6e8
    keep = {k: v.clone() for k, v in sp.arrays.items()}
    rho0 = rho_nodes(torch, sim, sp)
    gp = pusher_for(sim, sp)

    # fused kernel, reference-observable prev arrays
    cur = torch.zeros_like(sim.cur)
    sim.ctx.push_deposit(gp, sp.arrays, n, sim.em, cur, mode=eb.DEPOSIT_AGGREGATED)
    torch.cuda.synchronize()
    assert int((sp.arrays["tag"][:n] == 1).sum()) == n
    for k in ("ux1", "ux2", "ux3", "dx1", "dx2"):
        assert bool(torch.isfinite(sp.arrays[k][:n]).all()), k
    assert torch.equal(sp.arrays["dx1_prev"][:n], keep["dx1"][:n])
    rho1 = rho_nodes(torch, sim, sp)
    J = folded(torch, sim, cur.clone())
    div = (J[0] - torch.roll(J[0], 1, dims=1)) + (J[1] - torch.roll(J[1], 1, dims=0))
    drho = (rho1 - rho0) / sim.dt
    scale = float(drho.abs().max())
    err = float((div + drho).abs().max())
    assert scale > 0 and err <= 2e-4 * scale, f"continuity violated by {err:.3e} of {scale:.3e}"

    # the same launch with i_prev / dx_prev as scratch
    lean = {k: v.clone() for k, v in keep.items()}
    for k in ("i1_prev", "i2_prev", "dx1_prev", "dx2_prev"):
        lean[k].fill_(7)
    cur_l = torch.zeros_like(sim.cur)
    sim.ctx.set_lean_prev(True)
    sim.ctx.push_deposit(gp, lean, n, sim.em, cur_l, mode=eb.DEPOSIT_AGGREGATED)
    sim.ctx.set_lean_prev(False)
    torch.cuda.synchronize()
    for k in ("i1", "i2", "dx1", "dx2", "ux1", "ux2", "ux3", "tag", "weight"):
        assert torch.equal(lean[k][:n], sp.arrays[k][:n]), k
    m = n // 4 * 4
    assert float((lean["dx1_prev"][:m] - 7).abs().max()) == 0.0
    jmax = float(cur.abs().max())
    assert float((cur_l - cur).abs().max()) <= 2e-4 * jmax
    del lean, cur_l

    # the unfused pair from the same initial state
    un = {k: v.clone() for k, v in keep.items()}
    cur_u = torch.zeros_like(sim.cur)
    sim.ctx.push(gp, un, n, sim.em)
    sim.ctx.deposit(un, n, sp.charge, sim.dt, cur_u, mode=eb.DEPOSIT_AGGREGATED)
    torch.cuda.synchronize()
    same = (un["i1"][:n] == sp.arrays["i1"][:n]) & (un["i2"][:n] == sp.arrays["i2"][:n])
    assert float((~same).float().mean()) <= 1e-5
    for k in ("ux1", "ux2", "ux3", "dx1", "dx2"):
        d = (un[k][:n] - sp.arrays[k][:n]).abs()[same]
        assert float(d.max()) <= 1e-5 * max(1.0, float(sp.arrays[k][:n].abs().max())), k
    assert float((cur_u - cur).abs().max()) <= 2e-4 * jmax


def test_full_size_against_the_cpu_oracle(state):
    """One species of the full-size state (1.66e8 particles on 4096 x 2048 cells: 32-bit byte
    offsets of the packed gather and int J keys at their real magnitudes) through one push +
    deposit of the CPU oracle -- the reference's own kernels compiled in place when
    oracle/_ref travelled with the repo (all host threads), else the C++ restatement on the first
    2^25 particles -- against (a) the strict build's push kernel: every particle array bit for
    bit; (b) the fast build's fused kernel (the one the bench times): same cell for all but
    <= 1e-5 of the particles, momenta / offsets rtol 1e-4 atol 2e-5, J within 2e-4 of max|J|."""
    import os
    torch, eb, sim = state
    from oracle import orc
    from helpers import assert_prtls_values_equal

    def to_host(arrays, m):
        ps = orc.ParticleSet(m)
        for nm in ps.names():
            if nm in arrays:
                getattr(ps, nm)[:] = arrays[nm][:m].cpu().numpy()
        return ps
    sp = sim.species[1]
    impl = orc.reference(0)
    n = sp.npart
    if impl is None:
        impl, n = orc.oracle(), min(sp.npart, 1 << 25)
    else:
        impl.set_threads(os.cpu_count() or 1)
    keep = {k: v[:n].clone() for k, v in sp.arrays.items()}
    g = orc.Grid.make(N, 2)
    em = sim.em.cpu().numpy()
    p = to_host(keep, n)
    octx = orc.make_pusher(dt=sim.dt, omegaB0=sim.scales["omegaB0"], mass=sp.mass, charge=sp.charge,
                           dx=sim.ctx.dx, xmin=list(sim.ctx.xmin), pbc=list(sim.params.pbc))
    impl.push(g, 0, octx, p, n, em)
    j_ref = np.zeros(g.shape(3), np.float32)
    impl.deposit(g, 0, p, n, sp.charge, octx.dt, sim.ctx.dx, j_ref)
    impl.set_threads(1)
    jmax = float(np.abs(j_ref).max())
    assert jmax > 0

    # (a) strict build, unfused push: bit for bit
    sctx = eb.Context(N, order=0, strict=True, dx=sim.ctx.dx, xmin=tuple(sim.ctx.xmin))
    gp = sctx.make_pusher(dt=sim.dt, omegaB0=sim.scales["omegaB0"], mass=sp.mass, charge=sp.charge,
                          dx=sim.ctx.dx, xmin=list(sim.ctx.xmin), pbc=list(sim.params.pbc))
    a = {k: v.clone() for k, v in keep.items()}
    sctx.push(gp, a, n, sim.em)
    cur = torch.zeros_like(sim.cur)
    sctx.deposit(a, n, sp.charge, sim.dt, cur, mode=eb.DEPOSIT_AGGREGATED)
    torch.cuda.synchronize()
    assert_prtls_values_equal(to_host(a, n), p, what="strict push at full size")
    assert np.abs(cur.cpu().numpy() - j_ref).max() <= 2e-4 * jmax
    del a
    sctx.close()

    # (b) fast build, fused kernel
    b = {k: v.clone() for k, v in keep.items()}
    cur.zero_()
    sim.ctx.push_deposit(pusher_for(sim, sp), b, n, sim.em, cur, mode=eb.DEPOSIT_AGGREGATED)
    torch.cuda.synchronize()
    q = to_host(b, n)
    same = (q.i1 == p.i1) & (q.i2 == p.i2)
    assert (~same).mean() <= 1e-5, f"{(~same).sum()} particles in another cell"
    assert np.array_equal(q.tag, p.tag)
    for nm in ("ux1", "ux2", "ux3", "dx1", "dx2"):
        np.testing.assert_allclose(getattr(q, nm)[same], getattr(p, nm)[same], rtol=1e-4, atol=2e-5, err_msg=nm)
    assert np.abs(cur.cpu().numpy() - j_ref).max() <= 2e-4 * jmax


def test_full_size_steps_conserve_particles_and_stay_finite(state):
    torch, eb, sim = state
    n0 = [sp.npart for sp in sim.species]
    sim.step(3)
    assert [sp.npart for sp in sim.species] == n0
    for sp in sim.species:
        assert int((sp.arrays["tag"][:sp.npart] == 1).sum()) == sp.npart
    assert bool(torch.isfinite(sim.em).all()) and bool(torch.isfinite(sim.cur).all())
