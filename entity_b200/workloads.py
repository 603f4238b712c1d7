"""Synthetic plasma states of the shapes named in BASELINE.json (``configs``), generated with
torch on whatever device the simulation lives on (there is no network for real data). These
build INITIAL CONDITIONS only -- the reference leaves that to its problem generators
(pgens/*/pgen.hpp), which are outside the hot path; parity tests copy the generated state
into the oracle rather than regenerating it.

* ``two_stream``   -- pgens/streaming/twostream.toml lifted to 2D: two counter-streaming e-
                      beams (ux = +-0.1) on a static e+ background, T = 1e-4, zig-zag.
* ``reconnection`` -- pgens/reconnection (Harris sheet, sigma0 = 100, T_bg = 1e-4, hot
                      overdense sheet) in its periodic-core variant: two sheets at +-L_y/4 so
                      that the box is doubly periodic (no MATCH layer / replenish injector).
* ``turbulence``   -- pgens/turbulence extended to 3D: uniform magnetised pair plasma, T = 1,
                      B0 along z (the antenna is an external-current source term, out of path).
"""
from __future__ import annotations

import math

from . import lib as L
from .srpic import PRTL_DTYPES, Scales, Simulation


def _alloc(sim: Simulation, n, cap=None):
    torch = sim.torch
    cap = cap or n
    arrays = {}
    for k, dt in PRTL_DTYPES.items():
        axis = [c for c in k if c in "123"]
        if k.startswith(("i", "dx")) and axis and int(axis[0]) > sim.dim:
            continue
        arrays[k] = torch.zeros(cap, dtype=getattr(torch, dt), device=sim.device)
    return arrays


def _uniform_positions(sim, arrays, n, gen, lo=0):
    """ppc-uniform: particle p sits in cell (p // per_cell) so the state starts cell-sorted,
    like the reference's uniform injector fills cells in index order."""
    torch = sim.torch
    g = sim.grid
    ncell = 1
    for a in range(g.dim):
        ncell *= g.n[a]
    per_cell = n // ncell
    assert per_cell * ncell == n, "particle count must be a multiple of the cell count"
    cell = torch.arange(n, device=sim.device, dtype=torch.int64) // per_cell
    names = ["i1", "i2", "i3"]
    for a in range(g.dim):
        arrays[names[a]][lo:lo + n] = (cell % g.n[a]).to(torch.int32)
        cell = cell // g.n[a]
    for a in range(g.dim):
        d = torch.rand(n, device=sim.device, generator=gen, dtype=torch.float32)
        arrays[f"dx{a + 1}"][lo:lo + n] = d.clamp_(max=0.99999994)
        arrays[f"i{a + 1}_prev"][lo:lo + n] = arrays[names[a]][lo:lo + n]
        arrays[f"dx{a + 1}_prev"][lo:lo + n] = arrays[f"dx{a + 1}"][lo:lo + n]


def _maxwellian(sim, arrays, n, gen, temperature, drift=(0.0, 0.0, 0.0), lo=0):
    torch = sim.torch
    if temperature < 0.3:
        sig = math.sqrt(temperature)
        for a, nm in enumerate(("ux1", "ux2", "ux3")):
            u = torch.randn(n, device=sim.device, generator=gen, dtype=torch.float32) * sig
            arrays[nm][lo:lo + n] = u + drift[a]
    else:
        # relativistic limit of Maxwell-Juttner: |u| ~ T * Gamma(3, 1), isotropic direction
        r = torch.rand((3, n), device=sim.device, generator=gen, dtype=torch.float32).clamp_(min=1e-12)
        mag = -temperature * torch.log(r[0] * r[1] * r[2])
        mu = 2.0 * torch.rand(n, device=sim.device, generator=gen, dtype=torch.float32) - 1.0
        ph = 2.0 * math.pi * torch.rand(n, device=sim.device, generator=gen, dtype=torch.float32)
        st = torch.sqrt((1.0 - mu * mu).clamp_(min=0.0))
        arrays["ux1"][lo:lo + n] = mag * st * torch.cos(ph) + drift[0]
        arrays["ux2"][lo:lo + n] = mag * st * torch.sin(ph) + drift[1]
        arrays["ux3"][lo:lo + n] = mag * mu + drift[2]
    arrays["weight"][lo:lo + n] = 1.0
    arrays["tag"][lo:lo + n] = 1


def _gen(sim, seed):
    g = sim.torch.Generator(device=sim.device)
    g.manual_seed(seed)
    return g


def two_stream(n=(256, 256), ppc0=64, nfilter=4, seed=0x1234, **kw) -> Simulation:
    """configs[0]: 2D two-stream, 4 species of which 2 are pushed."""
    dim = len(n)
    dx = 100.0 / 1024.0  # physical cell size is irrelevant for throughput; keep O(0.1)
    sim = Simulation(n, 0, Scales(dim, dx, larmor0=100.0, skindepth0=10.0, ppc0=ppc0),
                     nfilter=nfilter, **kw)
    ncell = math.prod(n)
    per = ncell * (ppc0 // 4)
    gen = _gen(sim, seed)
    for k, (charge, pusher, drift) in enumerate([(-1.0, L.PUSHER_BORIS, 0.1),
                                                 (+1.0, L.PUSHER_NONE, 0.0),
                                                 (-1.0, L.PUSHER_BORIS, -0.1),
                                                 (+1.0, L.PUSHER_NONE, 0.0)]):
        arr = _alloc(sim, per)
        _uniform_positions(sim, arr, per, gen)
        _maxwellian(sim, arr, per, gen, 1e-4, (drift, 0.0, 0.0))
        sim.add_species(1.0, charge, arr, per, pusher)
    return sim


def reconnection(n=(4096, 2048), ppc0=32, nfilter=8, seed=0x5678, sheets=True,
                 capacity_factor=1.0, walls=False, **kw) -> Simulation:
    """configs[1]: 2D pair-plasma Harris sheet(s), periodic-core variant. The box is doubly
    periodic, so a multi-domain run tiles it: every block of `n` cells holds the same two
    sheets (own seed) and the global plasma is an array of Harris sheets. `capacity_factor`
    leaves room in the particle arrays for migration (maxnpart of the reference)."""
    torch_n1, torch_n2 = n
    Lx = 1000.0 * (n[0] / 4096.0)
    dx = Lx / n[0]
    Ly = dx * n[1]
    if walls:
        # the x2 boundaries of pgens/reconnection/reconnection.toml:16-21: fields MATCH (to the
        # initial profile, ds = 20), particles ABSORB; no replenishing injector
        kw.setdefault("fbc", [L.FBC_PERIODIC, L.FBC_PERIODIC, L.FBC_NONE, L.FBC_NONE, 0, 0])
        kw.setdefault("pbc", [L.PBC_PERIODIC, L.PBC_PERIODIC, L.PBC_ABSORB, L.PBC_ABSORB, 0, 0])
    sim = Simulation(n, 0, Scales(2, dx, larmor0=0.1, skindepth0=1.0, ppc0=ppc0), nfilter=nfilter,
                     xmin=(-0.5 * Lx, -0.5 * Ly, 0.0), **kw)
    torch = sim.torch
    g = sim.grid
    cs_width = 10.0 * (n[0] / 4096.0)
    y1, y2 = -0.25 * Ly, 0.25 * Ly
    # fields: B_x1 on (i, j+1/2) nodes; contravariant component = physical / dx
    jj = torch.arange(g.n[1] + 2 * g.ng, device=sim.device, dtype=torch.float32) - g.ng
    y = (jj + 0.5) * dx - 0.5 * Ly
    bx = torch.tanh((y - y1) / cs_width) - torch.tanh((y - y2) / cs_width) - 1.0
    sim.em[3] = (bx / dx)[:, None].expand(-1, g.n[0] + 2 * g.ng)
    if walls:
        import numpy as np
        target = torch.zeros_like(sim.em)
        target[3] = bx[:, None].expand(-1, g.n[0] + 2 * g.ng)  # tetrad components
        ds = 20.0 * (n[0] / 4096.0)
        nds = max(1, int(round(ds / dx)))
        ext = [g.n[0] + 2 * g.ng, g.n[1] + 2 * g.ng]
        ymin = float(np.float32(-0.5 * Ly))
        ymax = float(np.float32(ymin) + np.float32(dx) * np.float32(n[1]))
        sim.set_match([(1, ymin, ds, [0, 0], [ext[0], g.ng + nds]),
                       (1, ymax, ds, [0, g.ng + n[1] - nds], [ext[0], ext[1]])], target, 63)
        # the replenishing injector's boxes (pgen.hpp:236-241): 10 cells wide, inj_ypad = 50 (scaled
        # with the box) inside either x2 boundary, all of x1
        pad = max(1, int(round(50.0 * (n[0] / 4096.0) / dx)))
        sim.replenish_boxes = [([g.ng, g.ng + n[1] - pad - 10], [g.ng + n[0], g.ng + n[1] - pad]),
                               ([g.ng, g.ng + pad], [g.ng + n[0], g.ng + pad + 10])]
    gen = _gen(sim, seed)
    ncell = n[0] * n[1]
    n_bg = ncell * (ppc0 // 2)
    # hot sheet population: sech^2 profile, overdensity 3, T = sigma0 / (2 * overdensity)
    sigma0, over = 100.0, 3.0
    n_cs = 0
    if sheets:
        n_cs = int(over * 2.0 * (cs_width / dx) * n[0] * (ppc0 // 2)) * 2
    for charge in (-1.0, +1.0):
        arr = _alloc(sim, n_bg + n_cs, int((n_bg + n_cs) * capacity_factor))
        _uniform_positions(sim, arr, n_bg, gen)
        _maxwellian(sim, arr, n_bg, gen, 1e-4)
        if n_cs:
            half = n_cs // 2
            for s, yc in enumerate((y1, y2)):
                lo = n_bg + s * half
                # y ~ logistic around the sheet (density ~ sech^2), x uniform
                r = torch.rand(half, device=sim.device, generator=gen).clamp_(1e-6, 1 - 1e-6)
                yy = yc + 0.5 * cs_width * torch.log(r / (1.0 - r))
                yy = torch.remainder(yy + 0.5 * Ly, Ly)
                xx = torch.rand(half, device=sim.device, generator=gen) * Lx
                for nm_i, nm_d, pos, nn in (("i1", "dx1", xx, n[0]), ("i2", "dx2", yy, n[1])):
                    c = pos / dx
                    ci = torch.floor(c).clamp_(0, nn - 1)
                    arr[nm_i][lo:lo + half] = ci.to(torch.int32)
                    arr[nm_d][lo:lo + half] = (c - ci).clamp_(0.0, 0.99999994)
                    arr[nm_i + "_prev"][lo:lo + half] = arr[nm_i][lo:lo + half]
                    arr[nm_d + "_prev"][lo:lo + half] = arr[nm_d][lo:lo + half]
                _maxwellian(sim, arr, half, gen, 0.5 * sigma0 / over, lo=lo)
        sim.add_species(1.0, charge, arr, n_bg + n_cs)
    return sim


def turbulence(n=(128, 128, 128), ppc0=16, order=3, nfilter=4, seed=0x9abc, capacity_factor=1.0,
               **kw) -> Simulation:
    """configs[2] (per-GPU block): 3D pair plasma, T = 1, guide field along x3, 3rd-order shapes."""
    dx = 256.0 / 1024.0
    sim = Simulation(n, order, Scales(3, dx, larmor0=1.0, skindepth0=1.0, ppc0=ppc0),
                     nfilter=nfilter, **kw)
    # B_x3 = 1; in 3D every component is in-plane: contravariant = physical / dx
    sim.em[5] = 1.0 / dx
    gen = _gen(sim, seed)
    per = math.prod(n) * (ppc0 // 2)
    for charge in (-1.0, +1.0):
        arr = _alloc(sim, per, int(per * capacity_factor))
        _uniform_positions(sim, arr, per, gen)
        _maxwellian(sim, arr, per, gen, 1.0)
        sim.add_species(1.0, charge, arr, per)
    return sim


# --------------------------------------------------------------- curvilinear / GR shapes
def _node_metric(metric, n, mp, ng):
    """metric quantities (lib.metric_eval) on the four staggerings of the ghost-inclusive 2D
    mesh: dict key (s1, s2) in {0, 1}^2 -> array [N2, N1, nq]; host setup code, like the
    reference's pgens evaluate the metric when they fill the initial fields"""
    import numpy as np
    N1, N2 = n[0] + 2 * ng, n[1] + 2 * ng
    i = np.arange(N1, dtype=np.float32) - ng
    j = np.arange(N2, dtype=np.float32) - ng
    out = {}
    for s1 in (0, 1):
        for s2 in (0, 1):
            x1, x2 = np.meshgrid(i + 0.5 * s1, j + 0.5 * s2)
            # keep the query inside the coordinate range of the metric (ghost rows beyond the axis
            # mirror the first interior row)
            x2c = np.clip(x2, 0.0, float(n[1]))
            q = L.metric_eval(metric, n, mp, x1.ravel(), x2c.ravel())
            out[(s1, s2)] = q.reshape(N2, N1, -1)
    return out


def _cell_particles(sim_like, torch, device, cap, i_lo, i_hi, n2, per_cell, gen, temperature, phi=True):
    """per_cell particles in every cell of [i_lo, i_hi) x [0, n2), cell-sorted (i1 fastest)"""
    ncell = (i_hi - i_lo) * n2
    n = ncell * per_cell
    arrays = {}
    for k, dt in PRTL_DTYPES.items():
        axis = [c for c in k if c in "123"]
        if k.startswith(("i", "dx")) and axis and int(axis[0]) > 2:
            continue
        arrays[k] = torch.zeros(cap, dtype=getattr(torch, dt), device=device)
    if phi:
        arrays["phi"] = torch.zeros(cap, dtype=torch.float32, device=device)
    cell = torch.arange(n, device=device, dtype=torch.int64) // per_cell
    arrays["i1"][:n] = (i_lo + cell % (i_hi - i_lo)).to(torch.int32)
    arrays["i2"][:n] = (cell // (i_hi - i_lo)).to(torch.int32)
    for a in ("dx1", "dx2"):
        arrays[a][:n] = torch.rand(n, device=device, generator=gen).clamp_(max=0.99999994)
    for a in ("ux1", "ux2", "ux3"):
        arrays[a][:n] = torch.randn(n, device=device, generator=gen) * math.sqrt(temperature)
    arrays["weight"][:n] = 1.0
    arrays["tag"][:n] = 1
    for a in ("i1", "i2", "dx1", "dx2"):
        arrays[a + "_prev"][:n] = arrays[a][:n]
    return arrays, n


def wald(n=(512, 512), ppc=8, niter=10, seed=0x77, metric=L.METRIC_QKERR_SCHILD, spin=0.95,
         extent=(1.0, 10.0), nfilter=4, device=0):
    """configs[4]: pgens/wald (2D GR, qkerr_schild a = 0.95, 512 x 512) with two Boris species
    loaded uniformly in r in [2, 8] (the reference's wald.toml is a vacuum setup; species as in
    pgens/accretion/accretion.toml:43-55), pusher_niter = 10. Synthetic fields: the
    flat-space limit of a uniform vertical field in the orthonormal frame, no E."""
    import numpy as np
    import torch
    from .grpic import GRSimulation
    mp = [extent[0], extent[1], 0.0, float(np.float32(np.pi)), 0.0, 0.0, spin]
    # scales of wald.toml: larmor0 = 0.0025, skindepth0 = 0.05; dt = CFL dx0 with dx0 ~ the
    # smallest proper cell size (near the horizon)
    ng = 2
    q = _node_metric(metric, n, mp, ng)
    h11, h22 = q[(0, 0)][ng:-ng, ng:-ng, 0], q[(0, 0)][ng:-ng, ng:-ng, 1]
    dx0 = float(min(np.sqrt(h11).min(), np.sqrt(h22).min()))
    dt = 0.5 * dx0 / math.sqrt(2.0)
    larmor0, skin0 = 0.0025, 0.05
    V0 = float(q[(1, 1)][ng + n[1] // 2, ng + n[0] // 2, 10])  # sqrt_det_h at the mesh centre
    sim = GRSimulation(n, metric, mp, dt=dt, omegaB0=1.0 / larmor0, q0=V0 / (2 * ppc * skin0 ** 2),
                       B0=1.0 / larmor0, nfilter=nfilter, pusher_niter=niter, pusher_eps=1e-2,
                       deposit_mode=L.DEPOSIT_ATOMIC, device=device)
    # B^r = cos(theta) / sqrt(h_11) on (i, j + 1/2); B^theta = -sin(theta) / sqrt(h_22) on (i + 1/2, j)
    th01, th10 = q[(0, 1)][:, :, 25], q[(1, 0)][:, :, 25]
    b1 = np.cos(th01) / np.sqrt(q[(0, 1)][:, :, 0])
    b2 = -np.sin(th10) / np.sqrt(q[(1, 0)][:, :, 1])
    for f in (sim.em, sim.em0):
        f[3] = torch.from_numpy(np.nan_to_num(b1).astype(np.float32)).to(sim.device)
        f[4] = torch.from_numpy(np.nan_to_num(b2).astype(np.float32)).to(sim.device)
    # MATCH layer at the outer edge towards the initial field, 1 r_g thick
    tgt = sim.em.clone()
    chi_min = math.log(extent[0])
    dchi = (math.log(extent[1]) - chi_min) / n[0]
    lo = int(math.floor((math.log(extent[1] - 1.0) - chi_min) / dchi))
    sim.set_match(tgt, 0b111000, extent[1], 1.0, [lo + ng, 0], [n[0] + 2 * ng, n[1] + 2 * ng])
    gen = torch.Generator(device=sim.device)
    gen.manual_seed(seed)
    i_lo = int(math.ceil((math.log(2.0) - chi_min) / dchi))
    i_hi = int(math.floor((math.log(8.0) - chi_min) / dchi))
    for charge in (-1.0, 1.0):
        cap = (i_hi - i_lo) * n[1] * ppc
        arrays, npart = _cell_particles(sim, torch, sim.device, cap, i_lo, i_hi, n[1], ppc, gen, 0.01)
        sp = sim.alloc_species(1.0, charge, cap)
        sp.arrays, sp.npart = arrays, npart
    sim._species_c = None
    return sim


def magnetosphere(n=(2048, 1024), ppc=10, seed=0x88, extent=(1.0, 50.0), nfilter=4, device=0, inject=True):
    """configs[3]: pgens/magnetosphere/magnetosphere.toml (2D qspherical SR, 2048 x 1024): a
    dipole, ATMOSPHERE / MATCH / AXIS field boundaries, Boris + GCA pusher with the atmosphere's
    gravity, weighted particles. The synthetic state carries `ppc` pairs per cell throughout the
    domain (a filled magnetosphere instead of the start-up transient); inject = True registers the
    atmosphere injector of the toml (temperature 0.1, density 10, height 0.02, ds 2) with the step,
    so that srpic::ParticleInjector runs every step as in the reference."""
    import numpy as np
    import torch
    metric = L.METRIC_QSPHERICAL
    mp = [extent[0], extent[1], 0.0, float(np.float32(np.pi)), 0.0, 0.0, 0.0]
    ng = 2
    q = _node_metric(metric, n, mp, ng)
    sh11, sh22 = q[(0, 0)][ng:-ng, ng:-ng, 3], q[(0, 0)][ng:-ng, ng:-ng, 4]
    dx0 = float(min(sh11.min(), sh22.min()))
    dt = 0.5 * dx0 / math.sqrt(2.0)
    larmor0, skin0 = 2e-5, 0.01
    V0 = float(q[(1, 1)][ng + n[1] // 2, ng + n[0] // 2, 6])
    scales = dict(dt=dt, omegaB0=1.0 / larmor0, q0=V0 / (ppc * skin0 ** 2), B0=1.0 / larmor0, V0=V0,
                  n0=ppc / V0, ppc0=float(ppc), correction=1.0)
    fbc = [L.FBC_ATMOSPHERE, L.FBC_MATCH, L.FBC_AXIS, L.FBC_AXIS, 0, 0]
    pbc = [L.PBC_ABSORB, L.PBC_ABSORB, L.PBC_AXIS, L.PBC_AXIS, 0, 0]
    sim = Simulation(n, 0, scales, nfilter=nfilter, fused=False, deposit_mode=L.DEPOSIT_ATOMIC,
                     fbc=fbc, pbc=pbc, metric=metric, metric_params=mp, device=device)
    # dipole (pgen.hpp:40-57): B^r = cos(theta) / r^3, B^theta = sin(theta) / (2 r^3) in the
    # orthonormal frame; contravariant = / sqrt(h_ii)
    r01, th01 = q[(0, 1)][:, :, 8], q[(0, 1)][:, :, 9]
    r10, th10 = q[(1, 0)][:, :, 8], q[(1, 0)][:, :, 9]
    b1 = np.cos(th01) / r01 ** 3 / q[(0, 1)][:, :, 3]
    b2 = 0.5 * np.sin(th10) / r10 ** 3 / q[(1, 0)][:, :, 4]
    sim.em[3] = torch.from_numpy(np.nan_to_num(b1).astype(np.float32)).to(sim.device)
    sim.em[4] = torch.from_numpy(np.nan_to_num(b2).astype(np.float32)).to(sim.device)
    tgt = sim.em.clone()
    chi_min = math.log(extent[0])
    dchi = (math.log(extent[1]) - chi_min) / n[0]
    lo = int(math.floor((math.log(extent[1] - 1.0) - chi_min) / dchi))
    buf = max(nfilter + 2, 5)
    sim.set_field_bcs([
        dict(kind=L.FBC_ATMOSPHERE, o=0, sign=-1, target=tgt, mask=63, range_min=[0, 0],
             range_max=[buf + ng, n[1] + 2 * ng], i_edge=buf + ng - 1),
        dict(kind=L.FBC_MATCH, o=0, sign=+1, target=tgt, mask=0b011000, range_min=[lo + ng, 0],
             range_max=[n[0] + 2 * ng, n[1] + 2 * ng], xg_edge=extent[1], ds=1.0),
    ])
    sim.set_gca(1.0, 0.9)
    sim.set_atmosphere((-5.0, 0.0, 0.0), extent[0] * math.exp(buf * dchi), 2.0)
    gen = torch.Generator(device=sim.device)
    gen.manual_seed(seed)
    per = max(1, ppc // 2)
    for charge in (-1.0, 1.0):
        npart0 = n[0] * n[1] * per
        cap = int(npart0 * (1.25 if inject else 1.0))
        arrays, npart = _cell_particles(sim, torch, sim.device, cap, 0, n[0], n[1], per, gen, 0.1)
        sim.add_species(1.0, charge, arrays, npart, L.PUSHER_BORIS | L.PUSHER_GCA, cap)
    if inject:
        sim.set_atmosphere_injector(dict(dim=0, sign=-1, x_surf=extent[0] * math.exp(buf * dchi), ds=2.0,
                                         height=0.02, temperature=0.1, density=10.0, species=(0, 1)))
    return sim
