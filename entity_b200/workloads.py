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
