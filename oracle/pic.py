"""TEST INFRASTRUCTURE ONLY -- the SRPIC step in the reference's order on the CPU oracle.

Mirrors ``SRPICEngine::step_forward`` (src/engines/srpic/srpic.hpp:65-188) for one periodic
Minkowski domain, with the coefficient formulas of src/engines/srpic/fieldsolvers.h in fp32,
calling the oracle kernels (``orc.Impl``: the C++ port, or the compiled reference) on numpy
arrays. It is the checker for ``eb200_srpic_step`` and the CPU baseline of bench.py."""
from __future__ import annotations

import numpy as np

from . import orc

f32 = np.float32


class OracleSim:
    def __init__(self, impl: orc.Impl, n, order, scales: dict, dx, nfilter, fbc=None, pbc=None,
                 xmin=(0.0, 0.0, 0.0)):
        self.impl = impl
        self.order = order
        self.grid = orc.Grid.make(n, orc.nghosts_for(order))
        self.dim = len(n)
        self.s = scales
        self.dx = f32(dx)
        self.xmin = xmin
        self.nfilter = nfilter
        self.fbc = fbc or [orc.FBC_PERIODIC] * 6
        self.pbc = pbc or [orc.PBC_PERIODIC] * 6
        self.em = np.zeros(self.grid.shape(6), f32)
        self.cur = np.zeros(self.grid.shape(3), f32)
        self.buff = np.zeros(self.grid.shape(3), f32)
        self.species = []  # dicts: mass, charge, pusher, prtls(ParticleSet), npart
        self.step_index = 0
        self.time = 0.0
        self.match = None  # (faces, target, mask): MATCH field boundaries (oracle/bcs.py)
        self.ext = None  # mode table of the pgen's ext_current (oracle/antenna.py)

    # srpic::FieldBoundaries for MATCH faces (src/engines/srpic/fields_bcs.h:38-215, 600-672)
    def _field_boundaries(self, tags):
        if not self.match:
            return
        from . import bcs
        faces, target, mask = self.match
        for o, xg_edge, ds, rmin, rmax in faces:
            bcs.match_fields(self.grid, self.em, target, o, self.dx, self.xmin[o], xg_edge, ds,
                             tags, mask, rmin, rmax)

    def add_species(self, mass, charge, prtls: orc.ParticleSet, npart, pusher=orc.PUSHER_BORIS):
        self.species.append(dict(mass=mass, charge=charge, pusher=pusher, prtls=prtls, npart=npart))

    # fieldsolvers.h:36-139
    def _coeffs(self, fraction):
        dT = f32(fraction) * f32(self.s["correction"]) * f32(self.s["dt"])
        dx = np.sqrt(self.dx * self.dx)
        if self.dim == 2:
            return dT / (dx * dx), dT
        return dT / dx, f32(0.0)

    def step(self):
        g, im, s = self.grid, self.impl, self.s
        dt = f32(s["dt"])
        if self.step_index == 0:
            im.comm_fields(g, self.em, 0, 6, self.fbc)
            self._field_boundaries(3)
        c1, c2 = self._coeffs(0.5)
        im.faraday(g, self.em, c1, c2, None)
        im.comm_fields(g, self.em, 3, 6, self.fbc)
        self._field_boundaries(2)  # BC::B (srpic.hpp:93-101)
        for sp in self.species:
            if sp["pusher"] == orc.PUSHER_NONE or sp["npart"] == 0:
                continue
            ctx = orc.make_pusher(pusher_flags=sp["pusher"], mass=sp["mass"], charge=sp["charge"],
                                  time=self.time, dt=dt, omegaB0=s["omegaB0"], pbc=self.pbc,
                                  dx=self.dx, xmin=list(self.xmin))
            im.push(g, self.order, ctx, sp["prtls"], sp["npart"], self.em)
        self.cur[...] = 0
        for sp in self.species:
            if (sp["pusher"] == orc.PUSHER_NONE or sp["npart"] == 0
                    or abs(sp["charge"]) <= np.finfo(f32).eps):
                continue
            im.deposit(g, self.order, sp["prtls"], sp["npart"], sp["charge"], dt, self.dx, self.cur)
        im.sync_currents(g, self.cur, self.buff, self.fbc)
        im.comm_fields(g, self.cur, 0, 3, self.fbc)
        for _ in range(self.nfilter):
            self.buff[...] = self.cur
            im.filter_pass(g, self.cur, self.buff, self.fbc)
            im.comm_fields(g, self.cur, 0, 3, self.fbc)
        im.faraday(g, self.em, c1, c2, None)
        im.comm_fields(g, self.em, 3, 6, self.fbc)
        self._field_boundaries(2)  # BC::B (srpic.hpp:144-152)
        c1, c2 = self._coeffs(1.0)
        im.ampere(g, self.em, c1, c2)
        coeff = -dt * f32(s["q0"]) / (f32(s["B0"]) * f32(s["V0"]))
        if self.ext is not None:
            # CurrentsAmpere_kernel<D, ExtCurrent>: J += ppc0 * ext_current first (ampere_mink.hpp:134-215)
            from . import antenna
            antenna.add_ext_current(g, self.cur, self.ext, s["ppc0"], self.dx, self.xmin)
        im.currents_ampere(g, self.em, self.cur, coeff, f32(s["ppc0"]))
        im.comm_fields(g, self.em, 0, 3, self.fbc)
        im.comm_fields(g, self.cur, 0, 3, self.fbc)
        self._field_boundaries(1)  # BC::E (srpic.hpp:168-176)
        self.step_index += 1
        self.time += float(dt)

    def n_pushed(self):
        return sum(sp["npart"] for sp in self.species if sp["pusher"] != orc.PUSHER_NONE)


def from_device_sim(sim, impl=None) -> OracleSim:
    """Copy the full state of an entity_b200 Simulation into an oracle simulation."""
    impl = impl or orc.oracle()
    g = sim.grid
    n = [g.n[a] for a in range(g.dim)]
    o = OracleSim(impl, n, sim.order, sim.scales, sim.ctx.dx, sim.params.nfilter,
                  fbc=list(sim.params.fbc), pbc=list(sim.params.pbc), xmin=sim.ctx.xmin)
    o.em[...] = sim.em.cpu().numpy()
    o.cur[...] = sim.cur.cpu().numpy()
    o.step_index, o.time = sim.step_index, sim.time
    if getattr(sim, "_match", None) and sim._match[2]:
        o.match = (sim._match[2], sim._match[1].cpu().numpy(), sim._match[3])
    for sp in sim.species:
        cap = sp.maxnpart
        ps = orc.ParticleSet(cap)
        for nm in ps.names():
            if nm in sp.arrays:
                getattr(ps, nm)[:] = sp.arrays[nm].cpu().numpy()
        o.add_species(sp.mass, sp.charge, ps, sp.npart, sp.pusher)
    return o
