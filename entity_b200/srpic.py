"""Python driver of the SRPIC step (``eb200_srpic_step``): owns the device tensors of one
Minkowski domain and calls the C++ engine mirror in ``csrc/engine.cu``. Plumbing only --
all arithmetic on fields and particles happens in the CUDA library."""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

from . import lib as L
from .lib import ParamsC, SpeciesC

PRTL_DTYPES = {
    "i1": "int32", "i2": "int32", "i3": "int32",
    "dx1": "float32", "dx2": "float32", "dx3": "float32",
    "ux1": "float32", "ux2": "float32", "ux3": "float32", "weight": "float32",
    "i1_prev": "int32", "i2_prev": "int32", "i3_prev": "int32",
    "dx1_prev": "float32", "dx2_prev": "float32", "dx3_prev": "float32",
    "tag": "int16",
}


@dataclass
class Scales:
    """Derived scales exactly as SimulationParams computes them (parameters.cpp:47-78,
    grid.cpp:39-50, algorithms.cpp:21-22), in fp32."""
    dim: int
    dx: float
    larmor0: float
    skindepth0: float
    ppc0: float
    cfl: float = 0.5
    correction: float = 1.0

    def derive(self):
        import numpy as np
        f = np.float32
        dx = f(self.dx)
        dx0 = dx / np.sqrt(f(self.dim))
        V0 = dx if self.dim == 1 else (dx * dx if self.dim == 2 else dx * dx * dx)
        out = dict(
            dt=f(self.cfl) * dx0,
            omegaB0=f(1.0) / f(self.larmor0),
            B0=f(1.0) / f(self.larmor0),
            V0=f(V0),
            q0=f(V0) / (f(self.ppc0) * (f(self.skindepth0) * f(self.skindepth0))),
            n0=f(self.ppc0) / f(V0),
            ppc0=f(self.ppc0),
            correction=f(self.correction),
        )
        return {k: float(v) for k, v in out.items()}


@dataclass
class Species:
    mass: float
    charge: float
    pusher: int = L.PUSHER_BORIS
    drag: int = L.DRAG_NONE
    npart: int = 0
    maxnpart: int = 0
    arrays: dict = field(default_factory=dict)


class Simulation:
    """One Minkowski domain: em/cur/buff fields + species on one GPU."""

    def __init__(self, n, order, scales, nfilter=0, strict=False, fused=False,
                 deposit_mode=L.DEPOSIT_ATOMIC, fbc=None, pbc=None, sort_interval=0,
                 clear_interval=0, device=0, xmin=(0.0, 0.0, 0.0), stencil=None,
                 fieldsolver=True, deposit=True, metric=L.METRIC_MINKOWSKI, metric_params=None):
        """`scales`: a Scales (Minkowski: derived as the reference derives them) or a dict with
        dt, omegaB0, q0, B0, V0, n0, ppc0, correction (curvilinear: what the host passes)."""
        import torch
        self.torch = torch
        self.device = torch.device("cuda", device)
        self.dim = len(n)
        self.metric = metric
        if metric == L.METRIC_MINKOWSKI:
            self.ctx = L.Context(n, order=order, strict=strict, device=device, dx=scales.dx, xmin=xmin)
        else:
            self.ctx = L.Context(n, order=order, strict=strict, device=device, metric=metric,
                                 metric_params=metric_params)
        self.grid = self.ctx.grid
        self.scales = scales.derive() if isinstance(scales, Scales) else dict(scales)
        self.order = order
        shape6, shape3 = self.grid.shape(6), self.grid.shape(3)
        self.em = torch.zeros(shape6, dtype=torch.float32, device=self.device)
        self.cur = torch.zeros(shape3, dtype=torch.float32, device=self.device)
        self.buff = torch.zeros(shape3, dtype=torch.float32, device=self.device)
        self.species: list[Species] = []
        self.step_index = 0
        self.time = 0.0
        p = ParamsC()
        s = self.scales
        p.dt, p.correction, p.omegaB0 = s["dt"], s["correction"], s["omegaB0"]
        p.q0, p.B0, p.V0, p.ppc0 = s["q0"], s["B0"], s["V0"], s["ppc0"]
        p.nfilter = nfilter
        p.fieldsolver_enabled, p.deposit_enabled = int(fieldsolver), int(deposit)
        p.stencil = (C.c_float * 9)(*(stencil or [0.0] * 9))
        p.fbc = (C.c_int * 6)(*(fbc or [L.FBC_PERIODIC] * 6))
        p.pbc = (C.c_int * 6)(*(pbc or [L.PBC_PERIODIC] * 6))
        p.fuse_push_deposit = int(fused)
        p.deposit_mode = deposit_mode
        p.sort_interval, p.clear_interval = sort_interval, clear_interval
        p.n0 = s.get("n0", 0.0)
        self.params = p
        self._species_c = None
        self._field_bcs = None

    @property
    def dt(self):
        return self.scales["dt"]

    def set_match(self, faces, target, mask=63):
        """MATCH field boundaries inside the step (srpic::FieldBoundaries): faces = list of
        (o, xg_edge, ds, range_min, range_max); target = device tensor in the layout of em with
        the MatchFields values on every component's node. The tensors are kept alive here."""
        arr = (L.MatchFaceC * max(1, len(faces)))()
        d = self.dim
        for k, (o, xg_edge, ds, rmin, rmax) in enumerate(faces):
            arr[k].o, arr[k].xg_edge, arr[k].ds = o, xg_edge, ds
            arr[k].range_min = (C.c_int * 3)(*(list(rmin) + [0] * (3 - d)))
            arr[k].range_max = (C.c_int * 3)(*(list(rmax) + [1] * (3 - d)))
        self._match = (arr, target, list(faces), mask)
        self.ctx._check(self.ctx.lib.eb200_srpic_set_match(
            self.ctx.handle, arr, len(faces), target.data_ptr() if len(faces) else None, mask))

    def set_gca(self, larmor_max, e_ovr_b_max):
        self.params.gca_larmor_max, self.params.gca_e_ovr_b_max = larmor_max, e_ovr_b_max

    def set_atmosphere(self, g, x_surf, ds):
        """srpic::ParticlePush's atmosphere context (particle_pusher.h:45-80): g = (gx1, gx2, gx3)
        with the direction's sign, the surface coordinate and grid.boundaries.atmosphere.ds"""
        p = self.params
        p.has_atmosphere = 1
        p.atm_g = (C.c_float * 3)(*g)
        p.atm_x_surf, p.atm_ds = x_surf, ds

    def set_field_bcs(self, bcs):
        """MATCH / ATMOSPHERE faces of a curvilinear domain: list of dicts with kind, o, sign,
        target (device tensor, layout of em), mask, range_min, range_max and xg_edge, ds (MATCH) or
        i_edge (ATMOSPHERE). Tensors are kept alive here."""
        arr = (L.FieldBCC * max(1, len(bcs)))()
        for k, b in enumerate(bcs):
            arr[k].kind, arr[k].o, arr[k].sign = b["kind"], b["o"], b["sign"]
            arr[k].xg_edge, arr[k].ds = b.get("xg_edge", 0.0), b.get("ds", 1.0)
            arr[k].i_edge = b.get("i_edge", 0)
            arr[k].range_min = (C.c_int * 2)(*b["range_min"])
            arr[k].range_max = (C.c_int * 2)(*b["range_max"])
            arr[k].target = b["target"].data_ptr()
            arr[k].mask = b["mask"]
        self._field_bcs = (arr, [b["target"] for b in bcs])
        self.ctx._check(self.ctx.lib.eb200_srpic_set_field_bcs(self.ctx.handle, arr, len(bcs)))

    # ------------------------------------------------------------ injection / moments
    def particle_moment(self, what, species_indices, buff=None, comp=0, use_weights=False):
        """arch::ComputeMomentWithSpecies: zeroes `buff` (default: self.buff) and adds the moment
        of the listed species (0-based) into component comp"""
        buff = self.buff if buff is None else buff
        buff.zero_()
        inv_n0 = 1.0 / self.scales["n0"]
        for k in species_indices:
            sp = self.species[k]
            s = L.Context.prtls_struct(sp.arrays)
            self.ctx._check(self.ctx.lib.eb200_particle_moment(
                self.ctx.handle, C.byref(s), sp.npart, C.c_float(sp.mass), C.c_float(sp.charge),
                int(use_weights), what, C.c_float(inv_n0), C.c_void_p(buff.data_ptr()),
                int(buff.shape[0]), comp, L.Context._stream(None)))
        return buff

    def inject_nonuniform(self, pair, ppc, sdist_kind=L.SDIST_UNIFORM, field=None, comp=0,
                          target=1.0, temperatures=(0.0, 0.0), drifts=((0, 0, 0), (0, 0, 0)),
                          range_min=None, range_max=None, seed=0x123456789abcdef0, call=0,
                          target_field=None, target_max=0.0, atmosphere=None):
        """arch::InjectNonUniform for the species pair (0-based indices); returns the number of
        pairs injected. ppc = number_density * ppc0 / 2. target_field / target_max: the generic
        Replenish distribution; atmosphere = dict(dim, sign, nmax, height, xsurf, ds): Replenish
        with the AtmosphereDensityProfile."""
        if self._species_c is None:
            self._species_c = self._pack_species()
        arr = self._species_c
        for k, sp in enumerate(self.species):
            arr[k].npart = sp.npart
        g = self.grid
        rmin = range_min or [g.ng] * self.dim
        rmax = range_max or [g.ng + g.n[a] for a in range(self.dim)]
        sd = L.SpatialDistC()
        sd.kind, sd.comp, sd.target_density = sdist_kind, comp, target
        sd.field = field.data_ptr() if field is not None else None
        sd.target_field = target_field.data_ptr() if target_field is not None else None
        sd.target_max = target_max
        sd.inv_V0 = 1.0 / self.scales["V0"]
        if atmosphere is not None:
            sd.atm_dim, sd.atm_sign = atmosphere["dim"], atmosphere["sign"]
            sd.atm_nmax, sd.atm_height = atmosphere["nmax"], atmosphere["height"]
            sd.atm_xsurf, sd.atm_ds = atmosphere["xsurf"], atmosphere["ds"]
        eds = []
        for t, d in zip(temperatures, drifts):
            e = L.MaxwellianC()
            e.temperature = t
            e.drift_u = (C.c_float * 3)(*d)
            eds.append(e)
        n0 = int(arr[pair[0]].npart)
        self.ctx._check(self.ctx.lib.eb200_inject_nonuniform(
            self.ctx.handle, C.byref(arr[pair[0]]), C.byref(arr[pair[1]]), C.c_float(ppc),
            C.byref(sd), C.byref(eds[0]), C.byref(eds[1]), (C.c_int * 3)(*(list(rmin) + [0] * 3)[:3]),
            (C.c_int * 3)(*(list(rmax) + [1] * 3)[:3]), C.c_uint64(seed), self.step_index, call,
            L.Context._stream(None)))
        for k in pair:
            self.species[k].npart = int(arr[k].npart)
        return int(arr[pair[0]].npart) - n0

    def _atmosphere_c(self, atm, seed=0x123456789abcdef0):
        """dict(dim, sign, x_surf, ds, height, temperature, density, species) -> eb200_atmosphere_t"""
        a = L.AtmosphereC()
        a.dim, a.sign = atm["dim"], atm["sign"]
        a.x_surf, a.ds, a.height = atm["x_surf"], atm["ds"], atm["height"]
        a.temperature, a.density = atm["temperature"], atm["density"]
        a.species = (C.c_int * 2)(*atm["species"])
        a.inv_n0, a.inv_V0 = 1.0 / self.scales["n0"], 1.0 / self.scales["V0"]
        a.ppc0, a.seed = self.scales["ppc0"], atm.get("seed", seed)
        return a

    def atmosphere_particles(self, atm, plane=None, assume_empty=False):
        """srpic::AtmosphereParticlesIn (eb200_atmosphere_particles); returns pairs injected"""
        if self._species_c is None:
            self._species_c = self._pack_species()
        arr = self._species_c
        for k, sp in enumerate(self.species):
            arr[k].npart = sp.npart
        plane = self.buff if plane is None else plane
        a = self._atmosphere_c(atm)
        k0 = atm["species"][0]
        n0 = int(arr[k0].npart)
        self.ctx._check(self.ctx.lib.eb200_atmosphere_particles(
            self.ctx.handle, C.byref(a), arr, len(self.species), C.c_void_p(plane.data_ptr()),
            int(assume_empty), self.step_index, L.Context._stream(None)))
        for k, sp in enumerate(self.species):
            sp.npart = int(arr[k].npart)
        return int(arr[k0].npart) - n0

    def set_emission(self, species, photon_species, kind, photon_weight, photon_energy_min,
                     nominal_probability, nominal_photon_energy, should_drag=False,
                     seed=0x123456789abcdef0):
        """Emission policy of `species` (0-based) inside the step; kind None clears it"""
        if kind is None:
            self.ctx._check(self.ctx.lib.eb200_srpic_set_emission(self.ctx.handle, species, -1, None))
            return
        e = L.EmissionC()
        e.kind, e.photon_weight, e.photon_energy_min = kind, photon_weight, photon_energy_min
        e.nominal_probability, e.nominal_photon_energy = nominal_probability, nominal_photon_energy
        e.should_drag, e.seed = int(should_drag), seed
        self.ctx._check(self.ctx.lib.eb200_srpic_set_emission(self.ctx.handle, species, photon_species,
                                                              C.byref(e)))

    def set_atmosphere_injector(self, atm):
        """Registers the atmosphere injector of the step (None clears it)"""
        self._atm_c = self._atmosphere_c(atm) if atm is not None else None
        self.ctx._check(self.ctx.lib.eb200_srpic_set_atmosphere_injector(
            self.ctx.handle, C.byref(self._atm_c) if atm is not None else None))

    def replenish(self, boxes, pair=(0, 1), temperature=1e-4, target=1.0):
        """pgens/reconnection/pgen.hpp:222-278 (CustomPostStep): the mass density of the pair into
        buff[0], then arch::InjectNonUniform with ReplenishUniform(target) and a Maxwellian of the
        background temperature in each box (ghost-inclusive cell ranges); returns pairs injected"""
        self.particle_moment(L.STATS_RHO, list(pair), comp=0)
        n = 0
        for k, (rmin, rmax) in enumerate(boxes):
            n += self.inject_nonuniform(pair, 0.5 * self.scales["ppc0"], L.SDIST_REPLENISH, self.buff,
                                        comp=0, target=target, temperatures=(temperature, temperature),
                                        range_min=rmin, range_max=rmax, call=k)
        return n

    def set_ext_current(self, table):
        """The pgen's ext_current as a table of Fourier modes (eb200_ext_current_t; see
        lib.ExtCurrentC.from_table); None clears it. The host refills it whenever the pgen
        advances the amplitudes (pgens/turbulence/pgen.hpp CustomPostStep)."""
        x = L.ExtCurrentC.from_table(table) if table is not None else None
        self.ctx._check(self.ctx.lib.eb200_srpic_set_ext_current(
            self.ctx.handle, C.byref(x) if x is not None else None))

    def add_species(self, mass, charge, arrays: dict, npart: int, pusher=L.PUSHER_BORIS,
                    maxnpart=None):
        """arrays: name -> torch tensor on this device (capacity = maxnpart)."""
        cap = maxnpart or next(iter(arrays.values())).numel()
        sp = Species(mass, charge, pusher, L.DRAG_NONE, npart, cap, arrays)
        self.species.append(sp)
        self._species_c = None
        return sp

    def alloc_species(self, mass, charge, maxnpart, pusher=L.PUSHER_BORIS):
        torch = self.torch
        arrays = {}
        for k in PRTL_DTYPES:
            axis = [c for c in k if c in "123"]
            if k.startswith(("i", "dx")) and axis and int(axis[0]) > self.dim:
                continue
            arrays[k] = torch.zeros(maxnpart, dtype=getattr(torch, PRTL_DTYPES[k]),
                                    device=self.device)
        if self.metric != L.METRIC_MINKOWSKI:
            arrays["phi"] = torch.zeros(maxnpart, dtype=torch.float32, device=self.device)
        return self.add_species(mass, charge, arrays, 0, pusher, maxnpart)

    def _pack_species(self):
        arr = (SpeciesC * max(1, len(self.species)))()
        for k, sp in enumerate(self.species):
            arr[k].mass, arr[k].charge = sp.mass, sp.charge
            arr[k].pusher_flags, arr[k].drag_flags = sp.pusher, sp.drag
            arr[k].npart, arr[k].maxnpart = sp.npart, sp.maxnpart
            arr[k].arrays = L.Context.prtls_struct(sp.arrays)
        return arr

    def step(self, nsteps=1, stream=None):
        lib = self.ctx.lib
        if self._species_c is None:
            self._species_c = self._pack_species()
        arr = self._species_c
        st = L.Context._stream(stream)
        for _ in range(nsteps):
            rc = lib.eb200_srpic_step(self.ctx.handle, C.byref(self.params), self.em.data_ptr(),
                                      self.cur.data_ptr(), self.buff.data_ptr(), arr,
                                      len(self.species), self.step_index, self.time, st)
            if rc != 0:
                raise L.EB200Error(lib.eb200_last_error(self.ctx.handle).decode() or f"rc={rc}")
            self.step_index += 1
            self.time += self.dt
        for k, sp in enumerate(self.species):
            sp.npart = int(arr[k].npart)

    # ------------------------------------------------------------------ profiling
    def profile(self, on=True):
        self.ctx._check(self.ctx.lib.eb200_profile_enable(self.ctx.handle, int(on)))

    def read_profile(self):
        """{phase: (milliseconds, calls)} accumulated since the last read."""
        ms = (C.c_float * len(L.PHASES))()
        calls = (C.c_int * len(L.PHASES))()
        self.ctx._check(self.ctx.lib.eb200_profile_read(self.ctx.handle, ms, calls))
        return {nm: (float(ms[k]), int(calls[k])) for k, nm in enumerate(L.PHASES)}

    # ------------------------------------------------------- host-resident state (e2e)
    def host_state(self):
        """Pinned host copies of em, cur and all species arrays (for eb200_srpic_step_host)."""
        torch = self.torch
        pin = lambda t: t.detach().cpu().pin_memory()
        hs = dict(em=pin(self.em), cur=pin(self.cur), species=[])
        for sp in self.species:
            hs["species"].append({k: pin(v) for k, v in sp.arrays.items()})
        arr = (SpeciesC * max(1, len(self.species)))()
        for k, sp in enumerate(self.species):
            arr[k].mass, arr[k].charge = sp.mass, sp.charge
            arr[k].pusher_flags, arr[k].drag_flags = sp.pusher, sp.drag
            arr[k].npart, arr[k].maxnpart = sp.npart, sp.maxnpart
            arr[k].arrays = L.Context.prtls_struct(hs["species"][k])
        hs["c"] = arr
        hs["step"], hs["time"] = self.step_index, self.time
        return hs

    def step_host(self, hs):
        """One step through the host-buffer entry point; returns (h2d_bytes, d2h_bytes)."""
        up, down = C.c_uint64(0), C.c_uint64(0)
        rc = self.ctx.lib.eb200_srpic_step_host(
            self.ctx.handle, C.byref(self.params), hs["em"].data_ptr(), hs["cur"].data_ptr(),
            hs["c"], len(self.species), hs["step"], hs["time"], C.byref(up), C.byref(down))
        if rc != 0:
            raise L.EB200Error(self.ctx.lib.eb200_last_error(self.ctx.handle).decode() or f"rc={rc}")
        hs["step"] += 1
        hs["time"] += self.dt
        return int(up.value), int(down.value)

    # ------------------------------------------------------------ diagnostics (torch)
    def n_pushed(self):
        return sum(sp.npart for sp in self.species if sp.pusher != L.PUSHER_NONE)

    def field_energy(self):
        """Sum of E^2 and B^2 over active cells (plumbing-side diagnostic, not hot path)."""
        g = self.grid
        sl = (slice(None),) + tuple(slice(g.ng, g.ng + g.n[a]) for a in range(g.dim))[::-1]
        act = self.em[sl].double()
        return float((act[:3] ** 2).sum()), float((act[3:] ** 2).sum())
