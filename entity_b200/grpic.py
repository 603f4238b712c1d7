"""Python driver of the GRPIC step (``eb200_grpic_step``): owns the device tensors of one 2D
Kerr-Schild type domain (em, em0, cur, cur0, aux, buff + species) and calls the C++ engine
mirror in ``csrc/engine.cu``. Plumbing only: no arithmetic on field or particle data."""
from __future__ import annotations

import ctypes as C

from . import lib as L
from .lib import GRParamsC, SpeciesC
from .srpic import PRTL_DTYPES, Species


class GRSimulation:
    """One 2D GRPIC domain: x1 = {HORIZON, MATCH}, x2 = {AXIS, AXIS}, as the reference sets a
    Kerr-Schild mesh up (grid.cpp: fields [{horizon, match}, {axis, axis}])."""

    def __init__(self, n, metric, metric_params, dt, omegaB0, q0, B0, order=0, nfilter=0,
                 correction=1.0, pusher_eps=1e-6, pusher_niter=10, fieldsolver=True, deposit=True,
                 deposit_mode=L.DEPOSIT_ATOMIC, clear_interval=0, sort_interval=0, device=0):
        import torch
        self.torch = torch
        self.device = torch.device("cuda", device)
        self.ctx = L.Context(n, order=order, strict=False, device=device, metric=metric,
                             metric_params=metric_params)
        self.grid = self.ctx.grid
        g = self.grid
        self.shape6, self.shape3 = g.shape(6), g.shape(3)
        z = lambda shp: torch.zeros(shp, dtype=torch.float32, device=self.device)
        self.em, self.em0, self.aux = z(self.shape6), z(self.shape6), z(self.shape6)
        self.cur, self.cur0, self.buff = z(self.shape3), z(self.shape3), z(self.shape3)
        self.species: list[Species] = []
        self._species_c = None
        self.step_index = 0
        self.time = 0.0
        p = GRParamsC()
        p.dt, p.correction, p.omegaB0, p.q0, p.B0 = dt, correction, omegaB0, q0, B0
        p.nfilter = nfilter
        p.fieldsolver_enabled, p.deposit_enabled = int(fieldsolver), int(deposit)
        p.fbc = (C.c_int * 6)(L.FBC_HORIZON, L.FBC_MATCH, L.FBC_AXIS, L.FBC_AXIS, 0, 0)
        p.pbc = (C.c_int * 6)(L.PBC_ABSORB, L.PBC_ABSORB, L.PBC_AXIS, L.PBC_AXIS, 0, 0)
        p.pusher_eps, p.pusher_niter = pusher_eps, pusher_niter
        p.deposit_mode = deposit_mode
        p.sort_interval, p.clear_interval = sort_interval, clear_interval
        self.params = p
        self.match_target = None

    @property
    def dt(self):
        return float(self.params.dt)

    def set_match(self, target, mask, xg_edge, ds, range_min, range_max):
        """the +x1 MATCH layer: target = pgen.init_flds on every component's node (device tensor,
        layout of em)"""
        self.match_target = target
        p = self.params
        p.match_xg_edge, p.match_ds, p.match_mask = xg_edge, ds, mask
        p.match_range_min = (C.c_int * 2)(*range_min)
        p.match_range_max = (C.c_int * 2)(*range_max)

    def alloc_species(self, mass, charge, maxnpart, pusher=L.PUSHER_BORIS):
        torch = self.torch
        arrays = {}
        for k in list(PRTL_DTYPES) + ["phi"]:
            axis = [c for c in k if c in "123"]
            if k.startswith(("i", "dx")) and axis and int(axis[0]) > 2:
                continue
            dt = PRTL_DTYPES.get(k, "float32")
            arrays[k] = torch.zeros(maxnpart, dtype=getattr(torch, dt), device=self.device)
        sp = Species(mass, charge, pusher, L.DRAG_NONE, 0, maxnpart, arrays)
        self.species.append(sp)
        self._species_c = None
        return sp

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
            names = ["em", "em0", "cur", "cur0"]
            ptrs = [C.c_void_p(getattr(self, nm).data_ptr()) for nm in names]
            before = {p.value: getattr(self, nm) for p, nm in zip(ptrs, names)}
            rc = lib.eb200_grpic_step(
                self.ctx.handle, C.byref(self.params), C.byref(ptrs[0]), C.byref(ptrs[1]),
                C.byref(ptrs[2]), C.byref(ptrs[3]), C.c_void_p(self.aux.data_ptr()),
                C.c_void_p(self.buff.data_ptr()),
                C.c_void_p(self.match_target.data_ptr()) if self.match_target is not None else None,
                arr, len(self.species), self.step_index, C.c_double(self.time), st)
            if rc != 0:
                raise L.EB200Error(lib.eb200_last_error(self.ctx.handle).decode() or f"rc={rc}")
            # SwapFields exchanged the pointers: follow it with the tensors
            for p, nm in zip(ptrs, names):
                setattr(self, nm, before[p.value])
            self.step_index += 1
            self.time += self.dt
        for k, sp in enumerate(self.species):
            sp.npart = int(arr[k].npart)

    def n_pushed(self):
        return sum(sp.npart for sp in self.species if sp.pusher != L.PUSHER_NONE)

    def profile(self, on=True):
        self.ctx._check(self.ctx.lib.eb200_profile_enable(self.ctx.handle, int(on)))

    def read_profile(self):
        ms = (C.c_float * len(L.PHASES))()
        calls = (C.c_int * len(L.PHASES))()
        self.ctx._check(self.ctx.lib.eb200_profile_read(self.ctx.handle, ms, calls))
        return {nm: (float(ms[k]), int(calls[k])) for k, nm in enumerate(L.PHASES)}
