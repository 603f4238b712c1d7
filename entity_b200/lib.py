"""ctypes binding of include/entity_b200.h. Device memory comes from torch tensors."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EB200_LIB") or os.path.join(HERE, "libentity_b200.so")

PUSHER_NONE, PUSHER_PHOTON, PUSHER_BORIS, PUSHER_VAY, PUSHER_GCA = 0, 1, 2, 4, 8
DRAG_NONE, DRAG_SYNCHROTRON, DRAG_COMPTON = 0, 1, 2
PBC_NONE, PBC_PERIODIC, PBC_ABSORB, PBC_REFLECT, PBC_AXIS = 0, 1, 2, 3, 4
FBC_NONE, FBC_PERIODIC, FBC_CONDUCTOR, FBC_AXIS, FBC_SYNC = 0, 1, 2, 3, 4
FBC_MATCH, FBC_HORIZON, FBC_ATMOSPHERE = 5, 6, 7
DEPOSIT_ATOMIC, DEPOSIT_ORDERED, DEPOSIT_AGGREGATED = 0, 1, 2
STATS_B2, STATS_E2, STATS_EXB, STATS_JDOTE = 0, 1, 2, 3
BC_E, BC_B = 1, 2
STATS_NPART, STATS_N, STATS_RHO, STATS_CHARGE, STATS_T = 0, 1, 2, 3, 4

PHASES = ["FieldSolver", "PushDeposit", "CurrentFiltering", "Communications", "ParticleSort",
          "ParticleMigration"]

PRTL_FIELDS = ["i1", "i2", "i3", "dx1", "dx2", "dx3", "ux1", "ux2", "ux3", "weight",
               "i1_prev", "i2_prev", "i3_prev", "dx1_prev", "dx2_prev", "dx3_prev",
               "tag", "pld_r", "pld_i", "phi"]


class EB200Error(RuntimeError):
    pass


def nghosts_for(order: int) -> int:
    """N_GHOSTS of the reference build (src/global/global.h:130-136)."""
    return 2 if order == 0 else (order + 1) // 2 + 1


class Grid(C.Structure):
    _fields_ = [("dim", C.c_int), ("n", C.c_int * 3), ("ng", C.c_int)]

    @staticmethod
    def make(n, ng):
        g = Grid()
        g.dim = len(n)
        g.n = (C.c_int * 3)(*(list(n) + [1] * (3 - len(n))))
        g.ng = ng
        return g

    def shape(self, ncomp):
        ext = [self.n[a] + 2 * self.ng for a in range(self.dim)]
        return (ncomp, *ext[::-1])


class Prtls(C.Structure):
    _fields_ = [(name, C.c_void_p) for name in PRTL_FIELDS] + [
        ("npld_r", C.c_int), ("npld_i", C.c_int), ("pld_stride", C.c_uint32)]


class Pusher(C.Structure):
    _fields_ = [
        ("pusher_flags", C.c_int), ("drag_flags", C.c_int),
        ("mass", C.c_float), ("charge", C.c_float),
        ("time", C.c_double),
        ("dt", C.c_float), ("omegaB0", C.c_float),
        ("gca_larmor_max", C.c_float), ("gca_e_ovr_b_sqr_max", C.c_float),
        ("sync_coeff", C.c_float), ("compton_coeff", C.c_float),
        ("has_atmosphere", C.c_int),
        ("atm_gx1", C.c_float), ("atm_gx2", C.c_float), ("atm_gx3", C.c_float),
        ("atm_x_surf", C.c_float), ("atm_ds", C.c_float),
        ("pbc", C.c_int * 6),
        ("tag_outgoing", C.c_int),
        ("dx", C.c_float),
        ("xmin", C.c_float * 3),
    ]


class PusherGR(C.Structure):
    """eb200_pusher_gr_t"""
    _fields_ = [("pusher_flags", C.c_int), ("mass", C.c_float), ("charge", C.c_float),
                ("dt", C.c_float), ("omegaB0", C.c_float), ("epsilon", C.c_float),
                ("niter", C.c_int), ("pbc", C.c_int * 6), ("tag_outgoing", C.c_int)]


METRIC_MINKOWSKI, METRIC_SPHERICAL, METRIC_QSPHERICAL = 0, 1, 2
METRIC_KERR_SCHILD, METRIC_QKERR_SCHILD, METRIC_KERR_SCHILD_0 = 3, 4, 5


class Config(C.Structure):
    _fields_ = [
        ("device", C.c_int), ("strict_fp", C.c_int),
        ("grid", Grid), ("shape_order", C.c_int),
        ("metric", C.c_int), ("metric_params", C.c_float * 8),
        ("maxnpart", C.c_uint32),
    ]


class SpeciesC(C.Structure):
    _fields_ = [("mass", C.c_float), ("charge", C.c_float), ("pusher_flags", C.c_int),
                ("drag_flags", C.c_int), ("npart", C.c_uint32), ("maxnpart", C.c_uint32),
                ("arrays", Prtls)]


class MatchFaceC(C.Structure):
    _fields_ = [("o", C.c_int), ("xg_edge", C.c_float), ("ds", C.c_float),
                ("range_min", C.c_int * 3), ("range_max", C.c_int * 3)]


SDIST_UNIFORM, SDIST_TABLE, SDIST_REPLENISH, SDIST_REPLENISH_TABLE, SDIST_ATMOSPHERE = 0, 1, 2, 3, 4


class MaxwellianC(C.Structure):
    """eb200_maxwellian_t"""
    _fields_ = [("temperature", C.c_float), ("drift_u", C.c_float * 3)]


class SpatialDistC(C.Structure):
    """eb200_spatial_dist_t"""
    _fields_ = [("kind", C.c_int), ("field", C.c_void_p), ("comp", C.c_int),
                ("target_density", C.c_float), ("target_field", C.c_void_p),
                ("target_max", C.c_float), ("atm_dim", C.c_int), ("atm_sign", C.c_int),
                ("atm_nmax", C.c_float), ("atm_height", C.c_float), ("atm_xsurf", C.c_float),
                ("atm_ds", C.c_float), ("inv_V0", C.c_float)]


class EmissionC(C.Structure):
    """eb200_emission_t"""
    _fields_ = [("kind", C.c_int), ("photon_weight", C.c_float), ("photon_energy_min", C.c_float),
                ("nominal_probability", C.c_float), ("nominal_photon_energy", C.c_float),
                ("should_drag", C.c_int), ("photons", Prtls), ("photon_npart", C.c_uint32),
                ("photon_maxnpart", C.c_uint32), ("seed", C.c_uint64), ("step", C.c_uint32),
                ("call", C.c_uint32)]


EMISSION_SYNCHROTRON, EMISSION_COMPTON = 1, 2


class AtmosphereC(C.Structure):
    """eb200_atmosphere_t"""
    _fields_ = [("dim", C.c_int), ("sign", C.c_int), ("x_surf", C.c_float), ("ds", C.c_float),
                ("height", C.c_float), ("temperature", C.c_float), ("density", C.c_float),
                ("species", C.c_int * 2), ("inv_n0", C.c_float), ("inv_V0", C.c_float),
                ("ppc0", C.c_float), ("seed", C.c_uint64)]


MAX_MODES = 16


class ExtCurrentC(C.Structure):
    """eb200_ext_current_t"""
    _fields_ = [("nmodes", C.c_int), ("k", (C.c_float * MAX_MODES) * 3),
                ("pref", (C.c_float * MAX_MODES) * 3),
                ("a_real", C.c_float * MAX_MODES), ("a_imag", C.c_float * MAX_MODES),
                ("pref2", (C.c_float * MAX_MODES) * 3),
                ("a_real2", C.c_float * MAX_MODES), ("a_imag2", C.c_float * MAX_MODES)]

    @staticmethod
    def from_table(tab) -> "ExtCurrentC":
        """tab: dict with nmodes, k[3][n], pref[3][n], a_real[n], a_imag[n], pref2[3][n],
        a_real2[n], a_imag2[n] (array-likes)"""
        x = ExtCurrentC()
        n = int(tab["nmodes"])
        x.nmodes = n
        for c in range(3):
            for m in range(n):
                x.k[c][m] = float(tab["k"][c][m])
                x.pref[c][m] = float(tab["pref"][c][m])
                x.pref2[c][m] = float(tab["pref2"][c][m])
        for m in range(n):
            x.a_real[m], x.a_imag[m] = float(tab["a_real"][m]), float(tab["a_imag"][m])
            x.a_real2[m], x.a_imag2[m] = float(tab["a_real2"][m]), float(tab["a_imag2"][m])
        return x


class ParamsC(C.Structure):
    _fields_ = [
        ("dt", C.c_float), ("correction", C.c_float), ("omegaB0", C.c_float),
        ("q0", C.c_float), ("B0", C.c_float), ("V0", C.c_float), ("ppc0", C.c_float),
        ("nfilter", C.c_int), ("fieldsolver_enabled", C.c_int), ("deposit_enabled", C.c_int),
        ("stencil", C.c_float * 9), ("fbc", C.c_int * 6), ("pbc", C.c_int * 6),
        ("gca_larmor_max", C.c_float), ("gca_e_ovr_b_max", C.c_float),
        ("sync_gamma_rad", C.c_float), ("compton_gamma_rad", C.c_float),
        ("fuse_push_deposit", C.c_int), ("deposit_mode", C.c_int),
        ("sort_interval", C.c_int), ("clear_interval", C.c_int),
        ("n0", C.c_float), ("has_atmosphere", C.c_int), ("atm_g", C.c_float * 3),
        ("atm_x_surf", C.c_float), ("atm_ds", C.c_float),
    ]


class FieldBCC(C.Structure):
    """eb200_field_bc_t"""
    _fields_ = [("kind", C.c_int), ("o", C.c_int), ("sign", C.c_int),
                ("xg_edge", C.c_float), ("ds", C.c_float), ("i_edge", C.c_int),
                ("range_min", C.c_int * 2), ("range_max", C.c_int * 2),
                ("target", C.c_void_p), ("mask", C.c_int)]



class GRParamsC(C.Structure):
    """eb200_grpic_params_t"""
    _fields_ = [
        ("dt", C.c_float), ("correction", C.c_float), ("omegaB0", C.c_float),
        ("q0", C.c_float), ("B0", C.c_float),
        ("nfilter", C.c_int), ("fieldsolver_enabled", C.c_int), ("deposit_enabled", C.c_int),
        ("fbc", C.c_int * 6), ("pbc", C.c_int * 6),
        ("pusher_eps", C.c_float), ("pusher_niter", C.c_int),
        ("deposit_mode", C.c_int), ("sort_interval", C.c_int), ("clear_interval", C.c_int),
        ("match_xg_edge", C.c_float), ("match_ds", C.c_float),
        ("match_range_min", C.c_int * 2), ("match_range_max", C.c_int * 2),
        ("match_mask", C.c_int),
    ]


class MetadomainC(C.Structure):
    _fields_ = [("dim", C.c_int), ("rank", C.c_int), ("nranks", C.c_int),
                ("ndoms", C.c_int * 3), ("extents", C.POINTER(C.c_int) * 3),
                ("fbc", C.c_int * 6), ("pbc", C.c_int * 6)]


class DomainInfoC(C.Structure):
    _fields_ = [("offset", C.c_int * 3), ("n", C.c_int * 3), ("cell_offset", C.c_int * 3),
                ("face_fbc", C.c_int * 6), ("face_pbc", C.c_int * 6),
                ("dir_fbc", C.c_int * 27), ("neighbor", C.c_int * 27),
                ("enabled", C.c_int * 27)]


UNIQUE_ID_BYTES = 128

_lib = None


def load():
    """Load libentity_b200.so; there is no fallback when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EB200Error(
            f"{LIB_PATH} is missing: build it with `python -m entity_b200.build` "
            "(entity_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32p = C.c_void_p, C.POINTER(C.c_int)
    ctxp = C.c_void_p
    lib.eb200_version.restype = C.c_int
    lib.eb200_device_count.restype = C.c_int
    lib.eb200_last_error.restype = C.c_char_p
    lib.eb200_last_error.argtypes = [ctxp]
    lib.eb200_launch_count.restype = C.c_uint64
    lib.eb200_launch_count.argtypes = [ctxp]
    lib.eb200_init.argtypes = [C.POINTER(Config), C.POINTER(ctxp)]
    lib.eb200_finalize.argtypes = [ctxp]
    lib.eb200_finalize.restype = None
    lib.eb200_faraday.argtypes = [ctxp, vp, C.c_float, C.c_float, vp, vp]
    lib.eb200_ampere.argtypes = [ctxp, vp, C.c_float, C.c_float, vp]
    lib.eb200_currents_ampere.argtypes = [ctxp, vp, vp, C.c_float, C.c_float, vp]
    lib.eb200_filter.argtypes = [ctxp, vp, vp, C.c_int, i32p, vp]
    lib.eb200_push_sr.argtypes = [ctxp, C.POINTER(Pusher), C.POINTER(Prtls), C.c_uint32, vp, vp]
    lib.eb200_deposit.argtypes = [ctxp, C.POINTER(Prtls), C.c_uint32, C.c_float, C.c_float, vp,
                                  C.c_int, vp]
    lib.eb200_push_deposit_sr.argtypes = [ctxp, C.POINTER(Pusher), C.POINTER(Prtls), C.c_uint32,
                                          vp, vp, C.c_int, vp]
    lib.eb200_zero_currents.argtypes = [ctxp, vp, vp]
    lib.eb200_set_pd_kernel.argtypes = [ctxp, C.c_int]
    f32p = C.POINTER(C.c_float)
    lib.eb200_metric_eval.argtypes = [C.c_int, i32p, f32p, C.c_int, f32p, f32p, f32p]
    lib.eb200_faraday_sr.argtypes = [ctxp, vp, C.c_float, i32p, vp]
    lib.eb200_ampere_sr.argtypes = [ctxp, vp, C.c_float, i32p, vp]
    lib.eb200_currents_ampere_sr.argtypes = [ctxp, vp, vp, C.c_float, C.c_float, i32p, vp]
    lib.eb200_push_gr.argtypes = [ctxp, C.POINTER(PusherGR), C.POINTER(Prtls), C.c_uint32, vp,
                                  vp, vp]
    lib.eb200_gr_aux_e.argtypes = [ctxp, vp, vp, vp, i32p, vp]
    lib.eb200_gr_aux_h.argtypes = [ctxp, vp, vp, vp, i32p, vp]
    lib.eb200_faraday_gr.argtypes = [ctxp, vp, vp, vp, C.c_float, i32p, vp]
    lib.eb200_ampere_gr.argtypes = [ctxp, vp, vp, vp, C.c_float, i32p, vp]
    lib.eb200_currents_ampere_gr.argtypes = [ctxp, vp, vp, C.c_float, i32p, vp]
    lib.eb200_time_average.argtypes = [ctxp, vp, vp, C.c_int, vp]
    lib.eb200_comm_fields.argtypes = [ctxp, vp, C.c_int, C.c_int, C.c_int, i32p, vp]
    lib.eb200_sync_currents.argtypes = [ctxp, vp, vp, i32p, vp]
    lib.eb200_sort_particles.argtypes = [ctxp, C.POINTER(Prtls), C.POINTER(C.c_uint32), C.c_int, vp]
    lib.eb200_push_sr_emission.argtypes = [ctxp, C.POINTER(Pusher), C.POINTER(Prtls), C.c_uint32, vp,
                                           C.POINTER(EmissionC), vp]
    lib.eb200_push_sr_emission.restype = C.c_int
    lib.eb200_srpic_set_emission.argtypes = [ctxp, C.c_int, C.c_int, C.POINTER(EmissionC)]
    lib.eb200_srpic_set_emission.restype = C.c_int
    lib.eb200_set_lean_prev.argtypes = [ctxp, C.c_int]
    lib.eb200_set_sort_mode.argtypes = [ctxp, C.c_int]
    lib.eb200_srpic_step.argtypes = [ctxp, C.POINTER(ParamsC), vp, vp, vp, C.POINTER(SpeciesC),
                                     C.c_int, C.c_uint32, C.c_double, vp]
    lib.eb200_srpic_step.restype = C.c_int
    lib.eb200_profile_enable.argtypes = [ctxp, C.c_int]
    lib.eb200_profile_enable.restype = C.c_int
    lib.eb200_profile_read.argtypes = [ctxp, C.POINTER(C.c_float), C.POINTER(C.c_int)]
    lib.eb200_profile_read.restype = C.c_int
    lib.eb200_srpic_step_host.argtypes = [ctxp, C.POINTER(ParamsC), vp, vp, C.POINTER(SpeciesC),
                                          C.c_int, C.c_uint32, C.c_double,
                                          C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.eb200_srpic_step_host.restype = C.c_int
    lib.eb200_match_fields.argtypes = [ctxp, vp, vp, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int,
                                       i32p, i32p, vp]
    lib.eb200_match_fields.restype = C.c_int
    lib.eb200_match_layer.argtypes = [C.POINTER(Grid), C.c_float, C.POINTER(C.c_float),
                                      C.POINTER(C.c_float), C.c_float, C.c_float, C.c_int, C.c_int,
                                      C.c_float, C.POINTER(MatchFaceC)]
    lib.eb200_match_layer.restype = C.c_int
    lib.eb200_srpic_set_match.argtypes = [ctxp, C.POINTER(MatchFaceC), C.c_int, vp, C.c_int]
    lib.eb200_srpic_set_match.restype = C.c_int
    lib.eb200_stats_fields.argtypes = [ctxp, vp, vp, C.c_int, C.c_int, C.POINTER(C.c_double), vp]
    lib.eb200_stats_fields.restype = C.c_int
    lib.eb200_stats_particles.argtypes = [ctxp, C.POINTER(Prtls), C.c_uint32, C.c_float, C.c_float,
                                          C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(C.c_double), vp]
    lib.eb200_stats_particles.restype = C.c_int
    lib.eb200_pack_fields_hold.argtypes = [ctxp, vp, vp]
    lib.eb200_pack_fields_hold.restype = C.c_int
    lib.eb200_pack_fields_release.argtypes = [ctxp]
    lib.eb200_pack_fields_release.restype = C.c_int
    lib.eb200_decompose.argtypes = [C.c_int, C.c_int, i32p, i32p, i32p, i32p, i32p, i32p]
    lib.eb200_domain_info.argtypes = [C.POINTER(MetadomainC), C.POINTER(DomainInfoC)]
    lib.eb200_comm_unique_id.argtypes = [C.c_char_p]
    lib.eb200_comm_init.argtypes = [ctxp, C.POINTER(MetadomainC), C.c_char_p]
    lib.eb200_comm_particles.argtypes = [ctxp, C.POINTER(SpeciesC), C.c_int, vp]
    for name in ("decompose", "domain_info", "comm_unique_id", "comm_init", "comm_particles"):
        getattr(lib, "eb200_" + name).restype = C.c_int
    for name in ("init", "faraday", "ampere", "currents_ampere", "filter", "push_sr", "deposit",
                 "push_deposit_sr", "zero_currents", "comm_fields", "sync_currents",
                 "sort_particles"):
        getattr(lib, "eb200_" + name).restype = C.c_int
    _lib = lib
    return lib


def exported_symbols():
    """Names the header declares (used by the CPU-side ABI test)."""
    import re
    hdr = os.path.join(os.path.dirname(HERE), "include", "entity_b200.h")
    txt = open(hdr).read()
    return sorted(set(re.findall(r"\b(eb200_[a-z_0-9]+)\s*\(", txt)))


def match_layer(grid, dx, local_xmin, local_xmax, global_xmin_o, global_xmax_o, o, sign, ds):
    """srpic::MatchFieldsIn geometry for one face of the global box (pure host code): returns
    (o, xg_edge, ds, range_min, range_max) for Simulation.set_match / Context.match_fields, or
    None when the layer does not reach this domain."""
    lib = load()
    d = grid.dim
    lo = (C.c_float * 3)(*(list(local_xmin)[:d] + [0.0] * (3 - d)))
    hi = (C.c_float * 3)(*(list(local_xmax)[:d] + [0.0] * (3 - d)))
    f = MatchFaceC()
    rc = lib.eb200_match_layer(C.byref(grid), dx, lo, hi, global_xmin_o, global_xmax_o, o, sign, ds,
                               C.byref(f))
    if rc < 0:
        raise EB200Error(f"eb200_match_layer: bad argument (rc={rc})")
    if rc == 0:
        return None
    return (f.o, f.xg_edge, f.ds, list(f.range_min)[:d], list(f.range_max)[:d])


def decompose(ndomains, ncells, decomposition=None):
    """tools::Decompose through the library's host code: returns the list, per dimension, of
    the active-cell extents of the domains along it."""
    lib = load()
    dim = len(ncells)
    dec = list(decomposition) if decomposition is not None else [-1] * dim
    nd = (C.c_int * 3)()
    outs = [(C.c_int * max(ndomains, 1))() for _ in range(3)]
    rc = lib.eb200_decompose(ndomains, dim, (C.c_int * dim)(*ncells), (C.c_int * dim)(*dec), nd,
                             outs[0], outs[1], outs[2])
    if rc != 0:
        raise EB200Error(f"Decomposition error for {ndomains} domains over {list(ncells)} "
                         f"with {dec}")
    return [[int(outs[a][k]) for k in range(nd[a])] for a in range(dim)]


def make_metadomain(rank, extents, fbc=None, pbc=None):
    """eb200_metadomain_t for `rank` of the Cartesian product of `extents` (list per dimension)."""
    dim = len(extents)
    md = MetadomainC()
    md.dim, md.rank = dim, rank
    md.nranks = 1
    keep = []
    for a in range(3):
        if a < dim:
            arr = (C.c_int * len(extents[a]))(*extents[a])
            keep.append(arr)
            md.extents[a] = C.cast(arr, C.POINTER(C.c_int))
            md.ndoms[a] = len(extents[a])
            md.nranks *= len(extents[a])
        else:
            md.ndoms[a] = 1
    md.fbc = (C.c_int * 6)(*(fbc or [FBC_PERIODIC] * 6))
    md.pbc = (C.c_int * 6)(*(pbc or [PBC_PERIODIC] * 6))
    md._keep = keep  # the C struct borrows these arrays
    return md


def domain_info(md) -> DomainInfoC:
    info = DomainInfoC()
    if load().eb200_domain_info(C.byref(md), C.byref(info)) != 0:
        raise EB200Error("invalid metadomain description")
    return info


def unique_id() -> bytes:
    """ncclGetUniqueId through the library (needs NCCL, i.e. a GPU box)."""
    buf = C.create_string_buffer(UNIQUE_ID_BYTES)
    lib = load()
    if lib.eb200_comm_unique_id(buf) != 0:
        raise EB200Error(lib.eb200_last_error(None).decode())
    return buf.raw


def metric_eval(metric, n, metric_params, x1, x2):
    """Host evaluation of the metric functions the kernels use (eb200_metric_eval): numpy in,
    numpy out ([nq][16] for the SR metrics, [nq][32] for the GR ones)."""
    import numpy as np
    x1 = np.ascontiguousarray(x1, np.float32)
    x2 = np.ascontiguousarray(x2, np.float32)
    out = np.zeros((x1.size, 16 if metric <= METRIC_QSPHERICAL else 32), np.float32)
    f32p = C.POINTER(C.c_float)
    rc = load().eb200_metric_eval(metric, (C.c_int * 2)(*n),
                                  (C.c_float * 8)(*(list(metric_params) + [0.0] * 8)[:8]),
                                  x1.size, x1.ctypes.data_as(f32p), x2.ctypes.data_as(f32p),
                                  out.ctypes.data_as(f32p))
    if rc != 0:
        raise EB200Error(load().eb200_last_error(None).decode())
    return out


def _ptr(t):
    return None if t is None else t.data_ptr()


class Context:
    """One eb200 context (one local domain on one device)."""

    def __init__(self, n, order=0, strict=False, device=0, dx=1.0, xmin=(0.0, 0.0, 0.0),
                 ng=None, maxnpart=0, metric=METRIC_MINKOWSKI, metric_params=None):
        self.lib = load()
        self.order = order
        self.grid = Grid.make(n, nghosts_for(order) if ng is None else ng)
        cfg = Config()
        cfg.device = device
        cfg.strict_fp = int(strict)
        cfg.grid = self.grid
        cfg.shape_order = order
        cfg.metric = metric
        if metric == METRIC_MINKOWSKI:
            cfg.metric_params = (C.c_float * 8)(dx, *xmin, 0, 0, 0, 0)
        else:
            # x1min, x1max, x2min, x2max, qsph_r0, qsph_h, ks_a
            cfg.metric_params = (C.c_float * 8)(*(list(metric_params) + [0.0] * 8)[:8])
        self.metric = metric
        cfg.maxnpart = maxnpart
        self.dx = dx
        self.xmin = tuple(xmin)
        self.handle = C.c_void_p()
        rc = self.lib.eb200_init(C.byref(cfg), C.byref(self.handle))
        if rc != 0:
            raise EB200Error(self.lib.eb200_last_error(None).decode())
        which = int(os.environ.get("EB200_PD_KERNEL", "0"))
        if which:
            self.set_pd_kernel(which)

    def set_pd_kernel(self, which: int):
        """0 auto, 1 one particle per thread, 2 TMA-staged chunks, 3 four particles per thread."""
        self._check(self.lib.eb200_set_pd_kernel(self.handle, which))

    def set_lean_prev(self, on: bool):
        """i*_prev / dx*_prev are scratch of one step: not stored by the fused kernel, not sorted,
        not moved by step_host (eb200_set_lean_prev)"""
        self._check(self.lib.eb200_set_lean_prev(self.handle, int(on)))

    def set_sort_mode(self, mode: int):
        """-1 by build, 0 stable radix sort, 1 counting sort (eb200_set_sort_mode)"""
        self._check(self.lib.eb200_set_sort_mode(self.handle, mode))

    def close(self):
        if self.handle:
            self.lib.eb200_finalize(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise EB200Error(self.lib.eb200_last_error(self.handle).decode())

    @property
    def launch_count(self):
        return int(self.lib.eb200_launch_count(self.handle))

    @staticmethod
    def _stream(stream):
        if stream is None:
            import torch
            return C.c_void_p(torch.cuda.current_stream().cuda_stream)
        return C.c_void_p(stream)

    def make_pusher(self, **kw) -> Pusher:
        p = Pusher()
        p.pusher_flags = kw.get("pusher_flags", PUSHER_BORIS)
        p.drag_flags = kw.get("drag_flags", DRAG_NONE)
        p.mass = kw.get("mass", 1.0)
        p.charge = kw.get("charge", -1.0)
        p.time = kw.get("time", 0.0)
        p.dt = kw["dt"]
        p.omegaB0 = kw.get("omegaB0", 1.0)
        for k in ("gca_larmor_max", "gca_e_ovr_b_sqr_max", "sync_coeff", "compton_coeff",
                  "atm_gx1", "atm_gx2", "atm_gx3", "atm_x_surf", "atm_ds"):
            setattr(p, k, kw.get(k, 0.0))
        p.has_atmosphere = kw.get("has_atmosphere", 0)
        p.pbc = (C.c_int * 6)(*kw.get("pbc", [PBC_PERIODIC] * 6))
        p.tag_outgoing = kw.get("tag_outgoing", 0)
        p.dx = kw.get("dx", self.dx)
        p.xmin = (C.c_float * 3)(*kw.get("xmin", self.xmin))
        return p

    @staticmethod
    def prtls_struct(arrays: dict) -> Prtls:
        s = Prtls()
        for name in PRTL_FIELDS:
            t = arrays.get(name)
            setattr(s, name, _ptr(t) if t is not None and t.numel() else None)
        # payload planes: tensors of shape [npld, capacity] (plane k contiguous)
        for nm, cnt in (("pld_r", "npld_r"), ("pld_i", "npld_i")):
            t = arrays.get(nm)
            if t is not None and t.numel():
                setattr(s, cnt, int(t.shape[0]))
                s.pld_stride = int(t.shape[1])
        return s

    # -- field solvers
    def faraday(self, em, coeff1, coeff2, stencil=None, stream=None):
        st = None
        if stencil is not None:
            st = (C.c_float * 9)(*stencil)
        self._check(self.lib.eb200_faraday(self.handle, _ptr(em), coeff1, coeff2,
                                           C.cast(st, C.c_void_p) if st is not None else None,
                                           self._stream(stream)))

    def ampere(self, em, coeff1, coeff2, stream=None):
        self._check(self.lib.eb200_ampere(self.handle, _ptr(em), coeff1, coeff2,
                                          self._stream(stream)))

    def currents_ampere(self, em, cur, coeff, ppc0, stream=None):
        self._check(self.lib.eb200_currents_ampere(self.handle, _ptr(em), _ptr(cur), coeff, ppc0,
                                                   self._stream(stream)))

    def filter(self, cur, buff, nfilter, fbc, stream=None):
        self._check(self.lib.eb200_filter(self.handle, _ptr(cur), _ptr(buff), nfilter,
                                          (C.c_int * 6)(*fbc), self._stream(stream)))

    # -- curvilinear SR field solvers (spherical / qspherical contexts)
    def faraday_sr(self, em, coeff, fbc, stream=None):
        self._check(self.lib.eb200_faraday_sr(self.handle, _ptr(em), coeff, (C.c_int * 6)(*fbc),
                                              self._stream(stream)))

    def ampere_sr(self, em, coeff, fbc, stream=None):
        self._check(self.lib.eb200_ampere_sr(self.handle, _ptr(em), coeff, (C.c_int * 6)(*fbc),
                                             self._stream(stream)))

    def currents_ampere_sr(self, em, cur, coeff, inv_n0, fbc, stream=None):
        self._check(self.lib.eb200_currents_ampere_sr(self.handle, _ptr(em), _ptr(cur), coeff,
                                                      inv_n0, (C.c_int * 6)(*fbc),
                                                      self._stream(stream)))

    # -- GRPIC
    @staticmethod
    def make_pusher_gr(**kw) -> PusherGR:
        p = PusherGR()
        p.pusher_flags = kw.get("pusher_flags", PUSHER_BORIS)
        p.mass, p.charge = kw.get("mass", 1.0), kw.get("charge", -1.0)
        p.dt, p.omegaB0 = kw["dt"], kw.get("omegaB0", 1.0)
        p.epsilon, p.niter = kw.get("epsilon", 1e-2), kw.get("niter", 10)
        p.pbc = (C.c_int * 6)(*kw.get("pbc", [PBC_ABSORB, PBC_ABSORB, PBC_AXIS, PBC_AXIS, 0, 0]))
        p.tag_outgoing = kw.get("tag_outgoing", 0)
        return p

    def push_gr(self, pusher, arrays, npart, em, em0, stream=None):
        s = self.prtls_struct(arrays)
        self._check(self.lib.eb200_push_gr(self.handle, C.byref(pusher), C.byref(s), npart,
                                           _ptr(em), _ptr(em0), self._stream(stream)))

    def gr_aux_e(self, d, b, out, fbc, stream=None):
        self._check(self.lib.eb200_gr_aux_e(self.handle, _ptr(d), _ptr(b), _ptr(out),
                                            (C.c_int * 6)(*fbc), self._stream(stream)))

    def gr_aux_h(self, d, b, out, fbc, stream=None):
        self._check(self.lib.eb200_gr_aux_h(self.handle, _ptr(d), _ptr(b), _ptr(out),
                                            (C.c_int * 6)(*fbc), self._stream(stream)))

    def faraday_gr(self, b_in, b_out, e_aux, coeff, fbc, stream=None):
        self._check(self.lib.eb200_faraday_gr(self.handle, _ptr(b_in), _ptr(b_out), _ptr(e_aux),
                                              coeff, (C.c_int * 6)(*fbc), self._stream(stream)))

    def ampere_gr(self, d_in, d_out, h_aux, coeff, fbc, stream=None):
        self._check(self.lib.eb200_ampere_gr(self.handle, _ptr(d_in), _ptr(d_out), _ptr(h_aux),
                                             coeff, (C.c_int * 6)(*fbc), self._stream(stream)))

    def currents_ampere_gr(self, d, cur, coeff, fbc, stream=None):
        self._check(self.lib.eb200_currents_ampere_gr(self.handle, _ptr(d), _ptr(cur), coeff,
                                                      (C.c_int * 6)(*fbc), self._stream(stream)))

    def time_average(self, a, b, stream=None):
        self._check(self.lib.eb200_time_average(self.handle, _ptr(a), _ptr(b), a.shape[0],
                                                self._stream(stream)))

    # -- particles
    def push(self, pusher, arrays, npart, em, stream=None):
        s = self.prtls_struct(arrays)
        self._check(self.lib.eb200_push_sr(self.handle, C.byref(pusher), C.byref(s), npart,
                                           _ptr(em), self._stream(stream)))

    def push_emission(self, pusher, arrays, npart, em, kind, photons, photon_npart, photon_maxnpart,
                      photon_weight, photon_energy_min, nominal_probability, nominal_photon_energy,
                      should_drag=False, seed=0x123456789abcdef0, step=0, call=0, stream=None):
        """eb200_push_sr_emission; returns the new particle count of the emitted species"""
        s = self.prtls_struct(arrays)
        e = EmissionC()
        e.kind, e.photon_weight, e.photon_energy_min = kind, photon_weight, photon_energy_min
        e.nominal_probability, e.nominal_photon_energy = nominal_probability, nominal_photon_energy
        e.should_drag = int(should_drag)
        e.photons = self.prtls_struct(photons)
        e.photon_npart, e.photon_maxnpart = photon_npart, photon_maxnpart
        e.seed, e.step, e.call = seed, step, call
        rc = self.lib.eb200_push_sr_emission(self.handle, C.byref(pusher), C.byref(s), npart, _ptr(em),
                                             C.byref(e), self._stream(stream))
        self.last_emission_npart = int(e.photon_npart)
        self._check(rc)
        return int(e.photon_npart)

    def deposit(self, arrays, npart, charge, dt, cur, mode=DEPOSIT_ATOMIC, stream=None):
        s = self.prtls_struct(arrays)
        self._check(self.lib.eb200_deposit(self.handle, C.byref(s), npart, charge, dt, _ptr(cur),
                                           mode, self._stream(stream)))

    def push_deposit(self, pusher, arrays, npart, em, cur, mode=DEPOSIT_ATOMIC, stream=None):
        s = self.prtls_struct(arrays)
        self._check(self.lib.eb200_push_deposit_sr(self.handle, C.byref(pusher), C.byref(s),
                                                   npart, _ptr(em), _ptr(cur), mode,
                                                   self._stream(stream)))

    # -- matching field boundaries (fields_bcs.hpp MatchBoundaries_kernel)
    def match_fields(self, em, target, o, xg_edge, ds, tags, mask, range_min, range_max, stream=None):
        d = len(range_min)
        lo, hi = (C.c_int * 3)(*(list(range_min) + [0] * (3 - d))), \
            (C.c_int * 3)(*(list(range_max) + [1] * (3 - d)))
        self._check(self.lib.eb200_match_fields(self.handle, _ptr(em), _ptr(target), o, xg_edge, ds,
                                                tags, mask, lo, hi, self._stream(stream)))

    # -- reduced statistics (reduced_stats.hpp): local sums, host values
    def stats_fields(self, em, cur, what, comp=1, stream=None) -> float:
        out = C.c_double(0.0)
        self._check(self.lib.eb200_stats_fields(self.handle, _ptr(em),
                                                _ptr(cur) if cur is not None else None, what, comp,
                                                C.byref(out), self._stream(stream)))
        return out.value

    def stats_particles(self, arrays, npart, mass, charge, what, c1=0, c2=0, use_weights=False,
                        stream=None) -> float:
        out = C.c_double(0.0)
        s = self.prtls_struct(arrays)
        self._check(self.lib.eb200_stats_particles(self.handle, C.byref(s), npart, mass, charge,
                                                   1 if use_weights else 0, what, c1, c2,
                                                   C.byref(out), self._stream(stream)))
        return out.value

    def zero_currents(self, cur, stream=None):
        self._check(self.lib.eb200_zero_currents(self.handle, _ptr(cur), self._stream(stream)))

    def comm_fields(self, fld, c0, c1, fbc, stream=None):
        self._check(self.lib.eb200_comm_fields(self.handle, _ptr(fld), fld.shape[0], c0, c1,
                                               (C.c_int * 6)(*fbc), self._stream(stream)))

    def sync_currents(self, cur, buff, fbc, stream=None):
        self._check(self.lib.eb200_sync_currents(self.handle, _ptr(cur), _ptr(buff),
                                                 (C.c_int * 6)(*fbc), self._stream(stream)))

    def comm_init(self, md, uid: bytes | None):
        """Attach the decomposition (and, with `uid`, an NCCL communicator) to this context."""
        self._md = md
        self._check(self.lib.eb200_comm_init(self.handle, C.byref(md), uid))

    def sort_particles(self, arrays, npart, remove_dead=False, stream=None):
        s = self.prtls_struct(arrays)
        n = C.c_uint32(npart)
        self._check(self.lib.eb200_sort_particles(self.handle, C.byref(s), C.byref(n),
                                                  int(remove_dead), self._stream(stream)))
        return int(n.value)
