"""TEST INFRASTRUCTURE ONLY -- ctypes front end of the CPU oracle.

Loads ``oracle/liborc.so`` (the scalar fp32 restatement, ``oracle_kernels.cpp``) and,
when present, ``oracle/_ref/libref_o<O>.so`` (the reference's own kernel headers compiled
in place, ``ref_driver.cpp``). Both export the same signatures (``orc_*`` / ``ref_*``) so
tests can run them side by side on identical numpy buffers.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# --------------------------------------------------------------------------- enums
PUSHER_NONE, PUSHER_PHOTON, PUSHER_BORIS, PUSHER_VAY, PUSHER_GCA = 0, 1, 2, 4, 8
DRAG_NONE, DRAG_SYNCHROTRON, DRAG_COMPTON = 0, 1, 2
PBC_NONE, PBC_PERIODIC, PBC_ABSORB, PBC_REFLECT, PBC_AXIS = 0, 1, 2, 3, 4
FBC_NONE, FBC_PERIODIC, FBC_CONDUCTOR, FBC_AXIS, FBC_SYNC = 0, 1, 2, 3, 4


def nghosts_for(order: int) -> int:
    """N_GHOSTS of the reference build: src/global/global.h:130-136."""
    return 2 if order == 0 else (order + 1) // 2 + 1


class Grid(C.Structure):
    _fields_ = [("dim", C.c_int), ("n", C.c_int * 3), ("ng", C.c_int)]

    @staticmethod
    def make(n, ng):
        g = Grid()
        g.dim = len(n)
        nn = list(n) + [1] * (3 - len(n))
        g.n = (C.c_int * 3)(*nn)
        g.ng = ng
        return g

    def shape(self, ncomp):
        """numpy shape of a field in C order == LayoutLeft with i1 fastest."""
        ext = [self.n[a] + 2 * self.ng for a in range(self.dim)]
        return (ncomp, *ext[::-1])


PRTL_FIELDS = [
    ("i1", np.int32), ("i2", np.int32), ("i3", np.int32),
    ("dx1", np.float32), ("dx2", np.float32), ("dx3", np.float32),
    ("ux1", np.float32), ("ux2", np.float32), ("ux3", np.float32),
    ("weight", np.float32),
    ("i1_prev", np.int32), ("i2_prev", np.int32), ("i3_prev", np.int32),
    ("dx1_prev", np.float32), ("dx2_prev", np.float32), ("dx3_prev", np.float32),
    ("tag", np.int16),
    ("pld_r", np.float32), ("pld_i", np.uint32),
    ("phi", np.float32),
]


class Prtls(C.Structure):
    _fields_ = [(name, C.c_void_p) for name, _ in PRTL_FIELDS]


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


def make_pusher(**kw) -> Pusher:
    p = Pusher()
    p.pusher_flags = kw.get("pusher_flags", PUSHER_BORIS)
    p.drag_flags = kw.get("drag_flags", DRAG_NONE)
    p.mass = kw.get("mass", 1.0)
    p.charge = kw.get("charge", -1.0)
    p.time = kw.get("time", 0.0)
    p.dt = kw["dt"]
    p.omegaB0 = kw.get("omegaB0", 1.0)
    p.gca_larmor_max = kw.get("gca_larmor_max", 0.0)
    p.gca_e_ovr_b_sqr_max = kw.get("gca_e_ovr_b_sqr_max", 0.0)
    p.sync_coeff = kw.get("sync_coeff", 0.0)
    p.compton_coeff = kw.get("compton_coeff", 0.0)
    p.has_atmosphere = kw.get("has_atmosphere", 0)
    for k in ("atm_gx1", "atm_gx2", "atm_gx3", "atm_x_surf", "atm_ds"):
        setattr(p, k, kw.get(k, 0.0))
    p.pbc = (C.c_int * 6)(*kw.get("pbc", [PBC_PERIODIC] * 6))
    p.tag_outgoing = kw.get("tag_outgoing", 0)
    p.dx = kw.get("dx", 1.0)
    p.xmin = (C.c_float * 3)(*kw.get("xmin", [0.0, 0.0, 0.0]))
    return p


class ParticleSet:
    """SoA particle arrays as numpy; same member order as ParticleArrays."""

    def __init__(self, n: int):
        self.n = n
        for name, dt in PRTL_FIELDS:
            if name in ("pld_r", "pld_i"):
                setattr(self, name, np.zeros(0, dtype=dt))
            else:
                setattr(self, name, np.zeros(n, dtype=dt))

    def struct(self) -> Prtls:
        s = Prtls()
        for name, _ in PRTL_FIELDS:
            a = getattr(self, name)
            setattr(s, name, a.ctypes.data if a.size else None)
        return s

    def copy(self) -> "ParticleSet":
        q = ParticleSet(self.n)
        for name, _ in PRTL_FIELDS:
            setattr(q, name, getattr(self, name).copy())
        return q

    def names(self):
        return [n for n, _ in PRTL_FIELDS if n not in ("pld_r", "pld_i")]


# --------------------------------------------------------------------------- loading
def build(ref: bool | None = None) -> None:
    """Compile liborc.so (always) and _ref/libref_o*.so (when /root/reference exists)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "liborc.so"])
    if ref is None:
        ref = os.path.isdir(os.environ.get("EB200_REFERENCE", "/root/reference"))
    if ref:
        subprocess.check_call(
            ["make", "-s", "-j4", "-C", HERE, "ref",
             "REF=" + os.environ.get("EB200_REFERENCE", "/root/reference")])


def _declare(lib, prefix):
    f32p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int)
    G, P, U = C.POINTER(Grid), C.POINTER(Prtls), C.POINTER(Pusher)
    sig = {
        "faraday_mink": [G, C.c_void_p, C.c_float, C.c_float, C.c_void_p],
        "ampere_mink": [G, C.c_void_p, C.c_float, C.c_float],
        "currents_ampere_mink": [G, C.c_void_p, C.c_void_p, C.c_float, C.c_float],
        "filter_pass": [G, C.c_void_p, C.c_void_p, i32p],
        "push_sr_mink": [G, C.c_int, U, P, C.c_uint32, C.c_void_p],
        "deposit_mink": [G, C.c_int, P, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_void_p],
        "comm_fields_self": [G, C.c_void_p, C.c_int, C.c_int, C.c_int, i32p],
        "sync_currents_self": [G, C.c_void_p, C.c_void_p, i32p],
    }
    del f32p
    for name, args in sig.items():
        fn = getattr(lib, prefix + name, None)
        if fn is not None:
            fn.argtypes = args
            fn.restype = None


class Impl:
    """One implementation (oracle port or compiled reference) behind numpy arguments."""

    def __init__(self, lib, prefix, kind, fallback=None):
        self.lib, self.prefix, self.kind = lib, prefix, kind
        self.fallback = fallback
        _declare(lib, prefix)

    def _f(self, name):
        try:
            return getattr(self.lib, self.prefix + name)
        except AttributeError:
            # the compiled reference covers the kernels (src/kernels, src/metrics); the
            # framework-level ghost exchange (src/framework/domain) exists only as the port
            if self.fallback is None:
                raise
            return self.fallback._f(name)

    @staticmethod
    def _p(a):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data

    def set_threads(self, n):
        """Worker threads of the compiled reference (1 = serial program order). The port is
        always serial."""
        fn = getattr(self.lib, self.prefix + "set_threads", None)
        if fn is not None:
            fn.argtypes = [C.c_int]
            fn(int(n))
            return int(n)
        return 1

    def faraday(self, g, em, coeff1, coeff2, stencil=None):
        st = None
        if stencil is not None:
            st = np.ascontiguousarray(stencil, dtype=np.float32)
        self._f("faraday_mink")(C.byref(g), self._p(em), coeff1, coeff2,
                                st.ctypes.data if st is not None else None)

    def ampere(self, g, em, coeff1, coeff2):
        self._f("ampere_mink")(C.byref(g), self._p(em), coeff1, coeff2)

    def currents_ampere(self, g, em, cur, coeff, ppc0):
        self._f("currents_ampere_mink")(C.byref(g), self._p(em), self._p(cur), coeff, ppc0)

    def filter_pass(self, g, cur, buff, fbc):
        self._f("filter_pass")(C.byref(g), self._p(cur), self._p(buff), (C.c_int * 6)(*fbc))

    def push(self, g, order, ctx, prtls: ParticleSet, npart, em):
        s = prtls.struct()
        self._f("push_sr_mink")(C.byref(g), order, C.byref(ctx), C.byref(s), npart, self._p(em))

    def deposit(self, g, order, prtls: ParticleSet, npart, charge, dt, dx, cur):
        s = prtls.struct()
        self._f("deposit_mink")(C.byref(g), order, C.byref(s), npart, charge, dt, dx, self._p(cur))

    def comm_fields(self, g, fld, c0, c1, fbc):
        self._f("comm_fields_self")(C.byref(g), self._p(fld), fld.shape[0], c0, c1,
                                    (C.c_int * 6)(*fbc))

    def sync_currents(self, g, cur, buff, fbc):
        self._f("sync_currents_self")(C.byref(g), self._p(cur), self._p(buff), (C.c_int * 6)(*fbc))


_cache: dict = {}


def oracle() -> Impl:
    if "orc" not in _cache:
        path = os.path.join(HERE, "liborc.so")
        if not os.path.exists(path):
            build(ref=False)
        _cache["orc"] = Impl(C.CDLL(path), "orc_", "port")
    return _cache["orc"]


def reference(order: int) -> Impl | None:
    """The reference's own kernels for one compile-time SHAPE_ORDER, or None if not built."""
    key = f"ref{order}"
    if key not in _cache:
        path = os.path.join(HERE, "_ref", f"libref_o{order}.so")
        _cache[key] = (Impl(C.CDLL(path), "ref_", "reference", fallback=oracle())
                       if os.path.exists(path) else None)
    return _cache[key]
