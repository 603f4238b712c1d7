"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_ref/libref_curv_o<O>.so -- the
reference's own curvilinear-SR / GR kernels and metric classes compiled in place
(oracle/ref_curv_driver.cpp). Used by tests/golden/make_curv_golden.py to produce the committed
golden vectors and by the tests to re-check them wherever the reference tree was available at
build time."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import orc

HERE = os.path.dirname(os.path.abspath(__file__))

METRIC_SPHERICAL, METRIC_QSPHERICAL, METRIC_KERR_SCHILD, METRIC_QKERR_SCHILD, METRIC_KERR_SCHILD_0 = 1, 2, 3, 4, 5
SR_NQ, GR_NQ = 16, 32  # columns of metric_eval()


class Metric(C.Structure):
    _fields_ = [("kind", C.c_int), ("n1", C.c_int), ("n2", C.c_int),
                ("x1min", C.c_float), ("x1max", C.c_float), ("x2min", C.c_float),
                ("x2max", C.c_float), ("r0", C.c_float), ("h", C.c_float), ("a", C.c_float)]

    @staticmethod
    def make(kind, n, extent, r0=0.0, h=0.0, a=0.0):
        m = Metric()
        m.kind, m.n1, m.n2 = kind, n[0], n[1]
        m.x1min, m.x1max, m.x2min, m.x2max = extent
        m.r0, m.h, m.a = r0, h, a
        return m

    def params8(self):
        return [self.x1min, self.x1max, self.x2min, self.x2max, self.r0, self.h, self.a, 0.0]


class PusherGR(C.Structure):
    _fields_ = [("pusher_flags", C.c_int), ("mass", C.c_float), ("charge", C.c_float),
                ("dt", C.c_float), ("omegaB0", C.c_float), ("epsilon", C.c_float),
                ("niter", C.c_int), ("pbc", C.c_int * 6)]


def make_pusher_gr(**kw) -> PusherGR:
    p = PusherGR()
    p.pusher_flags = kw.get("pusher_flags", orc.PUSHER_BORIS)
    p.mass, p.charge = kw.get("mass", 1.0), kw.get("charge", -1.0)
    p.dt, p.omegaB0 = kw["dt"], kw.get("omegaB0", 1.0)
    p.epsilon, p.niter = kw.get("epsilon", 1e-2), kw.get("niter", 10)
    p.pbc = (C.c_int * 6)(*kw.get("pbc", [orc.PBC_ABSORB, orc.PBC_ABSORB, orc.PBC_AXIS,
                                          orc.PBC_AXIS, 0, 0]))
    return p


class RefCurv:
    def __init__(self, lib):
        self.lib = lib
        G, P, M = C.POINTER(orc.Grid), C.POINTER(orc.Prtls), C.POINTER(Metric)
        vp, i32p = C.c_void_p, C.POINTER(C.c_int)
        sig = {
            "refc_metric_eval": [M, C.c_int, vp, vp, vp],
            "refc_push_sr": [M, G, C.POINTER(orc.Pusher), P, C.c_uint32, vp],
            "refc_deposit": [M, G, P, C.c_uint32, C.c_float, C.c_float, vp],
            "refc_fields_sr": [M, C.c_int, G, vp, vp, C.c_float, C.c_float, i32p],
            "refc_filter_sph": [G, vp, vp, i32p],
            "refc_push_gr": [M, G, C.POINTER(PusherGR), P, C.c_uint32, vp, vp],
            "refc_fields_gr": [M, C.c_int, G, vp, vp, vp, C.c_float, i32p],
            "refc_time_average_db": [G, vp, vp],
            "refc_time_average_j": [G, vp, vp],
        }
        for k, a in sig.items():
            f = getattr(lib, k)
            f.argtypes, f.restype = a, None
        lib.refc_metric_dxmin.argtypes = [M]
        lib.refc_metric_dxmin.restype = C.c_float

    @staticmethod
    def _p(a):
        if a is None:
            return None
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data

    def metric_eval(self, m, x1, x2):
        x1 = np.ascontiguousarray(x1, np.float32)
        x2 = np.ascontiguousarray(x2, np.float32)
        out = np.zeros((x1.size, SR_NQ if m.kind <= 2 else GR_NQ), np.float32)
        self.lib.refc_metric_eval(C.byref(m), x1.size, self._p(x1), self._p(x2), self._p(out))
        return out

    def dxmin(self, m):
        return float(self.lib.refc_metric_dxmin(C.byref(m)))

    def push_sr(self, m, g, ctx, prtls, npart, em):
        s = prtls.struct()
        self.lib.refc_push_sr(C.byref(m), C.byref(g), C.byref(ctx), C.byref(s), npart, self._p(em))

    def deposit(self, m, g, prtls, npart, charge, dt, cur):
        s = prtls.struct()
        self.lib.refc_deposit(C.byref(m), C.byref(g), C.byref(s), npart, charge, dt, self._p(cur))

    def fields_sr(self, m, which, g, em, cur, coeff, inv_n0, fbc):
        self.lib.refc_fields_sr(C.byref(m), which, C.byref(g), self._p(em), self._p(cur), coeff,
                                inv_n0, (C.c_int * 6)(*fbc))

    def filter_sph(self, g, cur, buff, fbc):
        self.lib.refc_filter_sph(C.byref(g), self._p(cur), self._p(buff), (C.c_int * 6)(*fbc))

    def push_gr(self, m, g, ctx, prtls, npart, em, em0):
        s = prtls.struct()
        self.lib.refc_push_gr(C.byref(m), C.byref(g), C.byref(ctx), C.byref(s), npart,
                              self._p(em), self._p(em0))

    def fields_gr(self, m, which, g, a, b, c, coeff, fbc):
        self.lib.refc_fields_gr(C.byref(m), which, C.byref(g), self._p(a), self._p(b), self._p(c),
                                coeff, (C.c_int * 6)(*fbc))

    def time_average_db(self, g, em, em0):
        self.lib.refc_time_average_db(C.byref(g), self._p(em), self._p(em0))

    def time_average_j(self, g, cur, cur0):
        self.lib.refc_time_average_j(C.byref(g), self._p(cur), self._p(cur0))


_cache: dict = {}


def reference(order: int) -> RefCurv | None:
    """The compiled reference for one SHAPE_ORDER, or None when it was not built."""
    if order not in _cache:
        path = os.path.join(HERE, "_ref", f"libref_curv_o{order}.so")
        _cache[order] = RefCurv(C.CDLL(path)) if os.path.exists(path) else None
    return _cache[order]
