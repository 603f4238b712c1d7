"""Seeded inputs of the curvilinear-SR / GR parity cases (numpy only).

The same generators feed tests/golden/make_curv_golden.py (which runs the reference's own kernels,
oracle/_ref/libref_curv_o*.so, and stores their OUTPUTS under tests/golden/) and the tests (which
regenerate the inputs and compare the CUDA path -- and, where it was built, the compiled reference
again -- with the stored outputs)."""
from __future__ import annotations

import numpy as np

from oracle import orc
from oracle import refcurv as R

N = (20, 12)           # active cells
PI = float(np.float32(np.pi))

# name -> (kind, extent (r_min, r_max, th_min, th_max), r0, h, a)
SR_METRICS = {
    "sph": (R.METRIC_SPHERICAL, (1.0, 10.0, 0.0, PI), 0.0, 0.0, 0.0),
    "qsph": (R.METRIC_QSPHERICAL, (1.0, 30.0, 0.0, PI), 0.0, 0.0, 0.0),
    "qsph_h": (R.METRIC_QSPHERICAL, (1.0, 30.0, 0.0, PI), 0.2, 0.25, 0.0),
}
GR_METRICS = {
    "ks": (R.METRIC_KERR_SCHILD, (1.0, 8.0, 0.0, PI), 0.0, 0.0, 0.95),
    "qks": (R.METRIC_QKERR_SCHILD, (1.0, 20.0, 0.0, PI), 0.0, 0.0, 0.95),
    "qks_h": (R.METRIC_QKERR_SCHILD, (1.0, 20.0, 0.0, PI), 0.1, 0.3, 0.7),
    "ks0": (R.METRIC_KERR_SCHILD_0, (1.0, 8.0, 0.0, PI), 0.0, 0.0, 0.0),
}

# SR pusher variants (kw of orc.make_pusher); x1 faces absorb (what ATMOSPHERE / ABSORB map to,
# context.h:143-146) unless stated, x2 faces are the polar axis
AXIS_PBC = [orc.PBC_ABSORB, orc.PBC_ABSORB, orc.PBC_AXIS, orc.PBC_AXIS, 0, 0]
SR_PUSH = {
    "boris": dict(pusher_flags=2),
    "vay": dict(pusher_flags=4),
    "gca": dict(pusher_flags=2 | 8, gca_larmor_max=0.4, gca_e_ovr_b_sqr_max=0.9),
    "atm": dict(pusher_flags=2, has_atmosphere=1, atm_gx1=-0.4, atm_x_surf=1.0, atm_ds=1.5),
    "atm_gca": dict(pusher_flags=2 | 8, gca_larmor_max=0.4, gca_e_ovr_b_sqr_max=0.9,
                    has_atmosphere=1, atm_gx1=-0.4, atm_x_surf=1.0, atm_ds=1.5),
    "sync": dict(pusher_flags=2, drag_flags=1, sync_coeff=0.02),
    "photon": dict(pusher_flags=1),
    "reflect": dict(pusher_flags=2,
                    pbc=[orc.PBC_REFLECT, orc.PBC_REFLECT, orc.PBC_AXIS, orc.PBC_AXIS, 0, 0]),
}
AXIS_FBC = [orc.FBC_NONE, orc.FBC_NONE, orc.FBC_AXIS, orc.FBC_AXIS, 0, 0]
OPEN_FBC = [orc.FBC_NONE] * 6


def metric(name):
    kind, ext, r0, h, a = {**SR_METRICS, **GR_METRICS}[name]
    return R.Metric.make(kind, N, ext, r0, h, a)


def grid(order):
    return orc.Grid.make(N, orc.nghosts_for(order))


def fields(g, seed, ncomp=6, amp=0.3):
    """smooth + noise, every cell (ghosts included) non-zero"""
    rng = np.random.default_rng(seed)
    shp = g.shape(ncomp)
    j, i = np.meshgrid(np.arange(shp[1]), np.arange(shp[2]), indexing="ij")
    out = np.zeros(shp, np.float64)
    for c in range(ncomp):
        k1, k2 = rng.integers(1, 4, 2)
        out[c] = np.sin(2 * np.pi * k1 * i / shp[2] + rng.uniform(0, 6)) * \
            np.cos(2 * np.pi * k2 * j / shp[1] + rng.uniform(0, 6)) + \
            0.2 * rng.standard_normal(shp[1:])
    return (amp * out).astype(np.float32)


def particles(seed, n=600, umag=1.5, dead_frac=0.05):
    rng = np.random.default_rng(seed)
    p = orc.ParticleSet(n)
    p.i1[:] = rng.integers(0, N[0], n)
    p.i2[:] = rng.integers(0, N[1], n)
    # a band of particles next to every boundary so that each BC branch is taken
    p.i1[:40] = 0
    p.i1[40:80] = N[0] - 1
    p.i2[80:120] = 0
    p.i2[120:160] = N[1] - 1
    for nm in ("dx1", "dx2"):
        d = rng.random(n, dtype=np.float32)
        d[d >= 1.0] = 0.0
        getattr(p, nm)[:] = d
    p.dx2[80:100] *= 0.01          # hugging the axis
    p.i1_prev[:], p.i2_prev[:] = p.i1, p.i2
    p.dx1_prev[:], p.dx2_prev[:] = p.dx1, p.dx2
    p.phi[:] = rng.uniform(0, 2 * np.pi, n).astype(np.float32)
    for nm in ("ux1", "ux2", "ux3"):
        getattr(p, nm)[:] = (umag * rng.standard_normal(n)).astype(np.float32)
    p.weight[:] = rng.uniform(0.5, 1.5, n).astype(np.float32)
    p.tag[:] = 1
    p.tag[rng.random(n) < dead_frac] = 0
    return p


def gr_particles(seed, scale, n=600):
    """`scale(x1, x2) -> (sqrt(h_11), sqrt(h_22), sqrt(h_33))`: covariant momenta of order one in
    the local tetrad"""
    p = particles(seed, n=n, umag=1.0)
    s = scale(p.i1 + p.dx1, p.i2 + p.dx2)
    p.ux1[:] = (p.ux1 * s[0]).astype(np.float32)
    p.ux2[:] = (p.ux2 * s[1]).astype(np.float32)
    p.ux3[:] = (p.ux3 * np.minimum(s[2], 1e3)).astype(np.float32)
    return p


PRTL_OUT = ("i1", "i2", "dx1", "dx2", "ux1", "ux2", "ux3", "i1_prev", "i2_prev", "dx1_prev",
            "dx2_prev", "phi", "tag")


def prtl_state(p):
    return {nm: getattr(p, nm).copy() for nm in PRTL_OUT}


# ------------------------------------------------------------------------------ case lists
def sr_push_cases():
    out = []
    for v in SR_PUSH:
        out.append(("qsph", v, 0))
    for mname in ("sph", "qsph_h"):
        for v in ("boris", "atm_gca", "reflect"):
            out.append((mname, v, 0))
    for mname in SR_METRICS:
        out.append((mname, "boris", 2))
    out.append(("qsph", "gca", 1))
    out.append(("qsph", "vay", 3))
    return out


def gr_push_cases():
    out = []
    for mname in GR_METRICS:
        out.append((mname, "massive", 0))
        out.append((mname, "photon", 0))
    out.append(("qks", "massive", 1))
    out.append(("ks", "massive", 2))
    out.append(("qks_h", "massive", 3))
    return out


def _dt(ref, m):
    return float(np.float32(0.4 * ref.dxmin(m)))


def _sr_pusher(variant, dt):
    kw = dict(SR_PUSH[variant])
    kw.setdefault("pbc", AXIS_PBC)
    return dict(dt=dt, omegaB0=0.8, mass=1.0, charge=-1.0, **kw)


class Backend:
    """What a case needs from an implementation; see RefBackend / DeviceBackend."""


class RefBackend(Backend):
    """the compiled reference (oracle/_ref/libref_curv_o<O>.so)"""

    def __init__(self):
        self.refs = {o: R.reference(o) for o in range(4)}
        if any(v is None for v in self.refs.values()):
            raise RuntimeError("oracle/_ref/libref_curv_o*.so not built")

    def dxmin(self, m):
        return self.refs[0].dxmin(m)

    def push_sr(self, m, order, kw, p, em):
        self.refs[order].push_sr(m, grid(order), orc.make_pusher(**kw), p, p.n, em)

    def deposit(self, m, order, p, charge, dt, cur):
        self.refs[order].deposit(m, grid(order), p, p.n, charge, dt, cur)

    def fields_sr(self, m, which, em, cur, coeff, inv_n0, fbc):
        self.refs[0].fields_sr(m, which, grid(0), em, cur, coeff, inv_n0, fbc)

    def filter_sph(self, m, cur, buff, fbc):
        self.refs[0].filter_sph(grid(0), cur, buff, fbc)

    def push_gr(self, m, order, kw, p, em, em0):
        self.refs[order].push_gr(m, grid(order), R.make_pusher_gr(**kw), p, p.n, em, em0)

    def fields_gr(self, m, which, a, b, c, coeff, fbc):
        self.refs[0].fields_gr(m, which, grid(0), a, b, c, coeff, fbc)

    def time_average(self, m, a, b):
        # a = (a + b) / 2 on the active cells
        if a.shape[0] == 6:
            self.refs[0].time_average_db(grid(0), b, a)
        else:
            self.refs[0].time_average_j(grid(0), a, b)

    def metric_eval(self, m, x1, x2):
        return self.refs[0].metric_eval(m, x1, x2)


def _gr_scale(be, m):
    def scale(x1, x2):
        q = be.metric_eval(m, x1.astype(np.float32), x2.astype(np.float32))
        return np.sqrt(np.abs(q[:, 0])), np.sqrt(np.abs(q[:, 1])), np.sqrt(np.abs(q[:, 2]))
    return scale


def run_all(be: Backend, ref_for_setup=None):
    """Runs every case through `be`; returns {key: ndarray}. `ref_for_setup` supplies dxMin and
    the momentum scaling when `be` cannot (the device backend takes them from the golden file)."""
    setup = ref_for_setup or be
    out = {}
    # ---- metrics
    rng = np.random.default_rng(7)
    for mname in list(SR_METRICS) + list(GR_METRICS):
        m = metric(mname)
        x1 = rng.uniform(-1, N[0] + 1, 200).astype(np.float32)
        x2 = rng.uniform(0.02, N[1] - 0.02, 200).astype(np.float32)
        out[f"metric/{mname}"] = be.metric_eval(m, x1, x2)
    # ---- SR pusher + deposit
    for k, (mname, variant, order) in enumerate(sr_push_cases()):
        m, g = metric(mname), grid(order)
        dt = setup.dt(mname)
        em = fields(g, 100 + k)
        p = particles(200 + k)
        be.push_sr(m, order, _sr_pusher(variant, dt), p, em)
        key = f"sr_push/{mname}/{variant}/o{order}"
        for nm, a in prtl_state(p).items():
            out[f"{key}/{nm}"] = a
        cur = np.zeros(g.shape(3), np.float32)
        p.tag[(p.i1 < 0) | (p.i1 >= N[0]) | (p.i2 < 0) | (p.i2 >= N[1])] = 0
        be.deposit(m, order, p, -1.0, dt, cur)
        out[f"{key}/J"] = cur
    # ---- SR field solvers + filter
    g = grid(0)
    for mname in SR_METRICS:
        m = metric(mname)
        for bname, fbc in (("axis", AXIS_FBC), ("open", OPEN_FBC)):
            em, cur = fields(g, 300), fields(g, 301, ncomp=3)
            be.fields_sr(m, 0, em, cur, 0.013, 0.0, fbc)
            out[f"sr_fld/{mname}/{bname}/faraday"] = em.copy()
            be.fields_sr(m, 1, em, cur, 0.013, 0.0, fbc)
            out[f"sr_fld/{mname}/{bname}/ampere"] = em.copy()
            be.fields_sr(m, 2, em, cur, -0.021, 0.5, fbc)
            out[f"sr_fld/{mname}/{bname}/cur_ampere_E"] = em.copy()
            out[f"sr_fld/{mname}/{bname}/cur_ampere_J"] = cur.copy()
    for bname, fbc in (("axis", AXIS_FBC), ("open", OPEN_FBC)):
        cur = fields(g, 310, ncomp=3)
        for _ in range(2):
            buff = cur.copy()
            be.filter_sph(metric("qsph"), cur, buff, fbc)
        out[f"sr_fld/filter/{bname}"] = cur.copy()
    # ---- GR pusher + deposit
    for k, (mname, kind, order) in enumerate(gr_push_cases()):
        m, g = metric(mname), grid(order)
        dt = setup.dt(mname)
        em, em0 = fields(g, 400 + k, amp=0.05), fields(g, 450 + k, amp=0.05)
        p = gr_particles(500 + k, setup.gr_scale(mname))
        kw = dict(pusher_flags=1 if kind == "photon" else 2, dt=dt, omegaB0=0.8, niter=10)
        be.push_gr(m, order, kw, p, em, em0)
        key = f"gr_push/{mname}/{kind}/o{order}"
        for nm, a in prtl_state(p).items():
            out[f"{key}/{nm}"] = a
        cur = np.zeros(g.shape(3), np.float32)
        p.tag[(p.i1 < 0) | (p.i1 >= N[0]) | (p.i2 < 0) | (p.i2 >= N[1])] = 0
        be.deposit(m, order, p, -1.0, dt, cur)
        out[f"{key}/J"] = cur
    # ---- GR field kernels
    g = grid(0)
    for mname in GR_METRICS:
        m = metric(mname)
        for bname, fbc in (("axis", AXIS_FBC), ("open", OPEN_FBC)):
            key = f"gr_fld/{mname}/{bname}"
            d, b = fields(g, 600), fields(g, 601)
            e, h = fields(g, 602), fields(g, 603)
            be.fields_gr(m, 0, d, b, e, 0.0, fbc)
            out[f"{key}/aux_e"] = e.copy()
            be.fields_gr(m, 1, d, b, h, 0.0, fbc)
            out[f"{key}/aux_h"] = h.copy()
            b_out = fields(g, 604)
            be.fields_gr(m, 2, b, b_out, e, 0.017, fbc)       # out of place
            out[f"{key}/faraday_oop"] = b_out.copy()
            be.fields_gr(m, 2, b, b, e, 0.017, fbc)           # in place
            out[f"{key}/faraday_inp"] = b.copy()
            d_out = fields(g, 605)
            be.fields_gr(m, 3, d, d_out, h, 0.017, fbc)
            out[f"{key}/ampere_oop"] = d_out.copy()
            be.fields_gr(m, 3, d, d, h, 0.017, fbc)
            out[f"{key}/ampere_inp"] = d.copy()
            j = fields(g, 606, ncomp=3)
            be.fields_gr(m, 4, d, j, None, -0.03, fbc)
            out[f"{key}/cur_ampere"] = d.copy()
    a6, b6 = fields(g, 610), fields(g, 611)
    be.time_average(metric("ks"), a6, b6)
    out["gr_fld/time_average_db"] = a6
    a3, b3 = fields(g, 612, ncomp=3), fields(g, 613, ncomp=3)
    be.time_average(metric("ks"), a3, b3)
    out["gr_fld/time_average_j"] = a3
    return out


class RefSetup:
    """dxMin-derived time steps and momentum scales, from the compiled reference"""

    def __init__(self, be: RefBackend):
        self.be = be

    def dt(self, mname):
        return _dt(self.be.refs[0], metric(mname))

    def gr_scale(self, mname):
        return _gr_scale(self.be, metric(mname))
