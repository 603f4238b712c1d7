"""Whole-run golden cases (tests/golden/run_*.npz, made by tests/golden/make_run_golden.py from
the reference's own entity.xc): builders that import a dumped state into the oracle stepper
and into an entity_b200 Simulation. Shared by the CPU pin (test_run_golden.py) and the GPU
parity test (test_gpu_run_golden.py)."""
from __future__ import annotations

import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

PRTL = ["i1", "i2", "i3", "dx1", "dx2", "dx3", "ux1", "ux2", "ux3", "weight", "i1_prev", "i2_prev",
        "i3_prev", "dx1_prev", "dx2_prev", "dx3_prev", "tag"]

# name -> parameters of the input file (tests/golden/run_inputs/<name>.toml)
CASES = {
    "stream2d": dict(n=(48, 32), dx=0.1, xmin=(0.0, 0.0), larmor0=100.0, skindepth0=10.0, ppc0=16.0,
                     nfilter=4, pushers=[2, 0, 2, 0], cap=8192, walls=None),
    "reconnection_small": dict(n=(64, 48), dx=0.5, xmin=(-16.0, -12.0), larmor0=0.1, skindepth0=1.0,
                               ppc0=8.0, nfilter=8, pushers=[2, 2], cap=40000,
                               walls=dict(ds=2.0, bg_B=1.0, cs_width=1.5, cs_y=0.0)),
    "turbulence2d": dict(n=(32, 24), dx=0.25, xmin=(-4.0, -3.0), larmor0=1.0, skindepth0=1.0, ppc0=8.0,
                         nfilter=4, pushers=[2, 2], cap=4096, walls=None, antenna=True),
    "turbulence3d": dict(n=(16, 12, 10), dx=0.5, xmin=(-4.0, -3.0, -2.5), larmor0=1.0, skindepth0=1.0,
                         ppc0=8.0, nfilter=4, pushers=[2, 2], cap=8192, walls=None, antenna=True, order=3),
}


def load(name):
    return np.load(os.path.join(HERE, "golden", f"run_{name}.npz"))


def checksum(a: np.ndarray) -> np.uint64:
    b = np.ascontiguousarray(a).view(np.uint8)
    pad = (-b.size) % 4
    if pad:
        b = np.concatenate([b, np.zeros(pad, np.uint8)])
    w = b.view(np.uint32).astype(np.uint64)
    k = (np.arange(w.size, dtype=np.uint64) * np.uint64(2654435761) + np.uint64(1)) & np.uint64(0xFFFFFFFF)
    return np.uint64((w * k).sum(dtype=np.uint64))


def scales(case):
    from entity_b200.srpic import Scales
    c = CASES[case]
    return Scales(len(c["n"]), c["dx"], larmor0=c["larmor0"], skindepth0=c["skindepth0"], ppc0=c["ppc0"])


def antenna_table(case, z, s):
    """eb200_ext_current_t contents from the antenna state dumped after step s (what step s + 1
    adds in CurrentsAmpere)"""
    from oracle import antenna
    dim = len(CASES[case]["n"])
    return antenna.mode_table(dim, z[f"s{s}/ant_k"], z[f"s{s}/ant_a_real"], z[f"s{s}/ant_a_imag"],
                              z[f"s{s}/ant_a_real_inv"], z[f"s{s}/ant_a_imag_inv"])


def match_target(case, g):
    """BoundaryFieldsInX2 of pgens/reconnection/pgen.hpp:104-135 on every component's own node:
    bx1 = bg_B tanh((x2 - cs_y) / cs_width), everything else 0 (bg_Bguide = 0), fp32."""
    c = CASES[case]
    w = c["walls"]
    f32 = np.float32
    tgt = np.zeros(g.shape(6), f32)
    jj = np.arange(g.n[1] + 2 * g.ng, dtype=f32) - f32(g.ng)
    # bx1 lives on (i, j + 1/2)
    y = ((jj + f32(0.5)) * f32(c["dx"]) + f32(c["xmin"][1])).astype(f32)
    from oracle import bcs
    tgt[3] = (f32(w["bg_B"]) * bcs.tanhf(((y - f32(w["cs_y"])) / f32(w["cs_width"])).astype(f32))).astype(f32)[:, None]
    return tgt


def match_faces(case, g):
    """the two x2 layers as MatchFieldsIn hands them to the kernel (fields_bcs.h:72-114)"""
    c = CASES[case]
    f32 = np.float32
    ds, dx, n = c["walls"]["ds"], c["dx"], c["n"]
    ymin = float(f32(c["xmin"][1]))
    ymax = float(f32(c["xmin"][1]) + f32(dx) * f32(n[1]))
    nds = int(round(ds / dx))
    ext = [n[0] + 2 * g.ng, n[1] + 2 * g.ng]
    return [(1, ymin, ds, [0, 0], [ext[0], g.ng + nds]),
            (1, ymax, ds, [0, g.ng + n[1] - nds], [ext[0], ext[1]])]


# ------------------------------------------------------------------ GRPIC cases
GR_CASES = {
    # pgens/wald/wald.toml at fixture size (tests/golden/run_inputs/wald_small.toml)
    "wald_small": dict(n=(64, 48), metric="qkerr_schild", extent=(1.0, 10.0), r0=0.0, h=0.0, a=0.95,
                       larmor0=0.0025, skindepth0=0.05, ppc0=2.0, nfilter=0, deposit=False,
                       match_ds=1.0, pushers=[], cap=0),
    # pgens/accretion/accretion.toml at fixture size (run_inputs/accretion_small.toml)
    "accretion_small": dict(n=(48, 32), metric="qkerr_schild", extent=(1.0, 6.0), r0=0.0, h=0.0, a=0.95,
                            larmor0=0.025, skindepth0=0.5, ppc0=2.0, nfilter=4, deposit=True,
                            match_ds=1.0, pushers=[2, 2], cap=16384, niter=10, eps=1e-2),
}


def gr_match_range(case, ng):
    """grpic::MatchFieldsIn for the +x1 face (fields_bcs.h:42-96): box [x1max - ds, x1max] through
    Mesh::ExtentToRange with incl_ghosts = (false, true) along x1 and (true, true) along x2
    (mesh.h:133-200); qkerr_schild: x1 = (ln(r - r0) - chi_min) / dchi in fp32."""
    c = GR_CASES[case]
    f32 = np.float32
    n1, n2 = c["n"]
    r_min, r_max = f32(c["extent"][0]), f32(c["extent"][1])
    chi_min = np.log(r_min - f32(c["r0"]))
    dchi = (np.log(r_max - f32(c["r0"])) - chi_min) / f32(n1)
    to_cd = lambda r: (np.log(f32(r) - f32(c["r0"])) - chi_min) / dchi
    lo = max(float(np.floor(to_cd(r_max - f32(c["match_ds"])))), 0.0)
    hi = float(np.ceil(to_cd(r_max)))
    return [int(lo) + ng, 0], [int(hi) + 2 * ng, n2 + 2 * ng]


# ------------------------------------------------------------------ curvilinear SRPIC cases
SPH_CASES = {
    # pgens/magnetosphere/magnetosphere.toml at fixture size (run_inputs/magnetosphere_small.toml)
    "magnetosphere_small": dict(n=(64, 48), metric="qspherical", extent=(1.0, 10.0), r0=0.0, h=0.0,
                                nfilter=4, match_ds=1.0, pushers=[2 | 8, 2 | 8], cap=4096),
}


def sph_geometry(case, ng):
    """The ranges srpic::MatchFieldsIn (+x1) and srpic::AtmosphereFieldsIn (-x1) hand to their
    kernels (fields_bcs.h:39-215, 470-560) and the atmosphere's extent (utils.h:48-104) for a
    qspherical mesh (x1 = (ln(r - r0) - chi_min) / dchi), in fp32."""
    c = SPH_CASES[case]
    f32 = np.float32
    n1, n2 = c["n"]
    r0 = f32(c["r0"])
    r_min, r_max = f32(c["extent"][0]), f32(c["extent"][1])
    chi_min = np.log(r_min - r0)
    dchi = (np.log(r_max - r0) - chi_min) / f32(n1)
    to_cd = lambda r: (np.log(f32(r) - r0) - chi_min) / dchi
    to_ph = lambda x: r0 + np.exp(f32(x) * dchi + chi_min)
    # MATCH: box [r_max - ds, r_max], incl_ghosts (false, true) along x1
    lo = max(float(np.floor(to_cd(r_max - f32(c["match_ds"])))), 0.0)
    hi = float(np.ceil(to_cd(r_max)))
    match = dict(range_min=[int(lo) + ng, 0], range_max=[int(hi) + 2 * ng, n2 + 2 * ng],
                 xg_edge=float(r_max))
    # ATMOSPHERE at -x1: buffer of max(nfilter + 2, 5) cells from the inner edge
    buf = max(c["nfilter"] + 2, 5)
    xg_min, xg_max = to_ph(0.0), to_ph(float(buf))
    a_lo = float(np.floor(to_cd(max(xg_min, r_min))))
    a_hi = min(float(np.ceil(to_cd(xg_max))), float(n1))
    rmin, rmax = [int(a_lo) + 0, 0], [int(a_hi) + ng, n2 + 2 * ng]
    atm = dict(range_min=rmin, range_max=rmax, i_edge=rmax[0] - 1, x_surf=float(xg_max))
    return match, atm
