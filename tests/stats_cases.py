"""Seeded inputs of the reduced-statistics parity cases (tests/golden/stats_golden.npz): shared
by the generator (the compiled reference, oracle/_ref/libref_stats.so) and the GPU tests."""
import numpy as np

from helpers import random_particles, smooth_fields
from oracle import orc

DX = 0.5
GRIDS = {1: (96,), 2: (48, 40), 3: (14, 12, 10)}
FIELD_STATS = {"B2": 0, "E2": 1, "ExB": 2, "JdotE": 3}
PRTL_STATS = {"Npart": 0, "N": 1, "Rho": 2, "Charge": 3, "T": 4}
T_COMPONENTS = [(0, 0), (0, 1), (0, 3), (1, 1), (2, 3), (3, 3)]
SPECIES = [(1.0, -1.0), (0.0, 0.0), (1836.0, 1.0)]  # (mass, charge); massless = photons


def grid(dim):
    return orc.Grid.make(GRIDS[dim], 2)


def fields(dim):
    g = grid(dim)
    em = smooth_fields(g, 300 + dim, amp=0.8)
    cur = smooth_fields(g, 400 + dim, amp=0.3)[:3].copy()
    return g, em, cur


def particles(dim, k):
    g = grid(dim)
    n = 20000 + 7 * k
    p = random_particles(g, n, 500 + 10 * dim + k, umag=1.5, dead_frac=0.05)
    rng = np.random.default_rng(600 + 10 * dim + k)
    p.weight[:n] = rng.uniform(0.5, 2.0, n).astype(np.float32)
    return g, p, n


def field_cases():
    for dim in (1, 2, 3):
        for name, what in FIELD_STATS.items():
            for comp in ((0,) if name == "JdotE" else (1, 2, 3)):
                yield dim, name, what, comp


def particle_cases():
    for dim in (1, 2, 3):
        for k, (mass, charge) in enumerate(SPECIES):
            for name, what in PRTL_STATS.items():
                if name in ("Rho", "Charge") and mass == 0.0:
                    continue  # the reference raises (reduced_stats.hpp:428-431)
                for use_w in ((False, True) if name in ("N", "Rho", "Charge") else (False,)):
                    for c1, c2 in (T_COMPONENTS if name == "T" else [(0, 0)]):
                        yield dim, k, mass, charge, name, what, use_w, c1, c2


# ---------------------------------------------------------------- 2D curvilinear SRPIC meshes
# kind (ref driver): 1 spherical, 2 qspherical; ext = x1min, x1max, x2min, x2max, r0, h
CURV = {"spherical": (1, [1.0, 20.0, 0.0, float(np.float32(np.pi)), 0.0, 0.0]),
        "qspherical": (2, [1.0, 20.0, 0.0, float(np.float32(np.pi)), 0.2, 0.4])}


def curv_particles(k):
    g, p, n = particles(2, k)
    rng = np.random.default_rng(700 + k)
    p.phi[:n] = rng.uniform(0.0, 2 * np.pi, n).astype(np.float32)
    return g, p, n


def curv_field_cases():
    for name, what in FIELD_STATS.items():
        for comp in ((0,) if name == "JdotE" else (1, 2, 3)):
            yield name, what, comp


def curv_particle_cases():
    for d, k, mass, charge, name, what, use_w, c1, c2 in particle_cases():
        if d == 2:
            yield k, mass, charge, name, what, use_w, c1, c2
