"""TEST INFRASTRUCTURE ONLY -- numpy fp32 restatement of the reference's emission policies.

arch::emission::Synchrotron<M>::shouldEmit (src/archetypes/emission/synchrotron.h:139-224) and
arch::emission::Compton<M>::shouldEmit (src/archetypes/emission/compton.h:133-164), and what
kernel::sr::Pusher_kernel::processEmission does with their answer (src/kernels/pushers/sr.hpp:
1501-1555): recoil of the emitter, direction and momentum of the photon.

PARITY UNPINNED against a running reference: the policies draw from a Kokkos random pool and
append through an atomic counter, neither of which the serial shim provides, and no reference
test exercises them. What pins this file is the formulae's own limits (tests/test_gpu_emission.py:
zero field -> zero synchrotron probability, Compton probability = nominal * beta, recoil
antiparallel to the photon, energy gamma^2 * nominal) and the term-by-term citation below."""
from __future__ import annotations

import numpy as np

f32 = np.float32
SYNCHROTRON, COMPTON = 1, 2


def _cross(a, b):
    return np.stack([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])


def response(kind, u, e, b, photon_weight, nominal_probability, nominal_photon_energy, mass):
    """u, e, b: [3, n] fp32 (u = the mid-step four-velocity). Returns probability, delta_u [3, n],
    photon energy, gamma -- every operation in fp32, in the reference's order."""
    u, e, b = (np.asarray(x, f32) for x in (u, e, b))
    pw, npb, npe, m = f32(photon_weight), f32(nominal_probability), f32(nominal_photon_energy), f32(mass)
    u_sqr = u[0] * u[0] + u[1] * u[1] + u[2] * u[2]
    gamma_sqr = f32(1) + u_sqr
    energy = gamma_sqr * npe
    gamma = np.sqrt(gamma_sqr)
    if kind == COMPTON:
        du = -pw * energy / (np.sqrt(u_sqr) * m)
        return npb * np.sqrt(u_sqr / gamma_sqr), du * u, energy, gamma
    u_mag = np.sqrt(u_sqr)
    beta = u_mag / gamma
    epb = e + _cross(u, b) / gamma
    bde = (u[0] * e[0] + u[1] * e[1] + u[2] * e[2]) / gamma
    kap = _cross(epb, b) + bde * e
    chi = (epb[0] * epb[0] + epb[1] * epb[1] + epb[2] * epb[2]) - bde * bde
    prob = npb * (-(kap[0] * u[0] + kap[1] * u[1] + kap[2] * u[2]) / (gamma_sqr * u_mag) + beta * chi)
    d = -kap + gamma * u * chi
    du = -pw * energy / (np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) * m)
    return prob, du * d, energy, gamma


def decide(prob, energy, gamma, draw, mass, energy_min, should_drag):
    """(emit photon, apply recoil) per particle"""
    should = (draw < prob) & (energy < f32(mass) * (gamma - f32(1)) * f32(0.2))
    return should & (energy >= f32(energy_min)), should & bool(should_drag)
