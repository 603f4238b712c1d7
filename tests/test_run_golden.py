"""Pins the oracle's WHOLE STEP (oracle/pic.py: order of the substeps, coefficients, skip rules,
boundary calls) to the running reference: states dumped from the reference's own entity.xc
(Kokkos-OpenMP, one thread; tests/golden/make_run_golden.py) over a window of steps.

* stream2d            -- BASELINE configs[0] lifted to 2D, doubly periodic: every step of the window
                         bit for bit (fields, currents, every particle array).
* turbulence2d / 3d   -- pgens/turbulence (BASELINE configs[2]: the 3D case is the esirkepov,
                         shape_order = 3 build) with the antenna's ext_current: bit for bit, the
                         antenna amplitudes of every step imported from the dump (the pgen advances
                         them on the host with its own RNG).
* reconnection_small  -- BASELINE configs[1] at fixture size with its MATCH / ABSORB x2 walls and the
                         replenishing injector: also bit for bit (the MATCH profiles go through
                         glibc's tanhf, oracle/bcs.py, as the reference's host build does). The
                         particles the pgen's injector appended after each step are imported from
                         the dump (its Kokkos RNG stream is not part of the hot path)."""
import numpy as np
import pytest

from oracle import orc, pic

import run_cases as rc


def build_oracle(case, z, s0):
    c = rc.CASES[case]
    sc = rc.scales(case).derive()
    walls = c["walls"]
    fbc = [orc.FBC_PERIODIC] * 6
    pbc = [orc.PBC_PERIODIC] * 6
    if walls:
        fbc[2] = fbc[3] = orc.FBC_NONE
        pbc[2] = pbc[3] = orc.PBC_ABSORB
    o = pic.OracleSim(orc.oracle(), c["n"], c.get("order", 0), sc, c["dx"], c["nfilter"], fbc=fbc,
                      pbc=pbc, xmin=(tuple(c["xmin"]) + (0.0,))[:3])
    o.em[...] = z[f"s{s0}/em"]
    o.cur[...] = z[f"s{s0}/cur"]
    for k, pusher in enumerate(c["pushers"]):
        n = int(z[f"s{s0}/sp{k}_npart"][1])
        ps = orc.ParticleSet(c["cap"])
        for a in rc.PRTL:
            key = f"s{s0}/sp{k}_{a}"
            if key in z.files:
                getattr(ps, a)[:n] = z[key]
        m, q = z[f"meta/sp{k}_mass_charge"]
        o.add_species(float(m), float(q), ps, n, pusher)
    o.step_index = s0 + 1
    o.time = float(z[f"s{s0}/time"][0]) + float(np.float32(sc["dt"]))
    if walls:
        o.match = (rc.match_faces(case, o.grid), rc.match_target(case, o.grid), 63)
    return o


def check_step(case, z, o, s, s1, exact):
    c = rc.CASES[case]
    em, cur = z[f"s{s}/em"], z[f"s{s}/cur"]
    if exact:
        assert np.array_equal(o.em.view(np.uint32), em.view(np.uint32)), f"step {s}: E/B differ"
        assert np.array_equal(o.cur.view(np.uint32), cur.view(np.uint32)), f"step {s}: J differs"
    else:
        for nm, a, b in (("E/B", o.em, em), ("J", o.cur, cur)):
            tol = 3e-6 * np.abs(b).max()
            assert np.abs(a - b).max() <= tol, f"step {s}: {nm} off by {np.abs(a - b).max():.3e} > {tol:.3e}"
    for k, sp in enumerate(o.species):
        npre, n = (int(v) for v in z[f"s{s}/sp{k}_npart"])
        assert sp["npart"] == npre, f"step {s}: species {k} npart {sp['npart']} != {npre}"
        if c["pushers"][k] == 0:
            continue
        for a in rc.PRTL:
            v = getattr(sp["prtls"], a)
            if f"s{s}/sp{k}_{a}" in z.files:  # full arrays (last step)
                assert np.array_equal(v[:npre], z[f"s{s}/sp{k}_{a}"][:npre]), f"step {s}: sp{k}.{a}"
            elif f"s{s}/sp{k}_{a}_sum" in z.files:
                assert rc.checksum(v[:npre]) == z[f"s{s}/sp{k}_{a}_sum"][0], f"step {s}: sp{k}.{a} checksum"
        # what the pgen's injector appended after this step
        if n > npre:
            for a in rc.PRTL:
                key = f"s{s}/sp{k}_{a}_inj" if s != s1 else f"s{s}/sp{k}_{a}"
                if key in z.files:
                    src = z[key] if s != s1 else z[key][npre:n]
                    getattr(sp["prtls"], a)[npre:n] = src
            sp["npart"] = n


@pytest.mark.parametrize("case,exact", [("stream2d", True), ("reconnection_small", True),
                                        ("turbulence2d", True), ("turbulence3d", True)])
def test_oracle_step_matches_running_reference(orc_mod, case, exact):
    z = rc.load(case)
    s0, s1 = (int(v) for v in z["meta/steps"])
    o = build_oracle(case, z, s0)
    for s in range(s0 + 1, s1 + 1):
        if rc.CASES[case].get("antenna"):
            o.ext = rc.antenna_table(case, z, s - 1)
        o.step()
        check_step(case, z, o, s, s1, exact)
