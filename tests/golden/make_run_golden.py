"""Generates tests/golden/run_*.npz: whole-run golden states from the reference's own
entity.xc (built by oracle/build_entity_xc.sh with the dump-wrapper pgens of oracle/pgens/),
Kokkos-OpenMP with ONE thread (= serial program order: particles of a species deposit in
array order, species in index order -- the order eb200's ORDERED deposit mode reproduces).

    python tests/golden/make_run_golden.py          # needs /root/reference + the built binaries

Per case: the full state after step S0 (import point) and, for every later step up to S1,
the fields em / cur, npart (before and after the pgen's own injection) and per-array
checksums; full particle arrays again at the last step; for runs with an injector the
injected tail [npart_pre, npart) of every step (the reference's Kokkos RNG stream is not
part of the hot path: the test re-imports what the injector appended)."""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refdump  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(ROOT, "baseline", "_ref")

PRTL = ["i1", "i2", "i3", "dx1", "dx2", "dx3", "ux1", "ux2", "ux3", "weight", "i1_prev", "i2_prev",
        "i3_prev", "dx1_prev", "dx2_prev", "dx3_prev", "tag", "phi"]

CASES = {
    # name: (binary, input, first step, last step)
    "stream2d": ("omp/entity_streaming.xc", "stream2d.toml", 0, 10),
    "reconnection_small": ("omp/entity_reconnection.xc", "reconnection_small.toml", 0, 7),
    # antenna-driven (ext_current); the 3D one is the esirkepov / shape_order = 3 build
    "turbulence2d": ("omp/entity_turbulence.xc", "turbulence2d.toml", 0, 8),
    "turbulence3d": ("omp3/entity_turbulence.xc", "turbulence3d.toml", 0, 4),
    # GRPIC: vacuum Wald solution (fields + boundaries only)
    "wald_small": ("omp/entity_wald.xc", "wald_small.toml", 0, 12),
    # GRPIC with particles: pusher, deposit into cur0, AbsorbCurrents, filter, both AmpereCurrents
    "accretion_small": ("omp/entity_accretion.xc", "accretion_small.toml", 0, 12),
    # curvilinear SRPIC: qspherical pulsar magnetosphere with atmosphere injection
    "magnetosphere_small": ("omp/entity_magnetosphere.xc", "magnetosphere_small.toml", 0, 12),
}


def checksum(a: np.ndarray) -> np.uint64:
    """order-sensitive 64-bit checksum of the raw bits (numpy only)"""
    b = np.ascontiguousarray(a).view(np.uint8)
    pad = (-b.size) % 4
    if pad:
        b = np.concatenate([b, np.zeros(pad, np.uint8)])
    w = b.view(np.uint32).astype(np.uint64)
    k = (np.arange(w.size, dtype=np.uint64) * np.uint64(2654435761) + np.uint64(1)) & np.uint64(0xFFFFFFFF)
    return np.uint64((w * k).sum(dtype=np.uint64))


def run_case(name):
    exe, inp, s0, s1 = CASES[name]
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        env = dict(os.environ, OMP_NUM_THREADS="1", EB_DUMP_DIR=tmp,
                   EB_DUMP_STEPS=",".join(str(s) for s in range(s0, s1 + 1)))
        r = subprocess.run([os.path.join(BIN, exe), "-input", os.path.join(HERE, "run_inputs", inp)],
                           cwd=tmp, env=env, capture_output=True, text=True, timeout=1200)
        if r.returncode != 0:
            raise RuntimeError(r.stdout[-2000:] + r.stderr[-2000:])
        prev_n = {}
        for s in range(s0, s1 + 1):
            d = refdump.read(os.path.join(tmp, f"s{s}_d0.bin"))
            nsp = sum(1 for k in d if k.endswith("_npart"))
            out[f"s{s}/time"] = d["time"]
            out[f"s{s}/em"] = d["em"]
            out[f"s{s}/cur"] = d["cur"]
            for extra in ("em0", "cur0", "aux"):
                if extra in d:
                    out[f"s{s}/{extra}"] = d[extra]
            scl = os.path.join(tmp, f"s{s}_scl.bin")
            if s == s0 and os.path.exists(scl):
                for k_, v_ in refdump.read(scl).items():
                    out[f"meta/{k_}"] = v_
            tgt = os.path.join(tmp, f"s{s}_tgt.bin")
            if s == s0 and os.path.exists(tgt):
                for k_, v_ in refdump.read(tgt).items():
                    out[f"meta/target_{k_}"] = v_
            ant = os.path.join(tmp, f"s{s}_ant.bin")
            if os.path.exists(ant):
                for k_, v_ in refdump.read(ant).items():
                    out[f"s{s}/ant_{k_}"] = v_
            for k in range(nsp):
                p = f"sp{k}_"
                n = int(d[p + "npart"][0])
                # what any injector (the engine's ParticleInjector or the pgen's CustomPostStep)
                # appended since the previous dump: nothing compacts the arrays inside the window
                npre = prev_n.get(k, n)
                prev_n[k] = n
                out[f"s{s}/{p}npart"] = np.array([npre, n], np.int64)
                out[f"meta/{p}mass_charge"] = d[p + "mass_charge"]
                for a in PRTL:
                    v = d[p + a]
                    if v.size == 0:
                        continue
                    if s in (s0, s1):
                        out[f"s{s}/{p}{a}"] = v[:n]
                    else:
                        out[f"s{s}/{p}{a}_sum"] = np.array([checksum(v[:npre])], np.uint64)
                        if n > npre:
                            out[f"s{s}/{p}{a}_inj"] = v[npre:n]
    out["meta/steps"] = np.array([s0, s1], np.int64)
    path = os.path.join(HERE, f"run_{name}.npz")
    np.savez_compressed(path, **out)
    print(name, os.path.getsize(path) // 1024, "KiB", len(out), "arrays")


if __name__ == "__main__":
    for nm in (sys.argv[1:] or CASES):
        run_case(nm)
