#!/usr/bin/env python
"""Drop-in demonstration on a GPU box: the reference's own entity.xc (Kokkos-CUDA sm_100 build)
with and without integration/eb200_shim.hpp patched in, same pgen, same input.

 1. parity: pgens/streaming (2D, periodic, four species) for 10 steps, states dumped by the
    CustomPostStep wrapper of both binaries: particle arrays in identical order (same Kokkos
    random pool, same sort), fields / currents within fp32 tolerance (both sides contract FMAs
    differently and sum J with atomics);
 2. timing: pgens/reconnection at 4096 x 2048 x 32 ppc, the reference's own per-step timers, with
    the reference's default particle order (never sorted) and with spatial_sorting_interval = 20.

usage: python integration/run_shim_check.py [out.json]"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import refrun  # noqa: E402
from oracle import refdump  # noqa: E402

BIN = os.path.join(ROOT, "baseline", "_ref")
INP = os.path.join(ROOT, "tests", "golden", "run_inputs", "stream2d.toml")
STEPS = (0, 5, 10)


def dump_run(flavour, extra_env=None):
    tmp = tempfile.mkdtemp(prefix=f"eb_shim_{flavour}_")
    env = dict(os.environ, **(extra_env or {}), EB_DUMP_DIR=tmp, EB_DUMP_STEPS=",".join(str(s) for s in STEPS),
               LD_LIBRARY_PATH=os.path.join(ROOT, "entity_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([os.path.join(BIN, flavour, "entity_streaming.xc"), "-input", INP], cwd=tmp, env=env,
                       capture_output=True, text=True, timeout=900)
    if r.returncode != 0:
        raise RuntimeError(f"{flavour}: rc {r.returncode}\n" + r.stdout[-3000:] + r.stderr[-3000:])
    return {s: refdump.read(os.path.join(tmp, f"s{s}_d0.bin")) for s in STEPS}


def parity(filter_too=False):
    """filter_too: also hand srpic::CurrentsFilter to eb200_filter (EB200_SHIM_FILTER=1)"""
    out = {}
    a, b = dump_run("cuda"), dump_run("cuda_shim", {"EB200_SHIM_FILTER": "1"} if filter_too else None)
    for s in STEPS:
        da, db = a[s], b[s]
        rec = {}
        for k in ("em", "cur"):
            x, y = da[k].astype(np.float64), db[k].astype(np.float64)
            rec[k + "_rel_err"] = float(np.abs(x - y).max() / max(np.abs(x).max(), 1e-30))
        nsp = sum(1 for k in da if k.endswith("_npart"))
        worst, moved, total = 0.0, 0, 0
        for sp in range(nsp):
            na, nb = int(da[f"sp{sp}_npart"][-1]), int(db[f"sp{sp}_npart"][-1])
            rec[f"sp{sp}_npart"] = [na, nb]
            if na != nb or na == 0:
                continue
            same = np.ones(na, bool)
            for nm in ("i1", "i2"):
                same &= da[f"sp{sp}_{nm}"][:na] == db[f"sp{sp}_{nm}"][:na]
            moved += int((~same).sum())
            total += na
            for nm in ("dx1", "dx2", "ux1", "ux2", "ux3"):
                d = np.abs(da[f"sp{sp}_{nm}"][:na] - db[f"sp{sp}_{nm}"][:na])[same]
                if d.size:
                    worst = max(worst, float(d.max()))
        rec["particles_in_another_cell"] = [moved, total]
        rec["max_abs_diff_dx_u"] = worst
        out[f"step{s}"] = rec
    return out


def main():
    out = {"parity": parity(), "parity_with_filter": parity(True), "timing": {}}
    os.environ["LD_LIBRARY_PATH"] = os.path.join(ROOT, "entity_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", "")
    # the reference's default (particles.spatial_sorting_interval = 0: particles stay in injection
    # order) and with its cell sort switched on every 20 steps
    for sort_interval in (0, 20):
        for flavour in ("cuda", "cuda_shim"):
            r = refrun.run(flavour, (4096, 2048), nsteps=45, skip=21, sort_interval=sort_interval)
            out["timing"][f"{flavour}_sort{sort_interval}"] = None if r is None else {k: r[k] for k in r if k != "raw"}
    os.environ["EB200_SHIM_FILTER"] = "1"
    r = refrun.run("cuda_shim", (4096, 2048), nsteps=45, skip=21, sort_interval=20)
    out["timing"]["cuda_shim_sort20_filter"] = None if r is None else {k: r[k] for k in r if k != "raw"}
    del os.environ["EB200_SHIM_FILTER"]
    text = json.dumps(out, indent=1, default=str)
    print(text)
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            f.write(text)


if __name__ == "__main__":
    main()
