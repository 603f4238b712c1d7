import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import run_cases as rc
import entity_b200 as eb
from entity_b200 import lib as L
from entity_b200.grpic import GRSimulation
import test_gpu_grpic as T
mods = (torch, eb, L, GRSimulation)
case = sys.argv[1] if len(sys.argv) > 1 else "wald_small"
z = rc.load(case)
s0, s1 = (int(v) for v in z["meta/steps"])
sim = T.build(mods, case, z, s0)
for s in range(s0 + 1, s0 + 3):
    sim.step()
    for nm in ("em", "em0", "aux"):
        a, b = getattr(sim, nm).cpu().numpy(), z[f"s{s}/{nm}"]
        for c in range(6):
            d = np.abs(a[c] - b[c])
            d[~np.isfinite(d)] = 0
            j, i = np.unravel_index(np.argmax(d), d.shape)
            print(f"step {s} {nm}[{c}] maxerr {d.max():.3e} at (i1={i}, i2={j}) ref {b[c][j,i]:.6e} got {a[c][j,i]:.6e}  max|ref| {np.nanmax(np.abs(b[c])):.3e}; rows with err>1e-5*max: i1 in {sorted(set(np.nonzero(d > 1e-5*np.nanmax(np.abs(b[c])))[1].tolist()))[:12]} i2 in {sorted(set(np.nonzero(d > 1e-5*np.nanmax(np.abs(b[c])))[0].tolist()))[:12]}")
    if case == "accretion_small":
        T.import_injected(mods, sim, z, s, s1)
