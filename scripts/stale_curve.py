"""Push+deposit time per step as a function of the steps since the last cell sort (REC bench
workload): prints one line per step. Diagnosis tool for the sort interval."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import entity_b200 as eb
from entity_b200 import workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 45
sim = workloads.reconnection((4096, 2048), ppc0=32, nfilter=8, fused=True, sort_interval=0,
                             deposit_mode=eb.DEPOSIT_AGGREGATED, seed=0x5678)
for sp in sim.species:
    sim.ctx.sort_particles(sp.arrays, sp.npart, remove_dead=False)
sim.step_index = 1
sim.profile(True)
for k in range(n):
    sim.step()
    torch.cuda.synchronize()
    prof = sim.read_profile()
    nc = 0
    for sp in sim.species:
        a = sp.arrays
        nc += int(((a["i1"][:sp.npart] != a["i1_prev"][:sp.npart]) | (a["i2"][:sp.npart] != a["i2_prev"][:sp.npart])).sum())
    print(f"stale {k:3d}  push_deposit {prof['PushDeposit'][0]:7.3f} ms  crossers {nc}", flush=True)
