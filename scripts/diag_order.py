"""Diagnostic: how cell-ordered are the particles of the bench workload after k steps?"""
import sys
import torch
sys.path.insert(0, ".")
from entity_b200 import workloads
import entity_b200 as eb

sim = workloads.reconnection((2048, 1024), ppc0=32, nfilter=8, fused=True, sort_interval=20,
                             deposit_mode=eb.DEPOSIT_AGGREGATED)
for step in range(6):
    sp = sim.species[0]
    n = sp.npart
    key = sp.arrays["i1"][:n].long() + 2048 * sp.arrays["i2"][:n].long()
    changes = (key[1:] != key[:-1]).float().mean().item()
    kprev = sp.arrays["i1_prev"][:n].long() + 2048 * sp.arrays["i2_prev"][:n].long()
    cross = (key != kprev).float().mean().item()
    k4 = key[: n // 4 * 4].view(-1, 4)
    same4 = (k4 == k4[:, :1]).all(dim=1).float().mean().item()
    print(f"step {step}: npart {n} key-change frac {changes:.4f} (avg run {1/max(changes,1e-9):.1f}) "
          f"crossed-last-step {cross:.4f} quad-uniform {same4:.4f}")
    sim.step()
