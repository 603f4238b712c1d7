#!/usr/bin/env python
"""A/B of the ATOMIC and AGGREGATED deposit modes per (dimension, shape order) on cell-sorted
particles (16 per cell): which one AGGREGATED should select. Run under gpurun."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import entity_b200 as eb  # noqa: E402
from entity_b200 import workloads  # noqa: E402

for dim, order, n in ((2, 1, (1024, 512)), (2, 2, (1024, 512)), (2, 3, (1024, 512)),
                      (3, 1, (96, 96, 96)), (3, 2, (96, 96, 96)), (3, 3, (96, 96, 96))):
    res = {}
    for mode, name in ((eb.DEPOSIT_ATOMIC, "atomic"), (eb.DEPOSIT_AGGREGATED, "aggregated")):
        if dim == 2:
            sim = workloads.two_stream(n, ppc0=64, nfilter=0, fused=True, sort_interval=4, deposit_mode=mode)
            from entity_b200.srpic import Scales, Simulation
            s2 = Simulation(n, order, Scales(2, sim.ctx.dx, 100.0, 10.0, 64), nfilter=0, fused=True,
                            sort_interval=4, deposit_mode=mode)
            for sp in sim.species:
                s2.add_species(sp.mass, sp.charge, sp.arrays, sp.npart, sp.pusher)
            sim = s2
        else:
            sim = workloads.turbulence(n, ppc0=16, order=order, nfilter=0, fused=True, sort_interval=4,
                                       deposit_mode=mode)
        if mode == eb.DEPOSIT_AGGREGATED:
            sim.ctx.set_pd_kernel(2)  # force the aggregated (TMA stream) kernel, not the auto choice
        sim.step(5)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sim.step(8)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 8
        res[name] = dt * 1e3
        npart = sim.n_pushed()
        sim.ctx.close()
    print(f"D={dim} O={order} n={npart}: atomic {res['atomic']:.2f} ms  aggregated {res['aggregated']:.2f} ms", flush=True)
