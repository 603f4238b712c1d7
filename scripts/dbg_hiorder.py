import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import entity_b200 as eb
import hiorder_cases as hc
from hiorder_truth import deposit_1d
from helpers import to_device, to_host
from oracle import orc
z = np.load(os.path.join(ROOT, "tests/golden/hiorder_golden.npz"))
# truth, 1D: one deposit of the golden final state
for order in hc.ORDERS:
    g, octx, em, p, n = hc.setup(1, order)
    key = f"1d/o{order}/"
    for nm in hc.names(1):
        getattr(p, nm)[:] = z[key + nm]
    truth = deposit_1d(order, g.ng, g.n[0], p.i1, p.dx1, p.i1_prev, p.dx1_prev, (p.ux1, p.ux2, p.ux3), p.weight, p.tag, -1.0, float(octx.dt), hc.DX)
    ctx = eb.Context(hc.DIMS[1], order=order, strict=True, dx=hc.DX)
    d_j = torch.zeros(g.shape(3), dtype=torch.float32, device="cuda")
    ctx.deposit(to_device(p), n, -1.0, octx.dt, d_j, mode=eb.DEPOSIT_ATOMIC)
    ours = np.abs(d_j.cpu().numpy() - truth).max() / np.abs(truth).max()
    ref = orc.reference(order)
    line = f"truth 1D O {order}: ours {ours:.2e}"
    if ref is not None:
        jr = np.zeros(g.shape(3), np.float32)
        ref.deposit(g, order, p, n, -1.0, octx.dt, hc.DX, jr)
        line += f" reference {np.abs(jr - truth).max() / np.abs(truth).max():.2e}"
    print(line)
