"""Multi-domain parity worker: one process per GPU (torchrun) or a single process (world 1).

Every rank builds the same seeded GLOBAL problem on the CPU, keeps its own block on its GPU,
runs the CUDA library's exchange (eb200_comm_fields / eb200_sync_currents /
eb200_comm_particles / eb200_srpic_step with a communicator attached) and compares with
 * the single-domain oracle on the global problem (ghost fill, whole step), and
 * the numpy restatement of the exchange protocol, oracle/mdcomm.py (sync, migration).
Exits non-zero on the first mismatch.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/mgpu_worker.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main(decomp2d=None, decomp3d=None):
    import torch
    import torch.distributed as dist
    import entity_b200 as eb
    from entity_b200 import lib as L
    from entity_b200.metadomain import Metadomain, bootstrap_unique_id
    from entity_b200.srpic import Scales, Simulation
    from oracle import mdcomm, orc, pic
    from helpers import random_particles, to_device, to_host
    import test_metadomain as T

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # an ncclUniqueId names ONE communicator: every eb200_comm_init gets a fresh one
    fresh_uid = bootstrap_unique_id
    orc.build(ref=False)
    o = orc.oracle()

    def check(ok, what):
        if not ok:
            print(f"[rank {rank}] FAIL: {what}", flush=True)
            sys.exit(3)

    def dev(a):
        return torch.from_numpy(np.ascontiguousarray(a)).cuda()

    shapes = {1: [36 * world], 2: [24, 20], 3: [12, 10, 12]}
    for dim in (1, 2, 3):
        # global mesh grows with the number of ranks along the decomposed dimensions
        if dim == 1:
            extents = L.decompose(world, shapes[1], [-1])
        else:
            base = shapes[dim]
            dec = (decomp2d if dim == 2 else decomp3d) or [-1] * dim
            probe = L.decompose(world, [b * world for b in base], dec)
            N = [base[a] * len(probe[a]) + (1 if (a == 0 and len(probe[a]) > 1) else 0) for a in range(dim)]
            extents = L.decompose(world, N, dec)  # uneven along x1 when it is split
        for G, order in ((2, 0), (3, 3)):
            # ---------------- ghost fill against the single-domain oracle
            gg, glob, infos, doms, flds = T._blocks(extents, G, 6, 3 + dim)
            o.comm_fields(gg, glob, 0, 6, [orc.FBC_PERIODIC] * 6)
            dom, I = doms[rank], infos[rank]
            ctx = eb.Context(dom.n, order=order, strict=True, device=local)
            md = L.make_metadomain(rank, extents)
            ctx.comm_init(md, fresh_uid())
            f = dev(np.nan_to_num(flds[rank], nan=-777.0))
            ctx.comm_fields(f, 0, 6, [I.face_fbc[k] for k in range(6)])
            torch.cuda.synchronize()
            check(np.array_equal(f.cpu().numpy(), T._window(glob, gg, I, dom, G)),
                  f"ghost fill dim {dim} G {G} extents {extents}")
            # component sub-range: only B must change
            f2 = dev(np.nan_to_num(flds[rank], nan=-777.0))
            ctx.comm_fields(f2, 3, 6, [I.face_fbc[k] for k in range(6)])
            torch.cuda.synchronize()
            h2 = f2.cpu().numpy()
            check(np.array_equal(h2[3:], T._window(glob, gg, I, dom, G)[3:])
                  and np.array_equal(h2[:3], np.nan_to_num(flds[rank], nan=-777.0)[:3]),
                  f"ghost fill of B only, dim {dim}")
            # ---------------- additive sync against the protocol restatement
            _, _, _, doms2, fl2 = T._blocks(extents, G, 3, 50 + dim, integer=True)
            mine = dev(fl2[rank])
            buff = torch.zeros_like(mine)
            mdcomm.exchange_fields_loopback(doms2, fl2, 0, 3, True)
            ctx.sync_currents(mine, buff, [I.face_fbc[k] for k in range(6)])
            torch.cuda.synchronize()
            check(np.array_equal(mine.cpu().numpy(), fl2[rank]), f"current sync dim {dim} G {G}")
            ctx.close()
        # -------------------- particle migration against the protocol restatement
        N = [sum(e) for e in extents]
        gg, p = T._global_particles(N, 4000, 21 + dim)
        infos = T._infos(extents, None, None)
        doms = [mdcomm.Domain(r, J, dim, 2) for r, J in enumerate(infos)]
        sets, counts = T._split_particles(p, infos, doms, cap=6000)
        dom, I = doms[rank], infos[rank]
        ctx = eb.Context(dom.n, order=0, strict=True, device=local, dx=1.0)
        ctx.comm_init(L.make_metadomain(rank, extents), fresh_uid())
        arr = to_device(sets[rank])
        npart = counts[rank]
        # payload planes (pld_r x2, pld_i x1) derived from the particle's weight, which nothing on this
        # path modifies: after every migration each alive particle must still carry its own payloads
        # (records carry them: src/kernels/comm.hpp:75-105)
        w_ = arr["weight"]
        arr["pld_r"] = torch.stack([2.0 * w_, w_ + 1.0]).contiguous()
        arr["pld_i"] = w_.view(torch.int32).clone().reshape(1, -1).contiguous()
        em0 = torch.zeros(ctx.grid.shape(6), dtype=torch.float32, device="cuda")
        for step in range(3):
            lb = mdcomm.Loopback()
            holes = []
            for J, dm, q, k in zip(infos, doms, sets, range(len(doms))):
                g = orc.Grid.make(dm.n, 2)
                pc = T._pusher(dim, [J.face_pbc[f] for f in range(6)])
                o.push(g, 0, pc, q, counts[k], np.zeros(g.shape(6), np.float32))
                nbr_n = {d: [infos[dm.neighbor[d]].n[a] for a in range(dim)] for d in range(dm.ndir)}
                out, h = mdcomm.particle_outbox(dm, q, counts[k], nbr_n)
                lb.post(dm.rank, out)
                holes.append(h)
            for dm, q, k in zip(doms, sets, range(len(doms))):
                counts[k] = mdcomm.particle_apply(dm, q, counts[k], holes[k],
                                                  lb.collect(dm.rank, mdcomm.particle_wanted(dm)))
            pbc = [I.face_pbc[f] for f in range(6)]
            gp = ctx.make_pusher(dt=0.45, omegaB0=0.7, mass=1.0, charge=-1.0, dx=1.0, pbc=pbc,
                                 tag_outgoing=int(any(b == eb.PBC_NONE for b in pbc[:2 * dim])))
            ctx.push(gp, arr, npart, em0)
            spc = (L.SpeciesC * 1)()
            spc[0].mass, spc[0].charge, spc[0].pusher_flags = 1.0, -1.0, eb.PUSHER_BORIS
            spc[0].npart, spc[0].maxnpart = npart, 6000
            spc[0].arrays = eb.Context.prtls_struct(arr)
            ctx._check(ctx.lib.eb200_comm_particles(ctx.handle, spc, 1, eb.Context._stream(None)))
            npart = int(spc[0].npart)
            torch.cuda.synchronize()
            check(npart == counts[rank], f"npart after migration dim {dim} step {step}: {npart} vs {counts[rank]}")
            got, want = to_host(arr, npart), sets[rank]
            names = ["ux1", "ux2", "ux3", "weight", "tag"]
            for a in range(1, dim + 1):
                names += [f"i{a}", f"dx{a}", f"i{a}_prev", f"dx{a}_prev"]
            alive = want.tag[:npart] == 1
            for nm in names:
                a, b = getattr(got, nm)[:npart], getattr(want, nm)[:npart]
                if nm != "tag":
                    a, b = a[alive], b[alive]
                check(np.array_equal(a, b), f"migration dim {dim} step {step}: array {nm}")
            wv = arr["weight"][:npart][torch.from_numpy(alive).to("cuda")]
            sel_ = torch.from_numpy(alive).to("cuda")
            check(bool(torch.equal(arr["pld_r"][0, :npart][sel_], 2.0 * wv))
                  and bool(torch.equal(arr["pld_r"][1, :npart][sel_], wv + 1.0))
                  and bool(torch.equal(arr["pld_i"][0, :npart][sel_], wv.view(torch.int32))),
                  f"migration dim {dim} step {step}: payloads did not travel with their particles")
        ctx.close()

    # ------------------------------------------------ whole step against the global oracle
    for dim, order, fused, mode in ((2, 0, True, eb.DEPOSIT_AGGREGATED), (2, 0, False, eb.DEPOSIT_ATOMIC),
                                    (3, 2, True, eb.DEPOSIT_AGGREGATED)):
        base = [24, 20] if dim == 2 else [12, 10, 12]
        dec = (decomp2d if dim == 2 else decomp3d) or [-1] * dim
        probe = L.decompose(world, [b * world for b in base], dec)
        N = [base[a] * len(probe[a]) for a in range(dim)]
        mdm = Metadomain(N, world, rank, dec)
        G = orc.nghosts_for(order)
        ppc0 = 8
        dx = 0.25
        scales = Scales(dim, dx, larmor0=2.0, skindepth0=1.0, ppc0=ppc0)
        gsim = pic.OracleSim(o, N, order, scales.derive(), dx, nfilter=2)
        rng = np.random.default_rng(77 + dim)
        from helpers import smooth_fields
        gsim.em[...] = smooth_fields(gsim.grid, 5, amp=0.3)
        ncell = int(np.prod(N))
        nper = ncell * ppc0 // 2
        sim = Simulation(mdm.local_n, order, scales, nfilter=2, fused=fused, deposit_mode=mode,
                         device=local)
        mdm.attach(sim, fresh_uid())
        I = mdm.info
        win = (slice(None),) + tuple(
            slice(G + I.cell_offset[a], G + I.cell_offset[a] + mdm.local_n[a]) for a in reversed(range(dim)))
        act = (slice(None),) + tuple(slice(G, G + mdm.local_n[a]) for a in reversed(range(dim)))
        sim.em[act] = dev(gsim.em[win])
        for charge in (-1.0, 1.0):
            ps = random_particles(gsim.grid, nper, int(100 + charge), umag=0.5)
            ps.weight[:] = 1.0
            gsim.add_species(1.0, charge, ps.copy(), nper)
            sel = np.ones(nper, bool)
            for a in range(dim):
                ia = getattr(ps, mdcomm.INT_NAMES[a])
                sel &= (ia >= I.cell_offset[a]) & (ia < I.cell_offset[a] + mdm.local_n[a])
            idx = np.nonzero(sel)[0]
            cap = int(idx.size * 1.5) + 64
            q = orc.ParticleSet(cap)
            for nm in ps.names():
                getattr(q, nm)[:idx.size] = getattr(ps, nm)[idx]
            for a in range(dim):
                getattr(q, mdcomm.INT_NAMES[a])[:idx.size] -= I.cell_offset[a]
                getattr(q, mdcomm.INT_NAMES[a] + "_prev")[:idx.size] -= I.cell_offset[a]
            keep = ["ux1", "ux2", "ux3", "weight", "tag"]
            for a in range(1, dim + 1):
                keep += [f"i{a}", f"dx{a}", f"i{a}_prev", f"dx{a}_prev"]
            arrs = {k: v for k, v in to_device(q).items() if k in keep}
            sim.add_species(1.0, charge, arrs, idx.size, maxnpart=cap)
        for step in range(5):
            sim.step()
            gsim.step()
        torch.cuda.synchronize()
        cnt = torch.tensor([float(sim.n_pushed())], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(cnt)
        alive = sum(int((sp.arrays["tag"][:sp.npart] == 1).sum()) for sp in sim.species)
        ca = torch.tensor([float(alive)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ca)
        check(int(ca.item()) == gsim.n_pushed(), f"alive particles {int(ca.item())} vs {gsim.n_pushed()} (dim {dim})")
        e, eo = sim.em.cpu().numpy()[act], gsim.em[win]
        j, jo = sim.cur.cpu().numpy()[act], gsim.cur[win]
        # fp32 tolerance: atomics / block-edge sums reorder the additions into J
        check(np.abs(j - jo).max() <= 5e-4 * np.abs(gsim.cur).max(), f"J after 5 steps dim {dim} fused {fused}: "
              f"{np.abs(j - jo).max():.3e} vs scale {np.abs(gsim.cur).max():.3e}")
        check(np.abs(e - eo).max() <= 5e-4 * np.abs(gsim.em).max(), f"E/B after 5 steps dim {dim} fused {fused}")
        # ghost cells too (last comm of the step fills E and J ghosts)
        ew = T._window(gsim.em, gsim.grid, I, mdcomm.Domain(rank, I, dim, G), G)
        check(np.abs(sim.em.cpu().numpy()[:3] - ew[:3]).max() <= 5e-4 * np.abs(gsim.em).max(), f"E ghosts dim {dim}")
        sim.ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    print(f"[rank {rank}] multi-domain parity ok (world {world})", flush=True)


if __name__ == "__main__":
    d2 = [int(x) for x in os.environ["EB200_DECOMP2D"].split(",")] if os.environ.get("EB200_DECOMP2D") else None
    d3 = [int(x) for x in os.environ["EB200_DECOMP3D"].split(",")] if os.environ.get("EB200_DECOMP3D") else None
    main(d2, d3)
