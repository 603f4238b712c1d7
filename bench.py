#!/usr/bin/env python
"""Benchmark of the PIC timestep hot path (push + deposit + fields) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU kernels

One "step" is one full SRPIC step (SRPICEngine::step_forward order) over the synthetic
reconnection plasma of BASELINE.json configs[1]: 2D Cartesian pair plasma, Harris sheets,
4096x2048 cells, 32 ppc, zig-zag deposit, 8 filter passes, periodic-core variant. The metric
is particle-steps/s = (pushed, alive particles) x steps / time, whole job over all GPUs.

Prints ONE JSON line (rank 0). See DESIGN.md "Measurement" for how each key is obtained.
"""
from __future__ import annotations

import argparse
import json
import os

import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-steps/sec (push+deposit+fields)"
UNIT = "particle-steps/s"
B_P_2D = 78.0  # algorithmic bytes per particle-step, fused push+deposit, 2D
WORKLOAD = "reconnection 2D Cartesian SR pair-plasma Harris sheet (periodic core), 4096x2048 cells, 32 ppc"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def reference_sample(nsteps, warmup, size=(512, 256), ppc0=32, nfilter=8, threads=None):
    """The reference's CPU kernels (oracle/_ref, else the oracle port) on a bounded cut of the
    workload; returns (particle-steps/s, info dict). Needs a CUDA device only to generate the
    synthetic state when one exists -- otherwise numpy."""
    import numpy as np
    from oracle import orc, pic
    if not os.path.exists(os.path.join(ROOT, "oracle", "liborc.so")):
        orc.build()
    impl = orc.reference(0)
    kind = "reference"
    if impl is None:
        impl, kind = orc.oracle(), "port"
    threads = threads or host_threads()
    used = impl.set_threads(threads)
    osim = cpu_reconnection(impl, size, ppc0, nfilter)
    for _ in range(warmup):
        osim.step()
    n0 = osim.n_pushed()
    t0 = time.perf_counter()
    for _ in range(nsteps):
        osim.step()
    dt = time.perf_counter() - t0
    val = n0 * nsteps / dt
    info = {"value": val, "unit": UNIT, "cores": used, "kind": kind,
            "sample": f"{size[0]}x{size[1]} cut of the workload, {ppc0} ppc, {n0} particles, "
                      f"{nsteps} full steps ({dt:.2f} s), oracle/_ref kernels"
                      if kind == "reference" else
                      f"{size[0]}x{size[1]} cut, {ppc0} ppc, {n0} particles, {nsteps} steps, serial port"}
    return val, info, dt / nsteps


def cpu_reconnection(impl, size, ppc0, nfilter, seed=0x5678):
    """Numpy twin of workloads.reconnection (same shape; own RNG) for the CPU arms."""
    import numpy as np
    from oracle import orc, pic
    from entity_b200.srpic import Scales
    n1, n2 = size
    Lx = 1000.0 * (n1 / 4096.0)
    dx = Lx / n1
    Ly = dx * n2
    scales = Scales(2, dx, larmor0=0.1, skindepth0=1.0, ppc0=ppc0).derive()
    o = pic.OracleSim(impl, size, 0, scales, dx, nfilter, xmin=(-0.5 * Lx, -0.5 * Ly, 0.0))
    g = o.grid
    w = 10.0 * (n1 / 4096.0)
    y1, y2 = -0.25 * Ly, 0.25 * Ly
    jj = np.arange(n2 + 2 * g.ng, dtype=np.float32) - g.ng
    y = (jj + 0.5) * dx - 0.5 * Ly
    bx = np.tanh((y - y1) / w) - np.tanh((y - y2) / w) - 1.0
    o.em[3] = (bx / dx)[:, None].astype(np.float32)
    rng = np.random.default_rng(seed)
    ncell = n1 * n2
    n_bg = ncell * (ppc0 // 2)
    n_cs = int(3.0 * 2.0 * (w / dx) * n1 * (ppc0 // 2)) * 2
    for charge in (-1.0, 1.0):
        n = n_bg + n_cs
        ps = orc.ParticleSet(n)
        cell = np.arange(n_bg) // (ppc0 // 2)
        ps.i1[:n_bg] = cell % n1
        ps.i2[:n_bg] = cell // n1
        ps.dx1[:n_bg] = rng.random(n_bg, dtype=np.float32)
        ps.dx2[:n_bg] = rng.random(n_bg, dtype=np.float32)
        for nm in ("ux1", "ux2", "ux3"):
            getattr(ps, nm)[:n_bg] = (1e-2 * rng.standard_normal(n_bg)).astype(np.float32)
        half = n_cs // 2
        for s, yc in enumerate((y1, y2)):
            lo = n_bg + s * half
            r = np.clip(rng.random(half), 1e-6, 1 - 1e-6)
            yy = np.mod(yc + 0.5 * w * np.log(r / (1 - r)) + 0.5 * Ly, Ly) / dx
            xx = rng.random(half) * n1
            for nm_i, nm_d, c, nn in (("i1", "dx1", xx, n1), ("i2", "dx2", yy, n2)):
                ci = np.clip(np.floor(c), 0, nn - 1)
                getattr(ps, nm_i)[lo:lo + half] = ci.astype(np.int32)
                getattr(ps, nm_d)[lo:lo + half] = np.clip(c - ci, 0, 0.99999994).astype(np.float32)
            T = 0.5 * 100.0 / 3.0
            mag = -T * np.log(np.clip(rng.random((3, half)), 1e-12, None).prod(axis=0))
            mu = 2 * rng.random(half) - 1
            ph = 2 * np.pi * rng.random(half)
            st = np.sqrt(1 - mu * mu)
            ps.ux1[lo:lo + half] = mag * st * np.cos(ph)
            ps.ux2[lo:lo + half] = mag * st * np.sin(ph)
            ps.ux3[lo:lo + half] = mag * mu
        np.minimum(ps.dx1, np.float32(0.99999994), out=ps.dx1)
        np.minimum(ps.dx2, np.float32(0.99999994), out=ps.dx2)
        ps.i1_prev[:], ps.i2_prev[:] = ps.i1, ps.i2
        ps.dx1_prev[:], ps.dx2_prev[:] = ps.dx1, ps.dx2
        ps.weight[:] = 1.0
        ps.tag[:] = 1
        o.add_species(1.0, charge, ps, n)
    return o


REF_CUT = (1024, 512)  # CPU-sized cut of configs[1]


def entity_xc_sample(flavour, size, nsteps, skip, ppc0=32, nfilter=8):
    """The UNMODIFIED reference (entity.xc built from /root/reference by oracle/build_entity_xc.sh,
    binaries under baseline/_ref/) on reconnection.toml at `size`, timed by its own per-step
    timers. Returns (info dict for the JSON line, raw result) or (None, why)."""
    from baseline import refrun
    if refrun.binary(flavour) is None:
        return None, f"baseline/_ref/{flavour}/entity_reconnection.xc not built"
    r = refrun.run(flavour, size, ppc0=ppc0, nfilter=nfilter, nsteps=nsteps, skip=skip)
    if r is None or "error" in r:
        return None, (r or {}).get("error", "no result")
    t = r["s_per_step"]
    info = {"value": r["path_pss"], "unit": UNIT, "cores": r["cores"], "kind": "reference",
            "sample": (f"entity.xc ({'Kokkos-OpenMP' if flavour.startswith('omp') else 'Kokkos-CUDA sm_100'}, "
                       f"pgens/reconnection incl. MATCH walls + replenish injector), {size[0]}x{size[1]} cells, "
                       f"{ppc0} ppc, {int(r['particles'])} particles, median of the reference's own timers over "
                       f"steps {r['steps_run'] - r['steps_timed']}..{r['steps_run'] - 1}: "
                       "ParticlePusher+CurrentDeposit+FieldSolver+CurrentFiltering+Communications+FieldBoundaries"),
            "push_deposit_only": r["pushdep_pss"], "whole_step_incl_injector": r["total_pss"],
            "ms_per_step": {k: 1e3 * v for k, v in t.items()}}
    return info, r


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    info, r = entity_xc_sample("omp", REF_CUT, args.steps + args.warmup, args.warmup)
    if info is None:
        # entity.xc not prebuilt: the reference's kernel headers compiled in place (oracle/_ref)
        val, info, sec_per_step = reference_sample(args.steps, args.warmup)
        info["note"] = f"entity.xc unavailable ({r}); oracle/_ref kernels with a std::thread driver"
    else:
        val, sec_per_step = info["value"], r["s_per_step"]["path"]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": info["sample"], "shape_order": 0,
                   "current_filters": 8, "host_threads": info["cores"]},
        "cpu_baseline": info,
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; entity_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from entity_b200 import workloads
    if args.walls and world > 1:
        raise SystemExit("bench.py --walls: single-GPU variant (the multi-domain bench tiles a periodic box)")

    size = tuple(args.size)
    import entity_b200 as eb
    dmode = {"atomic": eb.DEPOSIT_ATOMIC, "aggregated": eb.DEPOSIT_AGGREGATED,
             "ordered": eb.DEPOSIT_ORDERED}[args.deposit]
    turb = args.turbulence is not None
    other = args.shape  # "wald" | "magnetosphere": secondary bench lines (configs[4], configs[3])
    if other:
        if world > 1:
            raise SystemExit("bench.py --shape: single-GPU lines (curvilinear multi-domain exchange is not built)")
        args.sort_interval = 0
        args.no_e2e = True
        if other == "wald":
            size = tuple(args.size) if tuple(args.size) != (4096, 2048) else (512, 512)
            sim = workloads.wald(size, ppc=args.ppc if args.ppc != 32 else 8, device=local)
        else:
            size = tuple(args.size) if tuple(args.size) != (4096, 2048) else (2048, 1024)
            sim = workloads.magnetosphere(size, ppc=args.ppc if args.ppc != 32 else 10, device=local)
    elif turb:
        # configs[2]-shaped block (3D, T = 1 pair plasma, 3rd-order shapes, 4 filter passes) at the
        # n^3 given: a second bench line for the Esirkepov path, single GPU, not the headline
        size = (args.turbulence,) * 3
        if args.sort_interval is None:
            args.sort_interval = 5  # hot plasma, wide windows: the order decays within a few steps
        sim = workloads.turbulence(size, ppc0=args.ppc if args.ppc != 32 else 16, order=3, nfilter=4,
                                   fused=not args.unfused, sort_interval=args.sort_interval,
                                   device=local, deposit_mode=dmode, seed=0x9abc + rank,
                                   capacity_factor=1.0 if world == 1 else 1.25)
    else:
        if args.sort_interval is None:
            args.sort_interval = 20
        sim = workloads.reconnection(size, ppc0=args.ppc, nfilter=args.filters, fused=not args.unfused,
                                     sort_interval=args.sort_interval, device=local,
                                     deposit_mode=dmode, seed=0x5678 + rank, walls=args.walls,
                                     capacity_factor=(1.05 if args.replenish else 1.0) if world == 1 else 1.1)
    # i*_prev / dx*_prev are scratch of one step (written by the pusher, read by the deposit of the
    # same step, overwritten by the next push before any read: sr.hpp:137-153). The caller of the
    # step mirror says so (eb200_set_lean_prev) unless --keep-prev asks for the reference's state
    # of those arrays after every step. Only the fused 2D zig-zag kernel makes use of it.
    lean_prev = (not args.keep_prev) and not turb and not other and not args.unfused
    if lean_prev:
        sim.ctx.set_lean_prev(True)
    decomposition = [1] * len(size)
    if world > 1:
        # spatial block decomposition as the reference's reconnection.toml asks ([-1, 2]); every
        # GPU owns one block of `size` cells (weak scaling), fields and particles cross block
        # boundaries through the library's NCCL exchange
        from entity_b200 import lib as L
        from entity_b200.metadomain import Metadomain, bootstrap_unique_id
        if turb:
            dec = [-1, -1, -1]  # configs[2]: 2x2x2 blocks on 8 GPUs
        else:
            dec = [-1, 2] if world % 2 == 0 else [-1, 1]
        if args.decomp:
            dec = list(args.decomp)
        nd = [len(e) for e in L.decompose(world, [s_ * world for s_ in size], dec)]
        mdm = Metadomain(tuple(s_ * n_ for s_, n_ in zip(size, nd)), world, rank, dec)
        assert mdm.local_n == size
        mdm.attach(sim, bootstrap_unique_id())
        decomposition = nd
    n_pushed0 = sim.n_pushed()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The cell sort runs every `sort_interval` steps. So that `value` does not depend on whether
    # --steps happens to contain one: (1) the window is aligned to the sort period -- the state is
    # sorted once, then run untimed up to the point where the K timed steps see the steady-state
    # mean staleness of the particle order ((interval - 1) / 2 steps since the last sort); (2)
    # the sorts inside the window are timed by their own events and replaced by the amortised
    # charge sort_ms / sort_interval per step (sort_ms measured separately when the window holds
    # none).
    si = args.sort_interval
    K, W = args.steps, args.warmup

    def sort_all():
        for sp in sim.species:
            if sp.npart:
                # EB200_SORT_SKIP_PREV | EB200_SORT_UNSTABLE: as the step sorts on a fast-build context
                sim.ctx.sort_particles(sp.arrays, sp.npart, remove_dead=2 | 4)

    pre = 0
    if si > 0:
        sort_all()
        sim.step_index = 1  # "a sort ran at the end of step 0"
        t_start = 1 + ((si - K) // 2 if K < si else 0)  # index of the first timed step
        pre = (t_start - W - 1) % si
        for _ in range(pre):
            sim.step()
    for _ in range(W):
        sim.step()
    barrier()
    sim.profile(True)
    launches0 = sim.ctx.launch_count
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_first = sim.step_index
    ev0.record()
    n_replenished = 0
    if args.replenish:
        # the reconnection pgen's CustomPostStep after every step (density moment + two
        # ReplenishUniform injections), as the reference's time loop runs it
        for _ in range(K):
            sim.step()
            n_replenished += sim.replenish(sim.replenish_boxes)
    else:
        sim.step(K)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches = sim.ctx.launch_count - launches0
    prof = sim.read_profile()
    sort_in_ms = prof["ParticleSort"][0]  # the phase is bracketed every step; empty when no sort ran
    sorts_in = sum(1 for t_ in range(t_first, t_first + K) if si > 0 and t_ % si == 0)
    if si > 0 and sorts_in == 0:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        sort_all()
        e1.record()
        barrier()
        sort_ms = e0.elapsed_time(e1)
    else:
        sort_ms = sort_in_ms / max(sorts_in, 1)
    sim.profile(False)
    # particles advanced per step: alive particles of the pushed species (migration leaves holes)
    n_pushed = sum(int((sp.arrays["tag"][:sp.npart] == 1).sum()) for sp in sim.species
                   if sp.pusher != eb.PUSHER_NONE)
    ms_amortised = ms - sort_in_ms + (K * sort_ms / si if si > 0 else 0.0)
    t = torch.tensor([ms_amortised, ms, sort_ms], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([float(n_pushed)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_max, ms_raw, sort_ms = (float(x) for x in t.tolist())
    total = float(cnt.item())
    value = total * K / (ms_max * 1e-3)
    timing = {"window_ms": ms_raw, "sorts_in_window": int(sorts_in), "sort_ms": sort_ms,
              "sort_interval": si, "sort_charge_ms_per_step": sort_ms / si if si > 0 else 0.0,
              "ms_per_step_raw": ms_raw / K, "alignment_steps": pre,
              "pairs_replenished_per_step": n_replenished / K,
              "note": "ms_per_step = (window - sorts inside) / steps + sort_ms / sort_interval"}

    # roofline of the dominant kernel (fused push+deposit), from the in-step CUDA events
    peak, peak_src = measured_peaks()
    pd_ms, pd_calls = prof["PushDeposit"]
    per_launch_s = 1e-3 * pd_ms / max(args.steps, 1)  # one push+deposit phase per step
    b_p = 102.0 if turb else B_P_2D  # 78 B (2D), 102 B (3D) per particle-step
    if other:
        b_p = 86.0  # 2D + phi carried (read 4 + write 4)
    if lean_prev:
        b_p = B_P_2D - 16.0  # i_prev / dx_prev (4 + 4 B per dimension) are not written: 62 B
    achieved = n_pushed * b_p / per_launch_s / 1e9 if per_launch_s > 0 else 0.0
    with_sort_s = per_launch_s + 1e-3 * timing["sort_charge_ms_per_step"]
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None,
                "frac_with_sort_charged": (n_pushed * b_p / with_sort_s / 1e9 / peak) if with_sort_s > 0 else 0.0,
                "bytes_per_particle_step": b_p,
                "frac_at_78B": (n_pushed * B_P_2D / per_launch_s / 1e9 / peak) if (lean_prev and per_launch_s > 0) else None,
                "kernel": "push_deposit (all species of one step)", "peak_source": peak_src,
                "phase_ms_per_step": {k: v[0] / max(args.steps, 1) for k, v in prof.items()}}
    # DRAM bytes of one launch from the ncu full capture of THIS workload's kernel, else null
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        with open(tr) as f:
            roofline["traffic"] = json.load(f).get(
                f"{other}_push_deposit_bytes_per_launch" if other else
                "turbulence_push_deposit_bytes_per_launch" if turb else
                "push_deposit_lean_bytes_per_launch" if lean_prev else "push_deposit_bytes_per_launch")
    if other:
        roofline["note"] = ("two passes (push, deposit), ALU bound: the metric's transcendental functions "
                            "per particle (x pusher_niter for GRPIC), see DESIGN.md")

    # end to end through the host-buffer C-ABI entry (pinned host state, H2D + D2H per step)
    e2e = None
    if not args.no_e2e:
        hs = sim.host_state()
        sim.step_host(hs)  # warm-up (allocates the device mirrors)
        barrier()
        t0 = time.perf_counter()
        up = down = 0
        for _ in range(args.e2e_steps):
            u, d = sim.step_host(hs)
            up, down = u, d
        barrier()
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": total * args.e2e_steps / float(te.item()), "unit": UNIT,
               "h2d_bytes_per_step": up, "d2h_bytes_per_step": down, "steps": args.e2e_steps}
        del hs

    cpu = gpu_ref = None
    if rank == 0 and world == 1 and not args.no_cpu and not other:
        # the reference's Kokkos-OpenMP entity.xc on all host cores, its own timers
        cpu, why = entity_xc_sample("omp", REF_CUT, args.cpu_steps + 10, 10)
        if cpu is None:
            _, cpu, _ = reference_sample(args.cpu_steps, 1)
            cpu["note"] = f"entity.xc unavailable ({why}); oracle/_ref kernels with a std::thread driver"
    if rank == 0 and world == 1 and not args.no_gpu_ref and not turb and not other:
        # the reference's own sm_100 build (Kokkos-CUDA, Kokkos_ARCH_BLACKWELL100) on this GPU,
        # full-size reconnection.toml at 32 ppc: "the existing Blackwell kernel" 
        del sim
        torch.cuda.empty_cache()
        gpu_ref, why = entity_xc_sample("cuda", tuple(args.size), 30, 10, ppc0=args.ppc, nfilter=args.filters)
        if gpu_ref is None:
            gpu_ref = {"unavailable": str(why)[-300:]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "timing": timing,
            "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": (f"wald-with-species: 2D GRPIC qkerr_schild a=0.95, {size[0]}x{size[1]} cells, two Boris species in r in [2, 8], pusher_niter=10"
                                    if other == "wald" else
                                    f"magnetosphere: 2D qspherical SRPIC dipole, {size[0]}x{size[1]} cells, Boris+GCA, atmosphere gravity + injector every step, ATMOSPHERE/MATCH/AXIS field boundaries"
                                    if other == "magnetosphere" else
                                    f"turbulence-shaped 3D Cartesian SR pair plasma, {size[0]}^3 cells, "
                                    f"{16 if args.ppc == 32 else args.ppc} ppc, 3rd-order shapes (reduced from 1024^3)"
                                    if turb else
                                    WORKLOAD if size == (4096, 2048) and args.ppc == 32 else
                                    f"reconnection 2D {size[0]}x{size[1]} cells, {args.ppc} ppc (reduced)")
                       + ((" [x2 walls: fields MATCH ds=20, particles ABSORB, replenishing injector every step]"
                           if args.replenish else
                           " [x2 walls: fields MATCH ds=20, particles ABSORB, no injector]")
                          if args.walls else ""),
                       "cells_per_gpu": list(size),
                       "ppc0": ((8 if args.ppc == 32 else args.ppc) if other == "wald" else
                                (10 if args.ppc == 32 else args.ppc) if other == "magnetosphere" else
                                (16 if args.ppc == 32 else args.ppc) if turb else args.ppc),
                       "shape_order": 3 if turb else 0,
                       "current_filters": 4 if (turb or other) else args.filters,
                       "particles_per_gpu": n_pushed0,
                       "fused_push_deposit": (not args.unfused) and not other,
                       "sort_interval": args.sort_interval,
                       "prev_arrays": ("scratch of one step, not stored (eb200_set_lean_prev): 62 B per particle-step"
                                       if lean_prev else "left as the reference leaves them"),
                       "deposit": "atomic" if other else args.deposit,
                       "parallelism": (f"domain decomposition {'x'.join(str(n_) for n_ in decomposition)}, "
                                       f"one block per GPU, NCCL halo + particle exchange")
                       if world > 1 else "single domain",
                       "l2": "inputs larger than L2 (particle state >> 126 MB), no flush"},
            "roofline": roofline, "cpu_baseline": cpu, "gpu_baseline": gpu_ref, "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


_JSON_OUT = None


def claim_stdout():
    """stdout must carry exactly ONE JSON line: libraries that print there (NCCL's version
    banner, for one) are sent to stderr; the line itself goes to a private copy of fd 1."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, nargs=2, default=[4096, 2048])
    ap.add_argument("--ppc", type=int, default=32)
    ap.add_argument("--filters", type=int, default=8)
    ap.add_argument("--sort-interval", type=int, default=None,
                    help="cell sort every N steps (default 20 for the 2D workloads: the measured optimum "
                         "with the 12 ms counting sort, profiles/bench_r2k_sort_interval_*.json; 5 for --turbulence)")
    ap.add_argument("--deposit", default="aggregated", choices=["atomic", "aggregated", "ordered"])
    ap.add_argument("--unfused", action="store_true")
    ap.add_argument("--turbulence", type=int, default=None, metavar="N",
                    help="bench the turbulence-shaped 3D block (N^3 cells, 16 ppc, O=3) instead")
    ap.add_argument("--shape", default=None, choices=["wald", "magnetosphere"],
                    help="secondary bench lines: the GRPIC step (configs[4] with species) or the "
                         "curvilinear SRPIC step (configs[3]) instead of the reconnection headline")
    ap.add_argument("--walls", action="store_true",
                    help="x2 boundaries of reconnection.toml (fields MATCH, particles ABSORB) instead "
                         "of the periodic core; single GPU, no replenishing injector")
    ap.add_argument("--replenish", action="store_true",
                    help="with --walls: run the reconnection pgen's replenishing injector (density "
                         "moment + ReplenishUniform in the two 10-cell boxes) after every step")
    ap.add_argument("--decomp", type=int, nargs="+", default=None,
                    help="override the block decomposition request (default -1 2, as reconnection.toml)")
    ap.add_argument("--keep-prev", action="store_true",
                    help="leave i*_prev / dx*_prev after every step exactly as the reference does (78 B per "
                         "particle-step instead of 62; default: the caller declares them scratch, eb200_set_lean_prev)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-gpu-ref", action="store_true",
                    help="skip the reference's own Kokkos-CUDA sm_100 build (gpu_baseline)")
    ap.add_argument("--cpu-steps", type=int, default=10)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.replenish:
        args.walls = True
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
