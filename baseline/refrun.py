"""Runs the UNMODIFIED reference (entity.xc built from /root/reference by
oracle/build_entity_xc.sh, binaries under baseline/_ref/<flavour>/) on the reconnection problem
of BASELINE.json configs[1] and reads the reference's OWN per-step timers
(src/global/utils/diag.cpp, `[diagnostics] interval = 1, blocking_timers = true`).

Used by bench.py only: `--impl reference` / `cpu_baseline` (Kokkos-OpenMP build, BASELINE.md
section 3) and `gpu_baseline` (the reference's Kokkos-CUDA sm_100 build on the same B200).
Nothing of this repo's kernels, engine or oracle is on that path.
"""
from __future__ import annotations

import os
import re
import shutil
import statistics
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")

TOML = """[simulation]
  name    = "rec"
  engine  = "srpic"
  runtime = {runtime}

[grid]
  resolution = [{n1}, {n2}]
  extent     = [[{x0}, {x1}], [{y0}, {y1}]]

  [grid.metric]
    metric = "minkowski"

  [grid.boundaries]
    fields    = [["PERIODIC"], ["MATCH", "MATCH"]]
    particles = [["PERIODIC"], ["ABSORB", "ABSORB"]]

    [grid.boundaries.match]
      ds = [[10.0], [20.0]]

[scales]
  larmor0    = 0.1
  skindepth0 = 1.0

[algorithms]
  current_filters = {nfilter}

  [algorithms.timestep]
    CFL = 0.5

[particles]
  ppc0 = {ppc0}
{sort_line}
  [[particles.species]]
    label    = "e-"
    mass     = 1.0
    charge   = -1.0
    maxnpart = {maxnpart}

  [[particles.species]]
    label    = "e+"
    mass     = 1.0
    charge   = 1.0
    maxnpart = {maxnpart}

[setup]
  bg_B           = 1.0
  bg_Bguide      = 0.0
  bg_temperature = 1e-4
  inj_ypad       = 50.0
  cs_width       = 10.0
  cs_overdensity = 3.0

[diagnostics]
  interval        = 1
  blocking_timers = true
  colored_stdout  = false

[output]
  enable = false
"""

_UNITS = {"ns": 1e-9, "µs": 1e-6, "us": 1e-6, "ms": 1e-3, "s": 1.0, "min": 60.0}
_TIMER = re.compile(r"^\s+([A-Za-z]+)\.+\s*([0-9.]+) (ns|µs|us|ms|s|min)\b")


def binary(flavour: str, pgen: str = "reconnection"):
    p = os.path.join(REF, flavour, f"entity_{pgen}.xc")
    return p if os.path.exists(p) else None


def reconnection_toml(size, ppc0=32, nfilter=8, nsteps=30, maxnpart=None, sort_interval=0):
    """reconnection.toml (pgens/reconnection/reconnection.toml) at `size` cells with the cell
    size of the 4096x2048 original (dx = 1000/4096) and the named ppc0."""
    n1, n2 = size
    dx = 1000.0 / 4096.0
    lx, ly = dx * n1, dx * n2
    dt = 0.5 * dx / 2.0 ** 0.5
    ncell = n1 * n2
    if maxnpart is None:
        # per species: background ppc0/2 per cell + the sheet (overdensity 3 over ~2 widths) + headroom
        maxnpart = int(ncell * ppc0 * 0.5 * 1.35 + 3.0 * 4.0 * (10.0 / dx) * n1 * ppc0 * 0.5)
    return TOML.format(runtime=f"{(nsteps + 0.5) * dt:.6f}", n1=n1, n2=n2, x0=-0.5 * lx, x1=0.5 * lx,
                       y0=-0.5 * ly, y1=0.5 * ly, nfilter=nfilter, ppc0=float(ppc0),
                       sort_line=(f"  spatial_sorting_interval = {int(sort_interval)}\n" if sort_interval else ""),
                       maxnpart=f"{maxnpart:.4e}")


def parse_timers(text: str):
    """list (one per step) of {timer: seconds}"""
    steps, cur = [], None
    for line in text.splitlines():
        if line.startswith("Step:"):
            cur = {}
            steps.append(cur)
            continue
        if cur is None:
            continue
        m = _TIMER.match(line)
        if m:
            cur[m.group(1)] = float(m.group(2)) * _UNITS[m.group(3)]
    return [s for s in steps if "ParticlePusher" in s]


def parse_npart(count_file: str):
    """exact particle counts per step (all species), written by the post-step hook of the
    wrapper pgen (oracle/pgens/dump_common.hpp, EB_COUNT_FILE)"""
    out = []
    for r in open(count_file).read().splitlines():
        f = r.split()
        if len(f) > 1:
            out.append(float(sum(int(x) for x in f[1:])))
    return out


def run(flavour, size, ppc0=32, nfilter=8, nsteps=30, skip=10, threads=None, timeout=1500, keep=None,
        sort_interval=0):
    """Returns a dict: particle-steps/s from the reference's own timers, median over steps
    [skip, nsteps): `pushdep` = ParticlePusher + CurrentDeposit (BASELINE.md section 3 headline),
    `path` = + FieldSolver + CurrentFiltering + Communications + FieldBoundaries, `total` =
    everything the reference timed in the step (incl. its injector / Custom)."""
    exe = binary(flavour)
    if exe is None:
        return None
    tmp = keep or tempfile.mkdtemp(prefix="eb_ref_")
    os.makedirs(tmp, exist_ok=True)
    with open(os.path.join(tmp, "rec.toml"), "w") as f:
        f.write(reconnection_toml(size, ppc0, nfilter, nsteps, sort_interval=sort_interval))
    env = dict(os.environ, EB_COUNT_FILE=os.path.join(tmp, "counts.txt"))
    ncores = threads or len(os.sched_getaffinity(0))
    if flavour.startswith("omp"):
        env.update(OMP_NUM_THREADS=str(ncores), OMP_PROC_BIND="spread", OMP_PLACES="threads")
    else:
        env["LD_LIBRARY_PATH"] = ":".join(filter(None, [
            "/usr/local/cuda/lib64", os.path.join(os.path.dirname(os.__file__), "site-packages",
                                                  "nvidia", "cuda_runtime", "lib"),
            env.get("LD_LIBRARY_PATH", "")]))
    try:
        r = subprocess.run([exe, "-input", "rec.toml"], cwd=tmp, env=env, capture_output=True,
                           text=True, timeout=timeout)
        text = r.stdout
        if r.returncode != 0:
            return {"error": (r.stdout[-600:] + r.stderr[-600:]).strip()}
        steps = parse_timers(text)
        npart = []
        sc = os.path.join(tmp, "counts.txt")
        if os.path.exists(sc):
            npart = parse_npart(sc)
    finally:
        if keep is None:
            shutil.rmtree(tmp, ignore_errors=True)
    if len(steps) <= skip:
        skip = max(0, len(steps) // 3)
    use = steps[skip:]
    if not use or not npart:
        return {"error": f"could not parse the reference's output ({len(steps)} steps, {len(npart)} stats rows)"}
    med = lambda keys: statistics.median(sum(s.get(k, 0.0) for k in keys) for s in use)
    n = npart[-1]
    mean = lambda keys: statistics.fmean(sum(s.get(k, 0.0) for k in keys) for s in use)
    pushdep = med(["ParticlePusher", "CurrentDeposit"])
    path = med(["ParticlePusher", "CurrentDeposit", "FieldSolver", "CurrentFiltering",
                "Communications", "FieldBoundaries"])
    total = med(["ParticlePusher", "CurrentDeposit", "FieldSolver", "CurrentFiltering",
                 "Communications", "FieldBoundaries", "Injector", "Custom", "ParticleSort",
                 "ParticleBoundaries"])
    return {"particles": n, "steps_timed": len(use), "steps_run": len(steps),
            "s_per_step": {"pushdep": pushdep, "path": path, "total": total,
                           "pusher": med(["ParticlePusher"]), "deposit": med(["CurrentDeposit"]),
                           "fieldsolver": med(["FieldSolver"]), "filter": med(["CurrentFiltering"]),
                           "custom_injector": med(["Custom", "Injector"]),
                           "sort_mean": mean(["ParticleSort"]),
                           "total_mean": mean(["ParticlePusher", "CurrentDeposit", "FieldSolver",
                                               "CurrentFiltering", "Communications", "FieldBoundaries",
                                               "Injector", "Custom", "ParticleSort", "ParticleBoundaries"])},
            "pushdep_pss": n / pushdep, "path_pss": n / path, "total_pss": n / total,
            "cores": ncores if flavour.startswith("omp") else None, "size": list(size), "ppc0": ppc0}


if __name__ == "__main__":
    import json
    import sys
    fl = sys.argv[1] if len(sys.argv) > 1 else "omp"
    sz = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (512, 256)
    ns = int(sys.argv[4]) if len(sys.argv) > 4 else 20
    print(json.dumps(run(fl, sz, nsteps=ns), indent=1))
