"""Builds entity_b200/libentity_b200.so in-tree with nvcc for sm_100a.

Every kernel translation unit that does floating-point work is compiled twice: with
``-DEB200_STRICT=1 --fmad=false`` (bit-exact with the reference's baseline CPU build) and with
nvcc's default FMA contraction (throughput path). ``python -m entity_b200.build`` or
``__graft_entry__.build()`` runs this; the .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libentity_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
          "-Xcompiler", "-fno-fast-math", "-Xcompiler", "-ffp-contract=off"]

# (source, object suffix, extra flags)
UNITS = [
    ("particles.cu", "strict", ["-DEB200_STRICT=1", "--fmad=false"]),
    ("particles.cu", "fast", ["-DEB200_STRICT=0", "-ftz=true"]),
    ("fields.cu", "strict", ["-DEB200_STRICT=1", "--fmad=false"]),
    ("fields.cu", "fast", ["-DEB200_STRICT=0"]),
    ("curv.cu", "one", ["--fmad=false"]),
    ("sort.cu", "one", []),
    ("stats.cu", "one", ["--fmad=false"]),
    ("bcs.cu", "one", ["--fmad=false"]),
    ("output.cu", "one", ["--fmad=false"]),
    ("inject.cu", "one", ["--fmad=false"]),
    ("comm.cu", "one", []),
    ("engine.cu", "one", []),
    ("capi.cu", "one", []),
]


def _deps():
    out = []
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h")):
                out.append(os.path.join(root, f))
    return out


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _compile(unit, verbose):
    src, tag, extra = unit
    srcp = os.path.join(CSRC, src)
    if not os.path.exists(srcp):
        return None
    obj = os.path.join(OBJ, f"{os.path.splitext(src)[0]}.{tag}.o")
    if _stale(obj, _deps()):
        cmd = [NVCC, *ARCH, *COMMON, *extra, "-c", srcp, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src} [{tag}]:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
    return obj


def build_variant(name: str, defines: list[str]) -> str:
    """Experiment helper: a second copy of the library with extra -D flags on particles.cu
    (fast variant only), written to build/variants/<name>.so; select it with EB200_LIB."""
    vdir = os.path.join(HERE, "variants")
    os.makedirs(vdir, exist_ok=True)
    build()
    obj = os.path.join(vdir, f"particles.fast.{name}.o")
    cmd = [NVCC, *ARCH, *COMMON, "-DEB200_STRICT=0", *defines, "-Xptxas=-v", "-c",
           os.path.join(CSRC, "particles.cu"), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr)
    info = [l for l in r.stderr.split("\n")]
    for k, l in enumerate(info):
        if "push_deposit_kernelILi2ELi0ELb1" in l and "Compiling" in l:
            print(name, info[k + 2].strip(), "|", info[k + 3].strip() if k + 3 < len(info) else "")
    objs = []
    for src, tag, _ in UNITS:
        o = os.path.join(OBJ, f"{os.path.splitext(src)[0]}.{tag}.o")
        if os.path.exists(o):
            objs.append(obj if (src, tag) == ("particles.cu", "fast") else o)
    out = os.path.join(vdir, f"{name}.so")
    r = subprocess.run([NVCC, *ARCH, "-shared", "-o", out, *objs, "-lcudart_static", "-ldl",
                        "-lrt", "-lpthread"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr)
    return out


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = [o for o in ex.map(lambda u: _compile(u, verbose), UNITS) if o]
    if _stale(LIB, objs):
        cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart_static", "-ldl", "-lrt",
               "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
