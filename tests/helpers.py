"""Shared generators for the parity tests (seeded, numpy only)."""
import numpy as np

from oracle import orc


def random_fields(g, ncomp, seed, amp=1.0):
    rng = np.random.default_rng(seed)
    return (amp * rng.standard_normal(g.shape(ncomp))).astype(np.float32)


def smooth_fields(g, seed, amp=1.0):
    """6-component field made of a few Fourier modes (periodic on the active box)."""
    rng = np.random.default_rng(seed)
    shape = g.shape(6)
    dim = g.dim
    axes = []
    for a in range(dim):
        n = g.n[a]
        idx = (np.arange(n + 2 * g.ng) - g.ng) / n
        axes.append(idx)
    grids = np.meshgrid(*axes[::-1], indexing="ij")  # slowest..fastest
    out = np.zeros(shape, dtype=np.float64)
    for c in range(6):
        for _ in range(3):
            k = rng.integers(1, 4, size=dim)
            ph = rng.uniform(0, 2 * np.pi)
            arg = ph
            for a in range(dim):
                arg = arg + 2 * np.pi * k[a] * grids[dim - 1 - a]
            out[c] += rng.standard_normal() * np.sin(arg)
    return (amp * out).astype(np.float32)


def random_particles(g, n, seed, umag=1.0, dead_frac=0.0):
    rng = np.random.default_rng(seed)
    p = orc.ParticleSet(n)
    names_i = ["i1", "i2", "i3"]
    names_d = ["dx1", "dx2", "dx3"]
    for a in range(g.dim):
        getattr(p, names_i[a])[:] = rng.integers(0, g.n[a], size=n)
        d = rng.random(n, dtype=np.float32)
        d[d >= 1.0] = 0.0
        getattr(p, names_d[a])[:] = d
        getattr(p, names_i[a] + "_prev")[:] = getattr(p, names_i[a])
        getattr(p, names_d[a] + "_prev")[:] = getattr(p, names_d[a])
    for nm in ("ux1", "ux2", "ux3"):
        getattr(p, nm)[:] = (umag * rng.standard_normal(n)).astype(np.float32)
    p.weight[:] = rng.uniform(0.5, 1.5, n).astype(np.float32)
    p.tag[:] = 1
    if dead_frac > 0:
        p.tag[rng.random(n) < dead_frac] = 0
    return p


def assert_prtls_equal(a, b, npart=None, what=""):
    for nm in a.names():
        x, y = getattr(a, nm), getattr(b, nm)
        if npart is not None:
            x, y = x[:npart], y[:npart]
        if x.dtype.kind == "f":
            ok = np.array_equal(x.view(np.uint32), y.view(np.uint32))
        else:
            ok = np.array_equal(x, y)
        if not ok:
            bad = np.nonzero(x != y)[0]
            raise AssertionError(
                f"{what}: particle array {nm} differs at {bad.size} entries, "
                f"first {bad[:5]}: {x[bad[:5]]} vs {y[bad[:5]]}")


def assert_bits_equal(x, y, what=""):
    if not np.array_equal(x.view(np.uint32), y.view(np.uint32)):
        bad = np.argwhere(x != y)
        raise AssertionError(
            f"{what}: {bad.shape[0]} of {x.size} entries differ; first at {bad[:3].tolist()} "
            f"max abs diff {np.max(np.abs(x.astype(np.float64) - y.astype(np.float64))):.3e}")


# ------------------------------------------------------------------ torch <-> numpy bridges
def to_device(p, device="cuda"):
    """ParticleSet (numpy) -> dict of torch tensors on the device."""
    import torch
    out = {}
    for nm in p.names():
        out[nm] = torch.from_numpy(getattr(p, nm).copy()).to(device)
    return out


def to_host(arrays, n):
    """dict of torch tensors -> ParticleSet (numpy)."""
    p = orc.ParticleSet(n)
    for nm in p.names():
        getattr(p, nm)[:] = arrays[nm][:n].cpu().numpy()
    return p


def values_equal(x, y):
    """Exact value equality (NaN == NaN; +0 == -0)."""
    return np.array_equal(x, y, equal_nan=True)


def assert_values_equal(x, y, what=""):
    if not values_equal(x, y):
        bad = np.argwhere(~((x == y) | (np.isnan(x) & np.isnan(y))))
        d = np.abs(x.astype(np.float64) - y.astype(np.float64))
        raise AssertionError(
            f"{what}: {bad.shape[0]} of {x.size} entries differ; first at {bad[:3].tolist()}, "
            f"max abs diff {np.nanmax(d):.3e}")


def assert_prtls_values_equal(a, b, what=""):
    for nm in a.names():
        x, y = getattr(a, nm), getattr(b, nm)
        if not values_equal(x, y):
            bad = np.nonzero(x != y)[0]
            raise AssertionError(
                f"{what}: particle array {nm} differs at {bad.size} entries, "
                f"first {bad[:5]}: {x[bad[:5]]} vs {y[bad[:5]]}")
