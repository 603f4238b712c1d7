"""Host-side logic of the multi-domain path (CPU, no GPU): decomposition, neighbour and
boundary tables (the library's host code), and the exchange protocol restated in numpy
(oracle/mdcomm.py) against the single-domain oracle on the same global problem -- in one
process (loopback, up to 8 domains) and across two gloo processes."""
import itertools
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from entity_b200 import lib as L  # noqa: E402  (loads the .so; host-only calls, no GPU needed)
from oracle import mdcomm, orc  # noqa: E402
from helpers import random_particles  # noqa: E402


# ------------------------------------------------------------------ tools::Decompose
@pytest.mark.parametrize("nd,ncells,dec,expect", [
    (4, [100], [-1], [[25, 25, 25, 25]]),
    (3, [100], [-1], [[34, 33, 33]]),
    (8, [4096, 2048], [-1, 2], [[1024] * 4, [1024] * 2]),       # reconnection.toml decomposition
    (8, [1024, 1024, 1024], [-1, -1, -1], [[512, 512]] * 3),    # turbulence on 8 GPUs
    (6, [100, 37], [-1, -1], [[17, 17, 17, 17, 16, 16], [37]]),
    (2, [64, 64], [-1, -1], [[64], [32, 32]]),                  # sqrt(2) -> n1 = 1
    (4, [64, 64], [2, -1], [[32, 32], [32, 32]]),
    (12, [120, 60, 30], [-1, 2, -1], None),
])
def test_decompose(nd, ncells, dec, expect):
    got = L.decompose(nd, ncells, dec)
    assert int(np.prod([len(e) for e in got])) == nd
    for a, e in enumerate(got):
        assert sum(e) == ncells[a]
        assert max(e) - min(e) <= 1 and sorted(e, reverse=True) == e
    if expect is not None:
        assert got == expect


def test_decompose_errors():
    with pytest.raises(L.EB200Error):
        L.decompose(3, [64, 64], [2, -1])      # does not divide evenly
    with pytest.raises(L.EB200Error):
        L.decompose(16, [40], [-1])            # ncells < 5 per domain
    with pytest.raises(L.EB200Error):
        L.decompose(4, [64, 64], [2, 3])       # n1 * n2 != ndomains


# ------------------------------------------------------------- neighbours / boundaries
CASES = [
    ([[12, 12]], None, None),
    ([[12, 11], [10, 9, 9]], None, None),
    ([[12], [10, 10]], None, None),
    ([[10, 10], [9, 9], [8, 8]], None, None),
    ([[10, 10, 10], [9], [8, 8]], None, None),
    # reconnection-like: periodic x, non-periodic y (conductor/absorb)
    ([[12, 12], [10, 10]], [L.FBC_PERIODIC, L.FBC_PERIODIC, L.FBC_CONDUCTOR, L.FBC_CONDUCTOR, 0, 0],
     [L.PBC_PERIODIC, L.PBC_PERIODIC, L.PBC_ABSORB, L.PBC_ABSORB, 0, 0]),
]


def _infos(extents, fbc, pbc):
    n = int(np.prod([len(e) for e in extents]))
    return [L.domain_info(L.make_metadomain(r, extents, fbc, pbc)) for r in range(n)]


@pytest.mark.parametrize("extents,fbc,pbc", CASES)
def test_domain_tables(extents, fbc, pbc):
    dim = len(extents)
    infos = _infos(extents, fbc, pbc)
    nd = [len(e) for e in extents]
    nd_ = 3 ** dim
    for r, I in enumerate(infos):
        # index = o1 + nd1 * (o2 + nd2 * o3)  (tools::TensorProduct ordering)
        o = [I.offset[a] for a in range(dim)]
        lin = 0
        for a in reversed(range(dim)):
            lin = lin * nd[a] + o[a]
        assert lin == r
        for a in range(dim):
            assert I.n[a] == extents[a][o[a]]
            assert I.cell_offset[a] == sum(extents[a][:o[a]])
        for d in range(nd_):
            if d == (nd_ - 1) // 2:
                continue
            nb = I.neighbor[d]
            md = nd_ - 1 - d
            # Metadomain::finalValidityCheck: the neighbour's neighbour in -d is me
            assert infos[nb].neighbor[md] == r
            # what I send in d, the neighbour expects from -d
            assert bool(I.enabled[d]) == bool(infos[nb].enabled[md])
            if I.dir_fbc[d] == L.FBC_PERIODIC:
                assert nb == r
            if I.dir_fbc[d] == L.FBC_SYNC:
                assert nb != r
        for a in range(dim):
            for side in range(2):
                edge = o[a] == (0 if side == 0 else nd[a] - 1)
                gf = (fbc or [L.FBC_PERIODIC] * 6)[2 * a + side]
                if not edge or (gf == L.FBC_PERIODIC and nd[a] > 1):
                    assert I.face_fbc[2 * a + side] == L.FBC_SYNC
                    assert I.face_pbc[2 * a + side] == L.PBC_NONE
                else:
                    assert I.face_fbc[2 * a + side] == gf


# ------------------------------------------------------------------- helpers: blocks
def _blocks(extents, G, ncomp, seed, integer=False):
    """Random global field (ghost-inclusive) and its decomposition into blocks."""
    dim = len(extents)
    N = [sum(e) for e in extents]
    rng = np.random.default_rng(seed)
    gg = orc.Grid.make(N, G)
    glob = rng.standard_normal(gg.shape(ncomp)).astype(np.float32)
    infos = _infos(extents, None, None)
    doms, flds = [], []
    for r, I in enumerate(infos):
        dom = mdcomm.Domain(r, I, dim, G)
        g = orc.Grid.make(dom.n, G)
        if integer:
            f = rng.integers(-8, 9, size=g.shape(ncomp)).astype(np.float32)
        else:
            f = np.full(g.shape(ncomp), np.nan, dtype=np.float32)
            src = (slice(None),) + tuple(
                slice(G + I.cell_offset[a], G + I.cell_offset[a] + dom.n[a]) for a in reversed(range(dim)))
            dst = (slice(None),) + tuple(slice(G, G + dom.n[a]) for a in reversed(range(dim)))
            f[dst] = glob[src]
        doms.append(dom)
        flds.append(f)
    return gg, glob, infos, doms, flds


def _window(glob, gg, I, dom, G):
    """The block's ghost-inclusive window of the ghost-filled global array (periodic)."""
    dim = dom.dim
    idx = []
    for a in reversed(range(dim)):
        N = gg.n[a]
        loc = np.arange(-G, dom.n[a] + G) + I.cell_offset[a]
        idx.append(np.mod(loc, N) + G)
    return glob[(slice(None),) + np.ix_(*idx)]


FIELD_CASES = [[[12, 12]], [[9, 8, 8]], [[12, 11], [10, 9, 9]], [[12], [10, 10]], [[10, 10], [24]],
               [[8, 8], [7, 7], [6, 6]], [[8, 8, 8], [9], [6, 6]]]


@pytest.mark.parametrize("extents", FIELD_CASES)
@pytest.mark.parametrize("G", [2, 3])
def test_ghost_fill_matches_single_domain(extents, G):
    gg, glob, infos, doms, flds = _blocks(extents, G, 6, 3)
    orc.oracle().comm_fields(gg, glob, 0, 6, [orc.FBC_PERIODIC] * 6)  # single-domain reference
    mdcomm.exchange_fields_loopback(doms, flds, 0, 6, False)
    for I, dom, f in zip(infos, doms, flds):
        assert np.array_equal(f, _window(glob, gg, I, dom, G)), f"rank {dom.rank}"


@pytest.mark.parametrize("extents", FIELD_CASES)
def test_current_sync_sums_every_deposit_once(extents):
    """Integer-valued blocks: after the additive sync every active cell holds the exact sum of
    all cells (ghosts included) of all blocks that map onto it."""
    G = 2
    gg, _, infos, doms, flds = _blocks(extents, G, 3, 5, integer=True)
    dim = len(extents)
    N = [gg.n[a] for a in range(dim)]
    expect = np.zeros((3, *N[::-1]), dtype=np.float64)
    for I, dom, f in zip(infos, doms, flds):
        idx = []
        for a in reversed(range(dim)):
            idx.append(np.mod(np.arange(-G, dom.n[a] + G) + I.cell_offset[a], N[a]))
        for c in range(3):
            np.add.at(expect[c], np.ix_(*idx), f[c].astype(np.float64))
    before = [f.copy() for f in flds]
    mdcomm.exchange_fields_loopback(doms, flds, 0, 3, True)
    for I, dom, f, f0 in zip(infos, doms, flds, before):
        act = (slice(None),) + tuple(slice(G, G + dom.n[a]) for a in reversed(range(dim)))
        win = (slice(None),) + tuple(
            slice(I.cell_offset[a], I.cell_offset[a] + dom.n[a]) for a in reversed(range(dim)))
        assert np.array_equal(f[act].astype(np.float64), expect[win]), f"rank {dom.rank}"
        ghost = np.ones(f.shape, bool)
        ghost[act] = False
        assert np.array_equal(f[ghost], f0[ghost])  # ghosts are untouched by the sync


# ----------------------------------------------------------------------- particles
def _global_particles(N, npart, seed):
    g = orc.Grid.make(N, 2)
    p = random_particles(g, npart, seed, umag=3.0, dead_frac=0.1)
    return g, p


def _split_particles(p, infos, doms, cap):
    dim = doms[0].dim
    sets, counts = [], []
    for I, dom in zip(infos, doms):
        sel = np.ones(p.n, bool)
        for a in range(dim):
            ia = getattr(p, mdcomm.INT_NAMES[a])
            sel &= (ia >= I.cell_offset[a]) & (ia < I.cell_offset[a] + dom.n[a])
        idx = np.nonzero(sel)[0]
        q = orc.ParticleSet(cap)
        for nm in p.names():
            getattr(q, nm)[:idx.size] = getattr(p, nm)[idx]
        for a in range(dim):
            getattr(q, mdcomm.INT_NAMES[a])[:idx.size] -= I.cell_offset[a]
            getattr(q, mdcomm.INT_NAMES[a] + "_prev")[:idx.size] -= I.cell_offset[a]
        sets.append(q)
        counts.append(idx.size)
    return sets, counts


def _rows(p, n, dim, offset=None):
    cols = []
    alive = p.tag[:n] == 1
    for a in range(dim):
        ia = getattr(p, mdcomm.INT_NAMES[a])[:n].astype(np.int64)
        if offset is not None:
            ia = ia + offset[a]
        cols.append(ia[alive])
        cols.append(getattr(p, mdcomm.DX_NAMES[a])[:n][alive].view(np.uint32).astype(np.int64))
    for nm in ("ux1", "ux2", "ux3", "weight"):
        cols.append(getattr(p, nm)[:n][alive].view(np.uint32).astype(np.int64))
    rows = np.stack(cols, axis=1)
    return rows[np.lexsort(rows.T[::-1])]


def _pusher(dim, pbc):
    return orc.make_pusher(dt=0.45, omegaB0=0.7, mass=1.0, charge=-1.0, dx=1.0, pbc=pbc,
                           tag_outgoing=int(any(b == orc.PBC_NONE for b in pbc[:2 * dim])))


@pytest.mark.parametrize("extents", [[[12, 12]], [[12, 11], [10, 9, 9]], [[12], [10, 10]],
                                     [[8, 8], [7, 7], [6, 6]]])
def test_particle_migration_matches_periodic_wrap(extents):
    """Push (zero fields: straight lines) + migration over the blocks == push with periodic
    wrap on the global domain, as a multiset of particles, over several steps."""
    dim = len(extents)
    N = [sum(e) for e in extents]
    gg, p = _global_particles(N, 3000, 11)
    infos = _infos(extents, None, None)
    doms = [mdcomm.Domain(r, I, dim, 2) for r, I in enumerate(infos)]
    sets, counts = _split_particles(p, infos, doms, cap=4000)
    o = orc.oracle()
    em_g = np.zeros(gg.shape(6), np.float32)
    glob_ctx = _pusher(dim, [orc.PBC_PERIODIC] * 6)
    for step in range(4):
        o.push(gg, 0, glob_ctx, p, p.n, em_g)
        lb = mdcomm.Loopback()
        holes = []
        for I, dom, q, k in zip(infos, doms, sets, range(len(doms))):
            g = orc.Grid.make(dom.n, 2)
            ctx = _pusher(dim, [I.face_pbc[f] for f in range(6)])
            o.push(g, 0, ctx, q, counts[k], np.zeros(g.shape(6), np.float32))
            nbr_n = {d: [infos[dom.neighbor[d]].n[a] for a in range(dim)] for d in range(dom.ndir)}
            out, h = mdcomm.particle_outbox(dom, q, counts[k], nbr_n)
            lb.post(dom.rank, out)
            holes.append(h)
        for dom, q, k in zip(doms, sets, range(len(doms))):
            inbox = lb.collect(dom.rank, mdcomm.particle_wanted(dom))
            counts[k] = mdcomm.particle_apply(dom, q, counts[k], holes[k], inbox)
        want = _rows(p, p.n, dim)
        got = np.concatenate([_rows(q, counts[k], dim, [I.cell_offset[a] for a in range(dim)])
                              for k, (I, q) in enumerate(zip(infos, sets))])
        got = got[np.lexsort(got.T[::-1])]
        assert want.shape == got.shape and np.array_equal(want, got), f"step {step}"
        for I, dom, q, k in zip(infos, doms, sets, range(len(doms))):
            alive = q.tag[:counts[k]] == 1
            for a in range(dim):
                ia = getattr(q, mdcomm.INT_NAMES[a])[:counts[k]][alive]
                assert ((ia >= 0) & (ia < dom.n[a])).all()


# ------------------------------------------------------------------- two gloo ranks
def _gloo_worker(rank, world, port, extents, q):
    try:
        import torch.distributed as dist
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                                world_size=world)
        G = 2
        dim = len(extents)
        gg, glob, infos, doms, flds = _blocks(extents, G, 6, 3)  # same seed on both ranks
        orc.oracle().comm_fields(gg, glob, 0, 6, [orc.FBC_PERIODIC] * 6)
        dom, f = doms[rank], flds[rank]
        mdcomm.exchange_fields_dist(dom, f, 0, 6, False)
        ok = np.array_equal(f, _window(glob, gg, infos[rank], dom, G))
        # additive sync against the loopback result
        _, _, _, doms2, fl2 = _blocks(extents, G, 3, 5, integer=True)
        mine = fl2[rank].copy()
        mdcomm.exchange_fields_loopback(doms2, fl2, 0, 3, True)
        mdcomm.exchange_fields_dist(doms2[rank], mine, 0, 3, True)
        ok = ok and np.array_equal(mine, fl2[rank])
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, bool(ok), ""))
    except Exception as e:  # pragma: no cover
        q.put((rank, False, repr(e)))


@pytest.mark.parametrize("extents", [[[12, 12]], [[12], [10, 10]], [[9, 8], [11]]])
def test_two_rank_gloo_exchange(extents):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, extents, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, msg in res:
        assert ok, f"rank {rank}: {msg}"
