"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's multi-domain exchange.

Restates, for numpy blocks (fields in C order = i1 fastest, component slowest):

* ``Metadomain::CommunicateFields``   src/framework/domain/metadomain_comm.cpp:205-367
* ``Metadomain::SynchronizeFields``   src/framework/domain/metadomain_comm.cpp:409-562
* ``GetSendRecvParams`` slabs         src/framework/domain/metadomain_comm.cpp:118-203
* ``Metadomain::CommunicateParticles`` + ``Particles::Communicate``
  src/framework/domain/metadomain_comm.cpp:565-653, src/framework/containers/particles_comm.cpp:180-389,
  src/kernels/comm.hpp:75-363 (with the deterministic ordering the CUDA library defines:
  index order inside a send tag; holes = dead slots first, then vacated slots by tag)

direction by direction, exactly like the reference loops over ``dir::Directions<D>::all``. The
neighbour / boundary tables are taken from the caller (``eb200_domain_info``, the product's
host logic, which is what the CPU tests exercise); the data movement here is independent
numpy code. Transports: an in-process loopback (all domains in one process) and
``torch.distributed`` point-to-point (gloo on CPU).

Parity unpinned by the reference itself: it has no multi-domain golden vectors and MPI is not
installed here, so this restatement is checked against the single-domain oracle on the same
global problem (tests/test_metadomain.py).
"""
from __future__ import annotations

import numpy as np


def ndir(dim):
    return 3 ** dim


def dir_vec(dim, lin):
    d = [0] * dim
    for a in range(dim - 1, -1, -1):
        d[a] = lin % 3 - 1
        lin //= 3
    return d


def dir_index(d):
    lin = 0
    for x in d:
        lin = lin * 3 + (x + 1)
    return lin


def _slab(n, G, d, sync, recv):
    """Index ranges (per dimension, ghost-inclusive) of GetSendRecvParams."""
    out = []
    for a, na in enumerate(n):
        s = -d[a] if recv else d[a]
        if not sync:
            if s == 0:
                out.append((G, G + na))
            elif s == 1:
                out.append((G + na, 2 * G + na) if recv else (na, na + G))
            else:
                out.append((0, G) if recv else (G, 2 * G))
        else:
            if s == 0:
                out.append((0, na + 2 * G))
            elif s == 1:
                out.append((na, na + 2 * G))
            else:
                out.append((0, 2 * G))
    return out


def _view(fld, ranges, c0, c1):
    sl = (slice(c0, c1),) + tuple(slice(lo, hi) for lo, hi in ranges[::-1])
    return fld[sl]


class Domain:
    """One block: tables (from eb200_domain_info) + local extents."""

    def __init__(self, rank, info, dim, G):
        self.rank, self.dim, self.G = rank, dim, G
        self.n = [info.n[a] for a in range(dim)]
        self.neighbor = [info.neighbor[k] for k in range(27)]
        self.enabled = [bool(info.enabled[k]) for k in range(27)]
        self.ndir = ndir(dim)
        self.centre = (self.ndir - 1) // 2


class Loopback:
    """All domains live in this process: post() every rank's outbox, then collect()."""

    def __init__(self):
        self.box = {}

    def post(self, src, outbox):
        for (dst, key), arr in outbox.items():
            self.box[(src, dst, key)] = arr

    def collect(self, dst, wanted):
        return {(src, key): self.box.pop((src, dst, key)) for (src, key) in wanted}


def gloo_exchange(rank, outbox, wanted, shapes):
    """torch.distributed point-to-point: outbox {(dst, key): array}, wanted [(src, key)] with
    shapes {(src, key): (shape, dtype)}. Messages between the same pair are ordered by key."""
    import torch
    import torch.distributed as dist
    ops, bufs = [], {}
    for (dst, key) in sorted(outbox):
        arr = np.ascontiguousarray(outbox[(dst, key)])
        if dst == rank:
            bufs[(rank, key)] = arr.copy()
            continue
        t = torch.from_numpy(arr.view(np.uint8).reshape(-1).copy())
        ops.append(dist.P2POp(dist.isend, t, dst))
    for (src, key) in sorted(wanted):
        if src == rank:
            continue
        shape, dt = shapes[(src, key)]
        nb = int(np.prod(shape)) * np.dtype(dt).itemsize
        t = torch.empty(nb, dtype=torch.uint8)
        bufs[(src, key)] = (t, shape, dt)
        ops.append(dist.P2POp(dist.irecv, t, src))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    out = {}
    for k, v in bufs.items():
        if isinstance(v, tuple):
            t, shape, dt = v
            out[k] = t.numpy().view(dt).reshape(shape).copy()
        else:
            out[k] = v
    return out


# ------------------------------------------------------------------------------ fields
def field_outbox(dom: Domain, fld, c0, c1, sync):
    """What this domain sends: {(peer, direction index): slab copy}."""
    out = {}
    for d in range(dom.ndir):
        if d == dom.centre or not dom.enabled[d]:
            continue
        r = _slab(dom.n, dom.G, dir_vec(dom.dim, d), sync, False)
        out[(dom.neighbor[d], d)] = _view(fld, r, c0, c1).copy()
    return out


def field_wanted(dom: Domain, c0, c1, sync):
    """[(src, d)] and shapes of what this domain receives, in iteration-direction order."""
    wanted, shapes = [], {}
    for d in range(dom.ndir):
        md = dom.ndir - 1 - d
        if d == dom.centre or not dom.enabled[md]:
            continue
        r = _slab(dom.n, dom.G, dir_vec(dom.dim, d), sync, True)
        shape = (c1 - c0,) + tuple(hi - lo for lo, hi in r[::-1])
        key = (dom.neighbor[md], d)
        wanted.append(key)
        shapes[key] = (shape, np.float32)
    return wanted, shapes


def field_apply(dom: Domain, fld, c0, c1, sync, inbox, buff=None):
    """Ghost fill (sync=False) or additive sync through `buff` (sync=True), direction by
    direction in Directions::all order."""
    if sync:
        if buff is None:
            buff = np.zeros_like(fld)
        buff[...] = 0.0
    for d in range(dom.ndir):
        md = dom.ndir - 1 - d
        if d == dom.centre or not dom.enabled[md]:
            continue
        r = _slab(dom.n, dom.G, dir_vec(dom.dim, d), sync, True)
        data = inbox[(dom.neighbor[md], d)]
        if sync:
            v = _view(buff, r, c0, c1)
            v += data
        else:
            _view(fld, r, c0, c1)[...] = data
    if sync:
        G = dom.G
        act = (slice(c0, c1),) + tuple(slice(G, G + na) for na in dom.n[::-1])
        fld[act] += buff[act]


def exchange_fields_loopback(doms, flds, c0, c1, sync):
    """All domains in one process."""
    lb = Loopback()
    for dom, f in zip(doms, flds):
        lb.post(dom.rank, field_outbox(dom, f, c0, c1, sync))
    for dom, f in zip(doms, flds):
        wanted, _ = field_wanted(dom, c0, c1, sync)
        field_apply(dom, f, c0, c1, sync, lb.collect(dom.rank, wanted))


def exchange_fields_dist(dom, fld, c0, c1, sync):
    wanted, shapes = field_wanted(dom, c0, c1, sync)
    inbox = gloo_exchange(dom.rank, field_outbox(dom, fld, c0, c1, sync), wanted, shapes)
    field_apply(dom, fld, c0, c1, sync, inbox)


# --------------------------------------------------------------------------- particles
INT_NAMES = ["i1", "i2", "i3"]
DX_NAMES = ["dx1", "dx2", "dx3"]


def particle_outbox(dom: Domain, p, npart, nbr_n):
    """Packs the particles carrying a send tag; returns (outbox, holes). nbr_n[d] = active
    extents of the neighbour in direction d (for the index shift). Record layout per particle
    = comm.cu's: [i, i_prev] x D, [dx, dx_prev] x D as float bits, ux1..3, weight."""
    D = dom.dim
    tag = p.tag[:npart]
    holes = list(np.nonzero(tag == 0)[0])
    out = {}
    for d in range(dom.ndir):
        if d == dom.centre:
            continue
        t = 2 + d - (1 if d > dom.centre else 0)
        idx = np.nonzero(tag == t)[0]
        holes.extend(idx)
        if not dom.enabled[d]:
            continue
        dv = dir_vec(D, d)
        rec = np.zeros((idx.size, 4 * D + 4), dtype=np.uint32)
        for a in range(D):
            shift = nbr_n[d][a] if dv[a] == -1 else (-dom.n[a] if dv[a] == 1 else 0)
            rec[:, 2 * a] = (getattr(p, INT_NAMES[a])[idx] + shift).astype(np.int32).view(np.uint32)
            rec[:, 2 * a + 1] = (getattr(p, INT_NAMES[a] + "_prev")[idx] + shift).astype(np.int32).view(np.uint32)
            rec[:, 2 * D + 2 * a] = getattr(p, DX_NAMES[a])[idx].view(np.uint32)
            rec[:, 2 * D + 2 * a + 1] = getattr(p, DX_NAMES[a] + "_prev")[idx].view(np.uint32)
        for k, nm in enumerate(("ux1", "ux2", "ux3", "weight")):
            rec[:, 4 * D + k] = getattr(p, nm)[idx].view(np.uint32)
        out[(dom.neighbor[d], d)] = rec
    p.tag[:npart][tag >= 2] = 0
    return out, np.asarray(holes, dtype=np.int64)


def particle_wanted(dom: Domain):
    wanted = []
    for d in range(dom.ndir):
        md = dom.ndir - 1 - d
        if d == dom.centre or not dom.enabled[md]:
            continue
        wanted.append((dom.neighbor[md], d))
    return wanted


def particle_apply(dom: Domain, p, npart, holes, inbox):
    """ExtractReceivedPrtls: holes first, then append. Returns the new npart."""
    D = dom.dim
    recs = [inbox[k] for k in particle_wanted(dom) if inbox[k].shape[0] > 0]
    if not recs:
        return npart
    rec = np.concatenate(recs, axis=0)
    nrecv = rec.shape[0]
    dest = np.empty(nrecv, dtype=np.int64)
    nh = min(nrecv, holes.size)
    dest[:nh] = holes[:nh]
    dest[nh:] = npart + np.arange(nrecv - nh)
    for a in range(D):
        getattr(p, INT_NAMES[a])[dest] = rec[:, 2 * a].view(np.int32)
        getattr(p, INT_NAMES[a] + "_prev")[dest] = rec[:, 2 * a + 1].view(np.int32)
        getattr(p, DX_NAMES[a])[dest] = rec[:, 2 * D + 2 * a].view(np.float32)
        getattr(p, DX_NAMES[a] + "_prev")[dest] = rec[:, 2 * D + 2 * a + 1].view(np.float32)
    for k, nm in enumerate(("ux1", "ux2", "ux3", "weight")):
        getattr(p, nm)[dest] = rec[:, 4 * D + k].view(np.float32)
    p.tag[dest] = 1
    return npart + max(0, nrecv - holes.size)
