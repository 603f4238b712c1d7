"""Host-side mirror of the reference's Metadomain for one domain per GPU
(src/framework/domain/metadomain.cpp:101-330): decomposition of the global mesh, this rank's
block, its neighbours and local boundary conditions. The tables come from the library's own
host code (``eb200_decompose`` / ``eb200_domain_info``); the exchange itself (pack kernels,
grouped ncclSend/ncclRecv, unpack kernels) lives in ``csrc/comm.cu``. torch.distributed is only
used to hand the ncclUniqueId from rank 0 to the other ranks (the reference bootstraps through
MPI_Init)."""
from __future__ import annotations

from . import lib as L


class Metadomain:
    def __init__(self, global_n, nranks=1, rank=0, decomposition=None, fbc=None, pbc=None):
        self.dim = len(global_n)
        self.global_n = tuple(global_n)
        self.nranks, self.rank = nranks, rank
        self.extents = L.decompose(nranks, list(global_n), decomposition)
        self.ndoms = tuple(len(e) for e in self.extents)
        self.md = L.make_metadomain(rank, self.extents, fbc, pbc)
        self.info = L.domain_info(self.md)
        self.local_n = tuple(self.info.n[a] for a in range(self.dim))
        self.offset = tuple(self.info.offset[a] for a in range(self.dim))
        self.cell_offset = tuple(self.info.cell_offset[a] for a in range(self.dim))
        self.face_fbc = [self.info.face_fbc[k] for k in range(6)]
        self.face_pbc = [self.info.face_pbc[k] for k in range(6)]

    def neighbor(self, direction):
        """Rank of the neighbour in `direction` (tuple of -1/0/1 per dimension)."""
        lin = 0
        for a in range(self.dim):
            lin = lin * 3 + (direction[a] + 1)
        return self.info.neighbor[lin]

    def attach(self, sim, uid: bytes | None):
        """Bind this decomposition (and an NCCL communicator when `uid` is given) to a
        Simulation whose grid is this rank's block; installs the local boundary conditions."""
        import ctypes as C
        assert tuple(sim.grid.n[a] for a in range(self.dim)) == self.local_n
        sim.ctx.comm_init(self.md, uid)
        sim.params.fbc = (C.c_int * 6)(*self.face_fbc)
        sim.params.pbc = (C.c_int * 6)(*self.face_pbc)
        sim.metadomain = self
        return sim


def bootstrap_unique_id(device=None) -> bytes | None:
    """ncclGetUniqueId on rank 0, handed to every rank through torch.distributed (whatever
    backend the process group uses). Returns None for a single process."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None
    on_gpu = dist.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    t = torch.zeros(L.UNIQUE_ID_BYTES, dtype=torch.uint8, device=dev)
    if dist.get_rank() == 0:
        raw = L.unique_id()
        t.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())
