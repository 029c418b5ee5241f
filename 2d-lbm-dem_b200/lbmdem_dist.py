"""Multi-GPU plumbing around liblbmdem_gpu.so: one process per GPU (torchrun), the lattice
split into x strips, grains replicated.  torch.distributed is only the bootstrap here -- it
carries the 128-byte NCCL unique id from rank 0 to the others and the max-over-ranks of the
timings; the halo rows and the force sums travel through the library's own NCCL communicator
(lbmdem_attach_nccl), inside the step.
"""
from __future__ import annotations

import os


def strip_bounds(lx: int, rank: int, nranks: int):
    """Rows [xlo, xhi) owned by `rank`: the same split as Sim::init_device (csrc/sim.cu)."""
    base, rem = divmod(lx, nranks)
    xlo = rank * base + min(rank, rem)
    return xlo, xlo + base + (1 if rank < rem else 0)


def env_rank():
    """(rank, local_rank, world_size) from the torchrun environment (1 process if unset)."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)))


def init_process_group(backend: str | None = None):
    """Joins the torch.distributed group described by the environment; returns (rank, world)."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_rank()
    if world == 1:
        return 0, 1
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0) -> bytes:
    """Rank `src` passes `payload` (nbytes long); everybody gets it back."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return payload
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    if dist.get_rank() == src:
        t = torch.tensor(list(payload), dtype=torch.uint8, device=dev)
    else:
        t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().tolist())


def max_over_ranks(value: float) -> float:
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def make_strip_solver(lx, ly, scale, prec, **over):
    """Solver for this process's strip, with the library's NCCL communicator attached."""
    import lbmdem_gpu as G
    rank, local_rank, world = env_rank()
    s = G.Solver(lx, ly, scale, prec, device=local_rank, rank=rank, nranks=world, **over)
    if world > 1:
        uid = G.nccl_unique_id() if rank == 0 else None
        uid = broadcast_bytes(uid, 128, src=0)
        s.attach_nccl(uid)
    return s


class LocalStrips:
    """The strips of a decomposed run inside ONE process (lbmdem_local_group_*): one context per strip, each driven by
    its own host thread, on the given devices -- all on device 0 by default, which exercises the whole strip logic
    (ghost rows, staged sweeps, exact integer force sums) on a one-GPU machine.  Mirrors lbmdem_gpu.Solver."""

    def __init__(self, lx, ly, scale, prec, nranks, devices=None, **over):
        import ctypes as C
        from concurrent.futures import ThreadPoolExecutor

        import lbmdem_gpu as G
        self.G, self.nranks = G, nranks
        devices = list(devices) if devices is not None else [0] * nranks
        assert len(devices) == nranks
        L = G.load_library()
        grp = C.c_void_p()
        rc = L.lbmdem_local_group_create(nranks, C.byref(grp))
        if rc:
            raise G.LbmdemError(rc, "lbmdem_local_group_create")
        self._L, self.group = L, grp
        self.ranks = [G.Solver(lx, ly, scale, prec, device=devices[k], rank=k, nranks=nranks, **over) for k in range(nranks)]
        for s in self.ranks:
            s.attach_local(grp)
        self.pool = ThreadPoolExecutor(max_workers=nranks)
        self.lx, self.ly = lx, ly

    def _all(self, fn):
        """the same call on every rank, concurrently (the ranks meet at barriers inside the step)"""
        futs = [self.pool.submit(fn, s) for s in self.ranks]
        errs, out = [], []
        for f in futs:
            try:
                out.append(f.result())
            except Exception as e:  # noqa: BLE001 - collect every rank's outcome before raising
                errs.append(e)
        if errs:
            raise errs[0]
        return out

    def init(self, path):
        return self._all(lambda s: s.init(path))[0]

    def init_arrays(self, r, x, y):
        return self._all(lambda s: s.init_arrays(r, x, y))[0]

    def step(self, n=1):
        self._all(lambda s: s.step(n))

    def lbm_step(self):
        self._all(lambda s: s.lbm_step())

    def set_f(self, f):
        for s in self.ranks:
            s.set_f(f[s.xlo:s.xhi])

    def set_grain_state(self, st):
        for s in self.ranks:
            s.set_grain_state(st)

    def f(self):
        import numpy as np
        return np.concatenate([s.f() for s in self.ranks], axis=0)

    def obst(self):
        import numpy as np
        return np.concatenate([s.obst() for s in self.ranks], axis=0)

    def state_checksum(self):
        tot = [0, 0]
        for s in self.ranks:
            a, b = s.state_checksum()
            tot[0] = (tot[0] + a) & ((1 << 64) - 1)
            tot[1] = (tot[1] + b) & ((1 << 64) - 1)
        return tuple(tot)

    def total_density(self):
        return sum(s.total_density() for s in self.ranks)

    def close(self):
        for s in self.ranks:
            s.close()
        self.ranks = []
        self.pool.shutdown()
        if self.group:
            self._L.lbmdem_local_group_destroy(self.group)
            self.group = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
