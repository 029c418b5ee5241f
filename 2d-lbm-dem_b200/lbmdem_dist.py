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
