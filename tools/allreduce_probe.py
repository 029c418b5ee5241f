"""tools/allreduce_probe.py -- how long the per-step collective of a strip-decomposed run takes by itself: ncclAllReduce
(sum) of 3 n int64 (the fixed-point force sums of n grains), back to back on one stream, CUDA-event timed, max over ranks.

    NCCL_PROTO=LL128 python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/allreduce_probe.py [n]
"""
import os
import sys

import torch
import torch.distributed as dist


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50840
    dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    out = []
    for elems in (3 * n, 3 * n // 8, 1024):
        x = torch.ones(elems, dtype=torch.int64, device="cuda")
        for _ in range(30):
            dist.all_reduce(x)
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(300):
            dist.all_reduce(x)
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / 300 * 1e3], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out.append("%d x int64: %.1f us" % (elems, t.item()))
    if rank == 0:
        print("world %d  NCCL_ALGO=%s NCCL_PROTO=%s :  %s" % (world, os.environ.get("NCCL_ALGO", "-"), os.environ.get("NCCL_PROTO", "-"), " | ".join(out)), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
