"""Worker of tests/test_strips_gloo.py: one of WORLD_SIZE CPU processes (gloo).  Each owns an
x-strip of the lattice and runs the strip-staged host check of the product's node headers
(tests/hostcheck: hc_strip_stage1/2/3), exchanging ghost rows and force sums the way
csrc/sim.cu does over NCCL.  Rank 0 also runs the undivided lattice and everybody compares."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "2d-lbm-dem_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch
import torch.distributed as dist

import lbmdem_dist as D
from oracle.oraclewrap import Oracle
from util import load_hostcheck, perturbed_f, random_kinematics, small_packing

GHOST = 4
dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def bind(hc):
    hc.hc_strip_stage1_f64.argtypes = [C.c_int] * 7 + [dp, dp, dp, ip, ip]
    hc.hc_strip_stage2_f64.argtypes = [C.c_int] * 8 + [dp, dp, dp, lp]
    hc.hc_strip_stage3_f64.argtypes = [C.c_int] * 6 + [dp, dp, dp]


def to_local(a_xyq, x0, nxl):
    """reference layout [x][y][q] (global) -> local planes [q][row][y]; rows outside the lattice are zero"""
    lx, ly, _ = a_xyq.shape
    out = np.zeros((9, nxl, ly))
    for r in range(nxl):
        x = x0 + r
        if 0 <= x < lx:
            out[:, r, :] = a_xyq[x].T
    return out


def local_map(obst, x0, nxl, ring):
    lx, ly = obst.shape
    out = np.full((nxl, ly), ring, dtype=np.int32)
    for r in range(nxl):
        x = x0 + r
        if 0 <= x < lx:
            out[r] = obst[x]
    return out


def peer_rule_sum(parts, gtab, scal, lx, n, world):
    """ipc_sum_kernel's rule on the CPU: grain i takes the partial sums of rank k iff its clamped bounding box
    (src/main.c:1009-1023, csrc/raster_node.cuh grain_geometry), widened by one row, meets rank k's rows."""
    dx, mgx = scal[0], scal[2]
    xc, rbl0 = (gtab[:, 0] - mgx) / dx, gtab[:, 5] / dx
    xi = np.maximum(np.trunc(xc - rbl0).astype(np.int64), 1)
    xf = np.minimum(np.trunc(xc + rbl0).astype(np.int64), lx - 2)
    out = np.zeros(3 * n, dtype=np.int64)
    used = 0
    for k in range(world):
        klo, khi = D.strip_bounds(lx, k, world)
        take = (xf + 1 >= klo) & (xi - 1 < khi) & (xf >= xi)
        used += int(take.sum())
        for c in range(3):
            out[c * n:(c + 1) * n] += np.where(take, parts[k][c * n:(c + 1) * n], 0)
    assert used < world * n, "the rule selects every rank for every grain: nothing is tested"
    return out


def run_strips(hc, lx, ly, n, scal, gtab, f_in, obst_old, rank, world):
    xlo, xhi = D.strip_bounds(lx, rank, world)
    multi = world > 1
    x0, nxl = (xlo - GHOST, xhi - xlo + 2 * GHOST) if multi else (0, lx)
    f = np.ascontiguousarray(to_local(f_in, x0, nxl))
    cell_old = np.ascontiguousarray(local_map(obst_old, x0, nxl, n))
    cell_new = np.empty_like(cell_old)
    assert hc.hc_strip_stage1_f64(lx, ly, n, x0, nxl, xlo, xhi, scal, gtab, f, cell_old, cell_new) == 0
    if multi:
        # ghost exchange: GHOST owned rows per side, all nine planes (Sim::halo_exchange)
        lo_send = torch.from_numpy(f[:, GHOST:2 * GHOST, :].copy())
        hi_send = torch.from_numpy(f[:, nxl - 2 * GHOST:nxl - GHOST, :].copy())
        ops, bufs = [], {}
        if rank > 0:
            bufs["lo"] = torch.empty_like(lo_send)
            ops += [dist.P2POp(dist.isend, lo_send, rank - 1), dist.P2POp(dist.irecv, bufs["lo"], rank - 1)]
        if rank < world - 1:
            bufs["hi"] = torch.empty_like(hi_send)
            ops += [dist.P2POp(dist.isend, hi_send, rank + 1), dist.P2POp(dist.irecv, bufs["hi"], rank + 1)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        if "lo" in bufs:
            f[:, 0:GHOST, :] = bufs["lo"].numpy()
        if "hi" in bufs:
            f[:, nxl - GHOST:nxl, :] = bufs["hi"].numpy()
    facc = np.zeros(3 * n, dtype=np.int64)
    assert hc.hc_strip_stage2_f64(lx, ly, n, x0, nxl, xlo, xhi, world, scal, gtab, f, facc) == 0
    if multi:
        t = torch.from_numpy(facc)
        # the default transport of a multi-process GPU run does not all-reduce: a rank reads, per grain, the partial sums of
        # the ranks whose rows the grain's bounding box touches (ipc_sum_kernel, csrc/aux_kernels.cu).  Same rule here,
        # from all-gathered partial sums; it must give the all-reduce's integers.
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.clone())
        by_rule = peer_rule_sum([p.numpy() for p in parts], gtab, scal, lx, n, world)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)   # integer sum: exact
        assert np.array_equal(by_rule, facc), f"rank {rank}: the peer-memory selection rule drops a contribution"
    f_out = np.empty((xhi - xlo, ly, 9))
    assert hc.hc_strip_stage3_f64(lx, ly, x0, nxl, xlo, xhi, scal, f, f_out) == 0
    return xlo, xhi, f_out, facc, cell_new[(xlo - x0):(xhi - x0)]


def main():
    rank, world = D.init_process_group("gloo")
    assert world == int(os.environ["WORLD_SIZE"]) and world >= 2
    hc = load_hostcheck()
    bind(hc)
    # bootstrap helpers the GPU path relies on
    token = os.urandom(128) if rank == 0 else None
    got = D.broadcast_bytes(token, 128, src=0)
    assert len(got) == 128 and (rank != 0 or got == token)
    assert D.max_over_ranks(float(rank)) == float(world - 1)
    D.barrier()

    lx, ly = 61, 48            # odd split; grains straddle the strip boundary
    o = Oracle(lx, ly, 1.0, "f64")
    r, x, y = small_packing(lx, ly, 1.0, seed=17, n_target=40)
    n = o.init_arrays(r, x, y)
    o.set_lid(0.02)
    o.set_f(perturbed_f(lx, ly, 18))
    v, w, a = random_kinematics(n, 19)
    rng = np.random.default_rng(20)
    sc = o.scalars()
    scal = np.array([sc["dx"], sc["c"], sc["Mgx"], sc["Mby"], 0.02])
    for step in range(3):
        st = o.grains()[:, :9].copy()
        st[:, 0:2] += rng.uniform(-0.6, 0.6, size=(n, 2)) * sc["dx"]
        st[:, 3:5], st[:, 5:6] = v * (1 + 0.2 * step), w
        o.set_grain_state(st)
        g = o.grains()
        gtab = np.ascontiguousarray(g[:, [0, 1, 3, 4, 5, 9, 12]])
        f_in, obst_old = o.f(), o.obst()
        o.lbm_step()
        f_ref, obst_ref, act_ref = o.f(), o.obst(), o.act()
        xlo, xhi, f_out, facc, cell_new = run_strips(hc, lx, ly, n, scal, gtab, f_in, obst_old, rank, world)
        # the strip reproduces the oracle's rows bit for bit
        assert np.array_equal(f_out, f_ref[xlo:xhi]), f"rank {rank} step {step}: populations differ from the oracle"
        own = np.where(cell_new >= 0, cell_new & ((1 << 29) - 1), -1)  # drop CELL_ACT, CELL_RIM
        assert np.array_equal(own, obst_ref[xlo:xhi]), f"rank {rank} step {step}: obstacle map"
        solid = (obst_ref[xlo:xhi] >= 0) & (obst_ref[xlo:xhi] < n)
        assert np.array_equal(((cell_new >> 30) & 1)[solid], act_ref[xlo:xhi][solid]), f"rank {rank} step {step}: act"
        # force sums: identical integers with one strip and with `world` strips
        _, _, f_one, facc_one, _ = run_strips(hc, lx, ly, n, scal, gtab, f_in, obst_old, 0, 1)
        assert np.array_equal(facc, facc_one), f"rank {rank} step {step}: fixed-point force sums depend on the decomposition"
        assert np.array_equal(f_one, f_ref)
        # and they are the oracle's forces up to the fixed-point resolution
        k = 1000.0 * 9 * 1e-6 * 1e-6 / (sc["dx"] * (0.504 - 0.5) ** 2)
        fh = facc.reshape(3, n).T.astype(np.float64) / np.array([2.0 ** 52, 2.0 ** 52, 2.0 ** 48]) * np.array([k, k, k * sc["dx"]])
        ref = o.fhf()
        assert np.abs(fh - ref).max() <= 1e-9 * max(np.abs(ref).max(), 1e-30), f"rank {rank} step {step}: fhf"
    D.barrier()
    dist.destroy_process_group()
    print(f"rank {rank}: ok")


if __name__ == "__main__":
    main()
