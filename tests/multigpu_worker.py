"""Worker of tests/test_multi_gpu.py: one process per GPU (torchrun, NCCL).  The strip-decomposed
run must be BIT-IDENTICAL to the one-GPU run of the same build (pure data movement plus an exact
integer all-reduce), and agree with the oracle like the one-GPU run does."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "2d-lbm-dem_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch

import lbmdem_dist as D
import lbmdem_gpu as G
from oracle.oraclewrap import Oracle
from util import perturbed_f, random_kinematics, small_packing


def main():
    rank, world = D.init_process_group("nccl")
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    for prec, lx, ly, steps in (("f64", 203, 160, 60), ("f32", 160, 131, 24)):
        r, x, y = small_packing(lx, ly, 1.0, seed=71, n_target=90)
        if prec == "f32":  # one grain fewer: an odd and an even count between the two runs, so that with two ranks the
            r, x, y = r[:-1], x[:-1], y[:-1]  # share form goes through both transports (all-gather / broadcasts)
        f0 = perturbed_f(lx, ly, 72)
        o = Oracle(lx, ly, 1.0, prec)
        n = o.init_arrays(r, x, y)
        v, w, a = random_kinematics(n, 73, vmax=0.02)
        st = o.grains()[:, :9].copy()
        st[:, 3:5], st[:, 5:6] = v, w
        o.set_f(f0)
        o.set_grain_state(st)

        strip = D.make_strip_solver(lx, ly, 1.0, prec)
        one = G.Solver(lx, ly, 1.0, prec, device=local_rank)
        xlo, xhi = strip.xlo, strip.xhi
        assert (xlo, xhi) == D.strip_bounds(lx, rank, world)
        for s in (strip, one):
            assert s.init_arrays(r, x, y) == n
            s.set_grain_state(st)
        one.set_f(f0)
        strip.set_f(f0[xlo:xhi])
        for chunk in range(3):
            for s in (strip, one, o):
                s.step(steps // 3)
            assert np.array_equal(strip.f(), one.f()[xlo:xhi]), f"{prec} rank {rank}: populations differ from the 1-GPU run"
            assert np.array_equal(strip.obst(), one.obst()[xlo:xhi])
            assert np.array_equal(strip.fhf(), one.fhf()), f"{prec} rank {rank}: fhf differs from the 1-GPU run"
            assert np.array_equal(strip.grains(), one.grains()), f"{prec} rank {rank}: grains differ from the 1-GPU run"
        assert np.array_equal(strip.obst(), o.obst()[xlo:xhi])
        tol = 1e-6 if prec == "f64" else 1e-4
        go, gs = o.grains(), strip.grains()
        assert np.abs(gs[:, :3] - go[:, :3]).max() <= tol * np.abs(go[:, :3]).max()
        # the density checksum of the strips adds up to the 1-GPU one
        part = torch.tensor([strip.total_density()], dtype=torch.float64, device="cuda")
        torch.distributed.all_reduce(part)
        assert abs(part.item() - one.total_density()) < 1e-9 * lx * ly
        # end-to-end call in its share form: every rank uploads / downloads its share of the grain rows, the ranks pass
        # them on over NCCL; same result as the one-GPU call with all rows
        npd = one.scalars()["npDEM"]
        i0, i1 = strip.share()
        assert (i0, i1) == D.strip_bounds(n, rank, world)
        rows = "f32" if prec == "f32" else "f64"
        full = one.grains()[:, :9].copy()
        for _ in range(2):
            so, fo, _d = one.step_host(full, npd, rows=rows)
            ss, fs, _d = strip.step_host(full[i0:i1], npd, rows=rows, share=True)
            assert np.array_equal(ss, so[i0:i1]) and np.array_equal(fs, fo[i0:i1]), f"{prec} rank {rank}: share form differs"
            full = np.array(so, dtype=np.float64)
        assert np.array_equal(strip.grains(), one.grains())
        # checkpoint / restart of a decomposed run: one file per rank, bit-exact continuation
        import tempfile
        ck = os.path.join(tempfile.gettempdir(), f"lbmdem_ck_{prec}_{os.environ.get('MASTER_PORT', '0')}")
        strip.save_state(ck)
        strip.step(23)
        again = D.make_strip_solver(lx, ly, 1.0, prec)
        assert again.load_state(ck) == n
        again.step(23)
        assert np.array_equal(again.f(), strip.f()) and np.array_equal(again.grains(), strip.grains())
        assert np.array_equal(again.fhf(), strip.fhf())
        again.close()
        strip.close()
        one.close()
    D.barrier()
    torch.distributed.destroy_process_group()
    print(f"rank {rank}: ok")


if __name__ == "__main__":
    main()
