"""Strip decomposition on ONE GPU: the ranks of a decomposed run as contexts of one process
(lbmdem_local_group_*), all on device 0, each driven by its own host thread.  Everything the multi-GPU path does --
ghost rows, the staged ring / bounce-back sweeps around the exchange, recomputed neighbour rims, the exact integer
sum of the force contributions -- runs here; only the transport differs (peer copies instead of NCCL).  The result
must be BIT-IDENTICAL to the undivided lattice for any number of strips."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from util import perturbed_f, random_kinematics, small_packing

import lbmdem_dist as D
import lbmdem_gpu as G


def _start(lx, ly, prec, seed, n_target):
    r, x, y = small_packing(lx, ly, 1.0, seed=seed, n_target=n_target)
    one = G.Solver(lx, ly, 1.0, prec)
    n = one.init_arrays(r, x, y)
    v, w, a = random_kinematics(n, seed + 2, vmax=0.02)
    st = one.grains()[:, :9].copy()
    st[:, 3:5], st[:, 5:6] = v, w
    f0 = perturbed_f(lx, ly, seed + 1)
    one.set_f(f0)
    one.set_grain_state(st)
    return one, (r, x, y), st, f0, n


@pytest.mark.parametrize("prec,lx,ly,nranks", [("f64", 203, 160, 2), ("f64", 203, 160, 3), ("f32", 160, 131, 4),
                                               ("f64", 97, 64, 7)])
def test_strips_on_one_device_equal_the_undivided_lattice(prec, lx, ly, nranks):
    one, (r, x, y), st, f0, n = _start(lx, ly, prec, 71, 90)
    grp = D.LocalStrips(lx, ly, 1.0, prec, nranks)
    assert grp.init_arrays(r, x, y) == n
    grp.set_f(f0)
    grp.set_grain_state(st)
    assert [(s.xlo, s.xhi) for s in grp.ranks] == [D.strip_bounds(lx, k, nranks) for k in range(nranks)]
    for chunk in range(3):
        one.step(23)
        grp.step(23)
        assert np.array_equal(grp.obst(), one.obst()), chunk
        assert np.array_equal(grp.f(), one.f()), f"populations differ from the one-context run after {(chunk + 1) * 23} calls"
        for s in grp.ranks:   # grains are replicated: every rank holds the one-context run's bits
            assert np.array_equal(s.grains(), one.grains()), (chunk, s.params.rank)
            assert np.array_equal(s.fhf(), one.fhf()), (chunk, s.params.rank)
        assert grp.state_checksum() == one.state_checksum()
    assert abs(grp.total_density() - one.total_density()) < 1e-9 * lx * ly
    grp.close()
    one.close()


def test_strips_on_one_device_real_sample_4096():
    """the benchmarked configuration (BASELINE configs[3]: bin/a08_a4b4r18_7000.data, 4096 x 4096, scale 2.7, fp32) in
    4 strips against the undivided lattice, through the exact fingerprint"""
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "a08_a4b4r18_7000.data")
    one = G.Solver(4096, 4096, 2.7, "f32")
    n = one.init(path)
    grp = D.LocalStrips(4096, 4096, 2.7, "f32", 4)
    assert grp.init(path) == n
    for z in (one, grp):
        z.step(12)
    assert grp.state_checksum() == one.state_checksum()
    for s in grp.ranks:
        assert np.array_equal(s.grains(), one.grains()) and np.array_equal(s.fhf(), one.fhf())
    grp.close()
    one.close()


def test_a_failing_rank_releases_its_peers():
    """neighbour capacity 1 makes the Verlet build fail on every rank: the call returns an error instead of hanging"""
    lx, ly = 120, 96
    r, x, y = small_packing(lx, ly, 1.0, seed=5, n_target=60)
    grp = D.LocalStrips(lx, ly, 1.0, "f64", 2, neighbour_capacity=1)
    grp.init_arrays(r, x, y)
    with pytest.raises(G.LbmdemError):
        grp.step(3)
    grp.close()


def test_strips_on_two_devices_equal_the_undivided_lattice():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    lx, ly, prec = 203, 160, "f64"
    one, (r, x, y), st, f0, n = _start(lx, ly, prec, 81, 90)
    grp = D.LocalStrips(lx, ly, 1.0, prec, 4, devices=[0, 1, 0, 1])
    assert grp.init_arrays(r, x, y) == n
    grp.set_f(f0)
    grp.set_grain_state(st)
    for chunk in range(2):
        one.step(31)
        grp.step(31)
        assert np.array_equal(grp.f(), one.f()) and np.array_equal(grp.obst(), one.obst())
        assert np.array_equal(grp.ranks[3].grains(), one.grains()) and np.array_equal(grp.ranks[1].fhf(), one.fhf())
    grp.close()
    one.close()
