"""GPU parity of the reference's dormant switches and rare branches, through
the C ABI against the oracle (which tests/test_oracle_pin.py pins to the compiled reference on the same cases):

  - dtt (src/main.c:117, :1555-1561): with dtt > 0 the confining right / top walls stay at the lattice extent, and a
    packing leaning on them drives force_WallT / force_WallR (:846-887, :923-951);
  - angleG (:98, :1841-1842): tilted gravity;
  - the order-dependent corner of the grain bounce-back sweep (:1176-1185): short links that face another grain's
    active node across a one-node gap -- the device evaluates them through its deferred list, which must be
    NON-EMPTY in this test.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.oraclewrap import Oracle
from util import mirrored_state, perturbed_f, random_kinematics, small_packing

import lbmdem_gpu as G


def _relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_top_right_walls_dtt_and_tilted_gravity_strict(prec):
    lx, ly = 96, 80
    o = Oracle(lx, ly, 1.0, prec)
    s = G.Solver(lx, ly, 1.0, prec, strict_fp=1, dtt=1.0, angleG=0.35)
    r, x, y = small_packing(lx, ly, 1.0, 71, n_target=60)
    n = o.init_arrays(r, x, y)
    assert s.init_arrays(r, x, y) == n
    o.set_dtt(1.0)
    o.set_angleG(0.35)
    sc = o.scalars()
    assert sc == s.scalars() and sc["xG"] != 0.0
    v, w, a = random_kinematics(n, 72, vmax=0.02)
    st = o.grains()[:, :9].copy()
    st[:, 3:5], st[:, 5:6] = v, w
    st = mirrored_state(st, sc["Mdx"], sc["Mhy"])
    rr = o.grains()[:, 9]
    assert (st[:, 0] + rr > sc["Mdx"]).any() and (st[:, 1] + rr > sc["Mhy"]).any()
    f0 = perturbed_f(lx, ly, 73)
    for z in (o, s):
        z.set_f(f0)
        z.set_grain_state(st)
    for chunk in range(3):
        o.step(41)
        s.step(41)
        assert o.scalars() == s.scalars()
        assert np.array_equal(o.grains()[:, :9], s.grains()[:, :9]), chunk
        assert np.array_equal(o.obst(), s.obst())
    assert np.array_equal(o.f(), s.f()) and np.array_equal(o.fhf(), s.fhf())
    B, T, L, R = o.wall_lists()
    assert len(T) > 0 and len(R) > 0
    wf = s.verlet_full()[2]
    # wall flags of the device: bit per wall (dem_node.cuh wall_flags, order B T L R)
    for bit, lst in enumerate((B, T, L, R)):
        assert np.array_equal(np.nonzero(wf & (1 << bit))[0].astype(np.int32), lst), "BTLR"[bit]
    # the walls stayed at the lattice extent (dtt = 0 would have moved them to 10 x that at the first VerletWall)
    assert o.scalars()["Mdx"] == sc["Mdx"] and o.scalars()["Mhy"] == sc["Mhy"]


def test_top_right_walls_default_build_dem_is_bit_exact():
    """one DEM step from identical input is bit-exact in the default build too (the DEM kernels are never contracted)"""
    lx, ly = 96, 80
    o = Oracle(lx, ly, 1.0, "f64")
    s = G.Solver(lx, ly, 1.0, "f64", dtt=1.0, angleG=-0.2)
    r, x, y = small_packing(lx, ly, 1.0, 75, n_target=60)
    n = o.init_arrays(r, x, y)
    assert s.init_arrays(r, x, y) == n
    o.set_dtt(1.0)
    o.set_angleG(-0.2)
    sc = o.scalars()
    v, w, a = random_kinematics(n, 76, vmax=0.05)
    st = o.grains()[:, :9].copy()
    st[:, 3:5], st[:, 5:6] = v, w
    st = mirrored_state(st, sc["Mdx"], sc["Mhy"])
    for z in (o, s):
        z.set_grain_state(st)
    o.phase("init_verlet")
    s.build_verlet()
    npd = sc["npDEM"]
    o.set_nbsteps(1)
    s.set_nbsteps(1)               # nbsteps % npDEM != 0: DEM sub-steps only, fhf stays zero
    o.step(npd - 1)
    s.step(npd - 1)
    assert np.array_equal(o.grains()[:, :9], s.grains()[:, :9])
    assert np.abs(o.grains()[:, 6:8]).max() > 10 * 9.81    # wall / pair contacts dominate gravity


def _gap_pairs(dx, rlb=4.6):
    """pairs of equal discs whose reduced discs (rLB = 0.85 r / dx = 4.6 nodes) leave exactly ONE fluid node between
    them along a lattice link, with the link fraction delta below 1/2 on one or both sides"""
    r = rlb * dx / 0.85
    cx, cy, rs = [], [], []
    for k, d in enumerate((9.7, 9.9, 10.3)):             # along x
        x0, y0 = 20.0, 14.0 + 14 * k
        cx += [x0, x0 + d]; cy += [y0, y0]; rs += [r, r]
    for k, d in enumerate((9.7, 10.3)):                  # along y
        x0, y0 = 50.0 + 16 * k, 14.0
        cx += [x0, x0]; cy += [y0, y0 + d]; rs += [r, r]
    for k, d in enumerate((6.9, 7.3)):                   # along the diagonal
        x0, y0 = 50.0 + 16 * k, 40.0
        cx += [x0, x0 + d]; cy += [y0, y0 + d]; rs += [r, r]
    return np.array(rs), np.array(cx) * dx, np.array(cy) * dx


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_one_node_gap_links_go_through_the_deferred_list_strict(prec):
    lx, ly = 96, 72
    o = Oracle(lx, ly, 1.0, prec)
    dx = o_dx = (1e-3 * lx / 10) / (lx - 1)
    r, x, y = _gap_pairs(dx)
    s = G.Solver(lx, ly, 1.0, prec, strict_fp=1)
    n = o.init_arrays(r, x, y)
    assert s.init_arrays(r, x, y) == n
    v, w, a = random_kinematics(n, 82, vmax=0.01)
    st = o.grains()[:, :9].copy()
    st[:, 3:5], st[:, 5:6] = v, w
    f0 = perturbed_f(lx, ly, 83)
    for z in (o, s):
        z.set_f(f0)
        z.set_grain_state(st)
    deferred = 0
    for k in range(6):
        o.lbm_step()
        s.lbm_step()
        deferred += s.list_counts()["deferred"]
        assert np.array_equal(o.obst(), s.obst())
        assert np.array_equal(o.f(), s.f()), k
        assert np.array_equal(o.fhf(), s.fhf()), k
    assert deferred > 0, "no link was evaluated through the deferred list: the test geometry lost its one-node gaps"
    # and in the coupled run (grains move)
    o.step(25)
    s.step(25)
    assert np.array_equal(o.f(), s.f()) and np.array_equal(o.grains()[:, :9], s.grains()[:, :9])


def test_one_node_gap_links_default_build():
    lx, ly = 96, 72
    o = Oracle(lx, ly, 1.0, "f64")
    dx = (1e-3 * lx / 10) / (lx - 1)
    r, x, y = _gap_pairs(dx)
    s = G.Solver(lx, ly, 1.0, "f64")
    n = o.init_arrays(r, x, y)
    assert s.init_arrays(r, x, y) == n
    f0 = perturbed_f(lx, ly, 84)
    for z in (o, s):
        z.set_f(f0)
    total = 0
    for k in range(4):
        o.lbm_step()
        s.lbm_step()
        total += s.list_counts()["deferred"]
    assert total > 0
    assert np.array_equal(o.obst(), s.obst())
    assert np.abs(o.f() - s.f()).max() < 1e-12
    assert _relerr(s.fhf(), o.fhf()) < 1e-9


def test_diverged_populations_raise_the_range_flag():
    """the 64-bit fixed-point force sums assume |sum| < 2^11: populations of 1e6 next to a grain must not wrap silently"""
    lx, ly = 96, 72
    r, x, y = small_packing(lx, ly, 1.0, 5, n_target=30)
    s = G.Solver(lx, ly, 1.0, "f64")
    s.init_arrays(r, x, y)
    s.set_f(perturbed_f(lx, ly, 6) * 1e6)
    with pytest.raises(G.LbmdemError) as ei:
        s.step(3)
    assert ei.value.code == -8
