"""CPU pin of the FORMULATION the CUDA kernels implement (no GPU needed).

The product's node headers (csrc/lbm_node.cuh, raster_node.cuh, dem_node.cuh) are compiled by
g++ (tests/hostcheck) and driven serially; results must be BIT-IDENTICAL to the oracle
(oracle/lbmdem_oracle.c, itself pinned bit-identical to the compiled reference):
  - the stored-state restatement of the LBM step: node-local reinit + collide (+ w-links), ring
    sweep in two in-place passes, grain bounce-back sweep in arbitrary link order with the
    deferred list, plain pull incl. ring nodes, solid nodes, moving grains;
  - the max-owner rasteriser with the act rule (src/main.c:991-1065);
  - forces_fluid from post-stream values (src/main.c:1285-1333);
  - the gather-form DEM step over a sorted full neighbour list (src/main.c:1336-1516, :1733-1763).
"""
import numpy as np
import pytest

from oracle.oraclewrap import Oracle
from util import DEM_CONST, scale_fhf, load_hostcheck, perturbed_f, random_kinematics, small_packing


def _grain_table(g):
    # oracle grains(): x1 x2 x3 v1 v2 v3 a1 a2 a3 r m It rLB -> hostcheck: x1 x2 v1 v2 v3 r rLB
    return np.ascontiguousarray(g[:, [0, 1, 3, 4, 5, 9, 12]])


def _lbm_case(prec, lx, ly, scale, seed, lid=0.0, steps=4, n_target=None, discs=None):
    hc = load_hostcheck()
    fn = getattr(hc, f"hc_lbm_step_{prec}")
    real = np.float64 if prec == "f64" else np.float32
    o = Oracle(lx, ly, scale, prec)
    r, x, y = discs if discs is not None else small_packing(lx, ly, scale, seed, n_target=n_target)
    n = o.init_arrays(r, x, y)
    if lid:
        o.set_lid(lid)
    o.set_f(perturbed_f(lx, ly, seed + 1))
    v, w, a = random_kinematics(n, seed + 2)
    rng = np.random.default_rng(seed + 3)
    sc = o.scalars()
    for step in range(steps):
        g = o.grains()
        st = g[:, :9].copy()
        # move the grains by a fraction of a lattice spacing so that nodes change state
        st[:, 0:2] += rng.uniform(-0.6, 0.6, size=(n, 2)) * sc["dx"]
        st[:, 3:5] = v * (1 + 0.1 * step)
        st[:, 5:6] = w
        o.set_grain_state(st)
        g = o.grains()
        f_in, obst_old = o.f(), o.obst()
        o.lbm_step()
        f_ref, obst_ref, act_ref, fhf_ref = o.f(), o.obst(), o.act(), o.fhf()
        scal = np.array([sc["dx"], sc["c"], sc["Mgx"], sc["Mby"], lid], dtype=np.float64)
        f_out = np.empty_like(f_in)
        obst_new = np.empty_like(obst_old)
        act_new = np.empty_like(obst_old)
        fhf = np.empty((n, 3))
        assert fn(lx, ly, n, scal, _grain_table(g), f_in, obst_old, f_out, obst_new, act_new, fhf) == 0
        assert np.array_equal(obst_new, obst_ref), f"obst differs at step {step}"
        solid = (obst_ref >= 0) & (obst_ref < n)
        assert solid.sum() > 0
        assert np.array_equal(act_new[solid], act_ref[solid]), f"act differs at step {step}"
        bad = np.argwhere(f_out != f_ref)
        assert bad.size == 0, f"step {step}: {len(bad)} populations differ, first {bad[:5]}"
        fh = scale_fhf(fhf, sc["dx"], prec)
        assert np.array_equal(fh, fhf_ref), f"fhf differs at step {step}"
    # the case must actually have exercised the boundary machinery
    return dict(n=n, solid=int(solid.sum()), changed=int((obst_old != obst_ref).sum()))


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_lbm_step_formulation_bit_exact_small(prec):
    info = _lbm_case(prec, 64, 48, 1.0, seed=10)
    assert info["n"] >= 6 and info["changed"] > 0


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_lbm_step_formulation_with_lid(prec):
    _lbm_case(prec, 40, 56, 1.0, seed=20, lid=0.05, steps=3)


def test_lbm_step_act_derived_on_demand():
    # the device kernels get a bare obstacle map and derive act[][] themselves (node_act)
    hc = load_hostcheck()
    hc.hc_set_act_folded(0)
    try:
        _lbm_case("f64", 64, 48, 1.0, seed=11, steps=2)
        _lbm_case("f64", 96, 80, 1.0, seed=31, steps=2, n_target=120)
    finally:
        hc.hc_set_act_folded(1)


def test_lbm_step_formulation_dense_scaled():
    # many small grains: one-node gaps between reduced discs, short links reading ring / solid nodes
    info = _lbm_case("f64", 96, 80, 1.0, seed=30, steps=3, n_target=120)
    assert info["n"] >= 60
    # links across one-node gaps went through the deferred list (evaluated in reverse sweep order)
    assert load_hostcheck().hc_last_deferred() > 100


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_lbm_step_overlapping_discs_and_discs_on_the_ring(prec):
    # three mutually overlapping reduced discs (the act rule needs the lowest covering index there),
    # discs reaching into the wall ring, a disc partly outside the lattice
    lx, ly = 96, 72
    dx = 1e-4 * lx / (lx - 1)
    r = np.array([10, 9, 8, 7, 9, 6.5]) * dx
    x = np.array([30, 39, 34, 3.0, 80, 95.0]) * dx
    y = np.array([30, 31, 38, 40, 2.5, 20]) * dx
    info = _lbm_case(prec, lx, ly, 1.0, seed=40, steps=3, discs=(r, x, y))
    assert info["n"] == 6 and info["solid"] > 500
    # the populations the fused kernel leaves unwritten were poisoned (NaN) during the sweeps and the force sums,
    # then materialised for the comparison with the oracle above: nothing on the path read them
    assert load_hostcheck().hc_last_dead() > 300


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_tile_rasteriser_formulation(prec):
    """The tile form of K2 (grains binned by 32 x 64 tiles with a halo, shared-memory painting in arbitrary order,
    act / rim bits from the tile's own neighbourhood) restated serially: the same map, bits and link counts as the
    whole-lattice rasteriser -- on one GPU's lattice and on the local rows of strips (ghost rows included)."""
    hc = load_hostcheck()
    fn = getattr(hc, f"hc_raster_tiles_check_{prec}")
    cases = []
    lx, ly = 200, 333   # several tiles each way, ly not a multiple of the tile width
    r, x, y = small_packing(lx, ly, 1.0, seed=5, n_target=300)
    cases.append((lx, ly, r, x, y, [(0, lx), (0, 104), (96, 104), (46, 37)]))
    lx, ly = 96, 72     # overlapping discs, discs in the wall ring, a disc partly outside the lattice
    dx = 1e-4 * lx / (lx - 1)
    cases.append((lx, ly, np.array([10, 9, 8, 7, 9, 6.5]) * dx, np.array([30, 39, 34, 3.0, 80, 95.0]) * dx,
                  np.array([30, 31, 38, 40, 2.5, 20]) * dx, [(0, lx), (20, 40)]))
    lx, ly = 160, 150   # discs larger than a tile
    dx = 1e-4 * lx / (lx - 1)
    cases.append((lx, ly, np.array([55.0, 30.0, 12.0]) * dx, np.array([70.0, 120.0, 20.0]) * dx, np.array([70.0, 100.0, 130.0]) * dx,
                  [(0, lx), (60, 50)]))
    for lx, ly, r, x, y, strips in cases:
        o = Oracle(lx, ly, 1.0, prec)
        n = o.init_arrays(r, x, y)
        sc = o.scalars()
        rng = np.random.default_rng(3)
        for rep in range(3):
            st = o.grains()[:, :9].copy()
            st[:, 0:2] += rng.uniform(-0.7, 0.7, size=(n, 2)) * sc["dx"]
            o.set_grain_state(st)
            scal = np.array([sc["dx"], sc["c"], sc["Mgx"], sc["Mby"], 0.0])
            for x0, nxl in strips:
                counts = np.zeros(3, dtype=np.int64)
                rc = fn(lx, ly, n, scal, _grain_table(o.grains()), x0, nxl, counts)
                assert rc == 0, (lx, ly, x0, nxl, rc)
                assert counts[0] > 0 and counts[2] >= 1
            # the whole-lattice map of the restatement is the oracle's
            o.lbm_step()


def _dem_params(sc):
    keys = ["kg", "kt", "km", "ktm", "nug", "num", "numb", "nugt", "mu", "mum", "mumb", "murf", "freq", "amp", "t",
            "distVerlet"]
    p = [DEM_CONST[k] for k in keys]
    p += [sc["dt"], sc["dt2"], sc["xG"], sc["yG"], sc["Mgx"], sc["Mdx"], sc["Mby"], sc["Mhy"]]
    return np.array(p, dtype=np.float64)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_dem_step_gather_form_bit_exact(prec):
    hc = load_hostcheck()
    fn = getattr(hc, f"hc_dem_step_{prec}")
    lx, ly = 64, 48
    o = Oracle(lx, ly, 1.0, prec)
    r, x, y = small_packing(lx, ly, 1.0, seed=5, overlap=0.03)
    n = o.init_arrays(r, x, y)
    v, w, a = random_kinematics(n, 6, vmax=0.02, wmax=5.0, amax=5.0)
    st = o.grains()[:, :9].copy()
    st[:, 3:5], st[:, 5:6], st[:, 6:9] = v, w, a
    o.set_grain_state(st)
    cap = 32
    cnt = np.zeros(n, dtype=np.int32)
    nbr = np.zeros((n, cap), dtype=np.int32)
    wfl = np.zeros(n, dtype=np.int32)
    touched = 0
    for step in range(230):
        nb = o.scalars()["nbsteps"]
        g = o.grains()
        state = np.ascontiguousarray(g[:, :9])
        props = np.ascontiguousarray(g[:, [9, 10, 11]])
        o.step(1)
        sc = o.scalars()          # Mdx/Mhy as left by VerletWall of this step
        fhf = o.fhf()             # as left by the LBM step of this step (if any)
        rc = fn(n, _dem_params(sc), int(nb % 8000 == 0), state, props, np.ascontiguousarray(fhf), int(nb % 100 == 0),
                cnt, nbr, cap, wfl)
        assert rc == 0
        ref = o.grains()[:, :9]
        assert np.array_equal(state, ref), f"DEM step {nb}: max diff {np.abs(state - ref).max()}"
        touched = max(touched, int(cnt.sum()))
    cum, half = o.verlet()
    assert touched == 2 * len(half) and touched > 0
    assert any(len(l) for l in o.wall_lists())
