"""Pins the oracle (oracle/lbmdem_oracle.c) to the reference.

Two anchors, both CPU-only:
  1. tests/golden/*.npz -- states written by the UNMODIFIED reference (src/main.c compiled in
     place, tools/make_golden.py), incl. the known-answer densities of SURVEY.md 4.4.  The
     oracle must reproduce them BIT FOR BIT (same compiler flags: -O2 -ffp-contract=off).
  2. the compiled reference itself (oracle/_ref/libref_*.so), where those libraries exist
     (built in the authoring container, shipped to the GPU box as binaries): random states are
     injected into both and single phases / whole steps compared bit for bit.
The reference's own repository holds no tests or golden vectors.
"""
import hashlib
import os

import numpy as np
import pytest

from oracle import build as obuild
from oracle.oraclewrap import Oracle
from util import perturbed_f, random_kinematics

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def sample_nodes(a):
    return np.ascontiguousarray(a[5::16, 7::16])


def check_snapshot(sim, gold, tag, n):
    """bit-exact comparison of a simulator state with a stored reference snapshot"""
    assert np.array_equal(sim.grains()[:, :9], gold[f"{tag}_grains"]), f"{tag}: grain state"
    assert np.array_equal(sim.fhf(), gold[f"{tag}_fhf"]), f"{tag}: hydrodynamic forces"
    obst = sim.obst()
    assert sha(obst) == str(gold[f"{tag}_obst_sha256"]), f"{tag}: obstacle map"
    assert int(((obst >= 0) & (obst < n)).sum()) == int(gold[f"{tag}_solid_nodes"])
    f = sim.f()
    if f"{tag}_f" in gold:
        assert np.array_equal(f, gold[f"{tag}_f"]), f"{tag}: populations"
    else:
        assert np.array_equal(sample_nodes(f), gold[f"{tag}_f_sample"]), f"{tag}: sampled populations"
    assert sha(f) == str(gold[f"{tag}_f_sha256"]), f"{tag}: populations (hash of all {f.size} values)"


def test_golden_a08d83_known_answers():
    gold = np.load(os.path.join(GOLD, "a08d83_512_f64.npz"))
    o = Oracle(512, 512, 1.0, "f64")
    n = o.init(os.path.join(GOLD, "a08d83.data"))
    assert n == 726
    sc = o.scalars()
    for k in sc:
        assert sc[k] == gold[f"scalar_{k}"], k
    # SURVEY.md 4.4, printed by the reference at start-up
    assert f"{sc['dx']:e}" == "1.001957e-04" and sc["npDEM"] == 10 and f"{sc['c']:f}" == "7.485352"
    assert np.array_equal(o.grains(), gold["init_grains"])
    assert sha(o.obst()) == str(gold["init_obst_sha256"])
    done = 0
    for upto in (15, 100, 202):
        o.step(upto - done)
        done = upto
        check_snapshot(o, gold, f"s{upto}", n)
        # check_density sums serially in the reference; the oracle does the same
        assert o.total_density() == float(gold[f"s{upto}_density"])
    assert f"{float(gold['s15_density']):.6f}" == "262144.004828"      # final_density, duration 2e-5
    assert f"{float(gold['s202_density']):.6f}" == "262144.109015"     # final_density, duration 2.7e-4
    cum, half = o.verlet()
    assert np.array_equal(half, gold["verlet_half"]) and np.array_equal(cum[:-1], gold["verlet_cumul"][:-1])
    for name, lst in zip("BTLR", o.wall_lists()):
        assert np.array_equal(lst, gold[f"wall_{name}"])


@pytest.mark.parametrize("name,prec", [("pack_64x48", "f64"), ("pack_64x48", "f32"),
                                       ("pack_256x256", "f64"), ("pack_256x256", "f32")])
def test_golden_packings(name, prec):
    gold = np.load(os.path.join(GOLD, f"{name}_{prec}.npz"))
    lx, ly = (int(v) for v in name.split("_")[1].split("x"))
    o = Oracle(lx, ly, 1.0, prec)
    n = o.init(os.path.join(GOLD, f"{name}_{prec}.data"))
    f0 = gold["start_f"] if "start_f" in gold else perturbed_f(lx, ly, int(gold["start_f_seed"]))
    o.set_f(f0)
    o.set_grain_state(gold["start_state"])
    o.step(int(gold["steps"]))
    check_snapshot(o, gold, "end", n)


# ---- live comparison with the compiled reference ----------------------------------------------
def _ref(lx, ly, prec):
    path = obuild.build_ref(lx, ly, "1.", prec)
    if path is None or not os.path.exists(path):
        pytest.skip(f"compiled reference {lx}x{ly} {prec} not available here")
    from oracle.refwrap import Reference
    return Reference(lx, ly, "1.", prec)


def _both(lx, ly, prec, sample):
    ref = _ref(lx, ly, prec)
    o = Oracle(lx, ly, 1.0, prec)
    cwd = os.getcwd()
    n = ref.init(sample)
    assert o.init(sample) == n
    os.chdir(cwd)
    return ref, o, n


def _assert_same(ref, o, what=""):
    assert np.array_equal(ref.obst(), o.obst()), what + " obst"
    assert np.array_equal(ref.act(), o.act()), what + " act"
    assert np.array_equal(ref.delta(), o.delta()), what + " delta"
    assert np.array_equal(ref.f(), o.f()), what + " f"
    assert np.array_equal(ref.fhf(), o.fhf()), what + " fhf"
    assert np.array_equal(ref.grains(), o.grains()), what + " grains"


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_oracle_equals_reference_phase_by_phase(prec, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)          # the reference writes stats.data into the cwd
    ref, o, n = _both(64, 48, prec, os.path.join(GOLD, f"pack_64x48_{prec}.data"))
    rng = np.random.default_rng(5)
    f0 = perturbed_f(64, 48, 6)
    for z in (ref, o):
        z.set_f(f0)
    for it in range(4):
        v, w, a = random_kinematics(n, 10 + it, vmax=0.05)
        st = ref.grains()[:, :9].copy()
        st[:, 0:2] += rng.uniform(-0.7, 0.7, size=(n, 2)) * ref.scalars()["dx"]
        st[:, 3:5], st[:, 5:6], st[:, 6:9] = v, w, a * 0.1
        for z in (ref, o):
            z.set_grain_state(st)
        for phase in ("reinit_obst_density", "obst_construction", "collision_streaming", "forces_fluid"):
            getattr(ref.lib, "ref_" + phase)()
            o.phase(phase)
            _assert_same(ref, o, f"iteration {it} after {phase}:")


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_oracle_equals_reference_coupled_run(prec, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    ref, o, n = _both(256, 256, prec, os.path.join(GOLD, f"pack_256x256_{prec}.data"))
    v, w, a = random_kinematics(n, 3, vmax=0.02)
    st = ref.grains()[:, :9].copy()
    st[:, 3:5], st[:, 5:6] = v, w
    f0 = perturbed_f(256, 256, 4)
    for z in (ref, o):
        z.set_f(f0)
        z.set_grain_state(st)
    for chunk in range(4):
        ref.step(53)
        o.step(53)
        _assert_same(ref, o, f"after {(chunk + 1) * 53} renderScene() calls:")
        cr, hr = ref.verlet()
        co, ho = o.verlet()
        assert np.array_equal(hr, ho) and np.array_equal(cr[:-1], co[:-1])
        for a_, b_ in zip(ref.wall_lists(), o.wall_lists()):
            assert np.array_equal(a_, b_)
    assert ref.total_density() == o.total_density()


def test_oracle_equals_reference_film_step(tmp_path, monkeypatch):
    """steps with nbsteps % 8000 == 0 use the in-lined alternate contact law (src/main.c:1342-1426)
    and write VTK/DEM files (into the cwd)."""
    monkeypatch.chdir(tmp_path)
    ref, o, n = _both(64, 48, "f64", os.path.join(GOLD, "pack_64x48_f64.data"))
    v, w, a = random_kinematics(n, 8, vmax=0.02)
    st = ref.grains()[:, :9].copy()
    st[:, 3:5], st[:, 5:6] = v, w
    for z in (ref, o):
        z.set_grain_state(st)
    ref.lib.ref_set_nbsteps(7997)
    o.set_nbsteps(7997)
    ref.step(6)
    o.step(6)
    _assert_same(ref, o, "across step 8000:")


def test_oracle_equals_reference_with_vibrating_walls(tmp_path, monkeypatch):
    """int vib = 1 (src/main.c:162, :1701-1706): the left / right walls and the lattice origin move every call"""
    monkeypatch.chdir(tmp_path)
    ref, o, n = _both(64, 48, "f64", os.path.join(GOLD, "pack_64x48_f64.data"))
    v, w, a = random_kinematics(n, 12, vmax=0.02)
    st = ref.grains()[:, :9].copy()
    st[:, 3:5], st[:, 5:6] = v, w
    for z in (ref, o):
        z.set_grain_state(st)
        z.set_vib(1)
    for chunk in range(3):
        ref.step(45)
        o.step(45)
        _assert_same(ref, o, f"vib, after {(chunk + 1) * 45} calls:")
        assert ref.scalars() == o.scalars()
    assert ref.scalars()["Mgx"] != 0.0
    ref.set_vib(0)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_oracle_equals_reference_top_right_walls_and_tilted_gravity(prec, tmp_path, monkeypatch):
    """The dormant switches of SURVEY 8(f)4: with dtt large the confining right / top walls stay at the lattice
    extent (src/main.c:117, :1555-1561), so a packing mirrored into the top-right corner drives force_WallT and
    force_WallR (:846-887, :923-951); angleG tilts gravity (:98, :1841-1842)."""
    from util import mirrored_state
    monkeypatch.chdir(tmp_path)
    ref, o, n = _both(64, 48, prec, os.path.join(GOLD, f"pack_64x48_{prec}.data"))
    try:
        for z in (ref, o):
            z.set_dtt(1.0)
            z.set_angleG(0.35)
        sc = ref.scalars()
        assert sc["xG"] != 0.0 and ref.scalars() == o.scalars()
        v, w, a = random_kinematics(n, 21, vmax=0.02)
        st = ref.grains()[:, :9].copy()
        st[:, 3:5], st[:, 5:6] = v, w
        st = mirrored_state(st, sc["Mdx"], sc["Mhy"])
        r = ref.grains()[:, 9]
        assert (st[:, 0] + r > sc["Mdx"]).any() and (st[:, 1] + r > sc["Mhy"]).any()   # grains INSIDE the two walls
        for z in (ref, o):
            z.set_grain_state(st)
        for chunk in range(3):
            ref.step(37)
            o.step(37)
            _assert_same(ref, o, f"dtt / angleG, after {(chunk + 1) * 37} calls:")
            assert ref.scalars() == o.scalars()
            wl = ref.wall_lists()
            for a_, b_ in zip(wl, o.wall_lists()):
                assert np.array_equal(a_, b_)
            assert len(wl[1]) > 0 and len(wl[3]) > 0, "top / right wall lists are empty"
        # the walls did NOT jump to 10 x the lattice extent (that is what dtt = 0 does at the first VerletWall)
        assert ref.scalars()["Mdx"] == sc["Mdx"] and ref.scalars()["Mhy"] == sc["Mhy"]
    finally:
        ref.set_dtt(0.0)
        ref.set_angleG(0.0)
