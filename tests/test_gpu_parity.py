"""GPU parity tests proper: the CUDA path, called through the C ABI (liblbmdem_gpu.so via
lbmdem_gpu.Solver), against the oracle on the same seeded inputs.

Bars (BASELINE.json north_star / SURVEY.md 4.3):
  - lattice node indices (obst) and act: bit-exact, always;
  - strict_fp=1 (no contraction, reference summation order): EVERYTHING bit-exact, any horizon;
  - default build (FMA contraction, fixed-point force sums): f, rho, u and grain trajectories
    within 1e-6 relative (fp64) / 1e-4 (fp32) at a <= 100-DEM-step horizon -- the packed-grain
    problem is chaotic beyond that even between two CPU builds of the reference;
  - one DEM step from identical input: bit-exact in both builds (aux kernels are never contracted).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.oraclewrap import Oracle
from util import perturbed_f, random_kinematics, small_packing

import lbmdem_gpu as G

REL_FAST = {"f64": 1e-6, "f32": 1e-4}      # the north_star tolerance, trajectory level
REL_STEP = {"f64": 1e-12, "f32": 2e-5}     # single step from identical state


def _pair(prec, lx, ly, scale=1.0, seed=0, n_target=None, lid=0.0, **over):
    o = Oracle(lx, ly, scale, prec)
    s = G.Solver(lx, ly, scale, prec, lid_u=lid, **over)
    r, x, y = small_packing(lx, ly, scale, seed, n_target=n_target)
    n = o.init_arrays(r, x, y)
    assert s.init_arrays(r, x, y) == n
    if lid:
        o.set_lid(lid)
    return o, s, n


def _same_start(o, s, n, seed, vmax=0.05):
    f0 = perturbed_f(o.lx, o.ly, seed)
    o.set_f(f0)
    s.set_f(f0)
    v, w, a = random_kinematics(n, seed + 1, vmax=vmax)
    st = o.grains()[:, :9].copy()
    st[:, 3:5], st[:, 5:6], st[:, 6:9] = v, w, a * 0.1
    o.set_grain_state(st)
    s.set_grain_state(st)


def _relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def test_setup_scalars_and_initial_map():
    for prec in ("f64", "f32"):
        o, s, n = _pair(prec, 64, 48, seed=3)
        so, ss = o.scalars(), s.scalars()
        for k in so:
            assert so[k] == ss[k], (prec, k, so[k], ss[k])
        assert np.array_equal(o.obst(), s.obst())
        assert np.array_equal(o.grains(), s.grains())
        assert np.array_equal(o.f(), s.f())


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("kernel", [0, 1])
def test_lbm_step_strict_is_bit_exact(prec, kernel):
    o, s, n = _pair(prec, 70, 131, seed=11, strict_fp=1, kernel=kernel)
    _same_start(o, s, n, 12)
    rng = np.random.default_rng(13)
    dx = o.scalars()["dx"]
    for step in range(4):
        st = o.grains()[:, :9].copy()
        st[:, 0:2] += rng.uniform(-0.6, 0.6, size=(n, 2)) * dx     # nodes change state
        o.set_grain_state(st)
        s.set_grain_state(st)
        o.lbm_step()
        s.lbm_step()
        assert np.array_equal(o.obst(), s.obst()), f"obst, step {step}"
        solid = (o.obst() >= 0) & (o.obst() < n)
        assert np.array_equal(o.act()[solid], s.act()[solid]), f"act, step {step}"
        fo, fs = o.f(), s.f()
        bad = np.argwhere(fo != fs)
        assert bad.size == 0, f"step {step}: {len(bad)} populations differ, first {bad[:4].tolist()}"
        assert np.array_equal(o.fhf(), s.fhf()), f"fhf, step {step}"


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("kernel", [0, 1])
def test_lbm_step_default_build(prec, kernel):
    o, s, n = _pair(prec, 130, 70, seed=21, kernel=kernel, lid=0.03)
    _same_start(o, s, n, 22)
    for step in range(3):
        o.lbm_step()
        s.lbm_step()
        assert np.array_equal(o.obst(), s.obst())
        assert _relerr(s.f(), o.f()) < REL_STEP[prec], f"f, step {step}"
        fo, fs = o.fhf(), s.fhf()
        assert _relerr(fs, fo) < (1e-9 if prec == "f64" else 5e-3), f"fhf, step {step}"
        # keep the two in lock-step so that every step is a single-step comparison
        s.set_f(o.f())
        s.set_fhf(fo)


def test_tiled_kernel_equals_generic_kernel_bitwise():
    """Same arithmetic, different plumbing (TMA tiles vs on-demand loads): identical bits."""
    lx, ly = 200, 333
    a = G.Solver(lx, ly, 1.0, "f64", kernel=0, strict_fp=1)
    b = G.Solver(lx, ly, 1.0, "f64", kernel=1, strict_fp=1)
    r, x, y = small_packing(lx, ly, 1.0, seed=5, n_target=300)
    n = a.init_arrays(r, x, y)
    b.init_arrays(r, x, y)
    f0 = perturbed_f(lx, ly, 6)
    v, w, acc = random_kinematics(n, 7)
    st = a.grains()[:, :9].copy()
    st[:, 3:5], st[:, 5:6] = v, w
    for slv in (a, b):
        slv.set_f(f0)
        slv.set_grain_state(st)
    for _ in range(3):
        a.step(a.scalars()["npDEM"])
        b.step(b.scalars()["npDEM"])
    assert np.array_equal(a.f(), b.f())
    assert np.array_equal(a.fhf(), b.fhf())
    assert np.array_equal(a.grains(), b.grains())


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_dem_steps_bit_exact_with_contacts(prec):
    o, s, n = _pair(prec, 64, 48, seed=5)
    v, w, a = random_kinematics(n, 6, vmax=0.02, wmax=5.0, amax=5.0)
    st = o.grains()[:, :9].copy()
    st[:, 3:5], st[:, 5:6], st[:, 6:9] = v, w, a
    o.set_grain_state(st)
    s.set_grain_state(st)
    # pure DEM: keep the LBM out by holding fhf fixed and stepping between LBM steps
    npd = o.scalars()["npDEM"]
    fh = np.random.default_rng(1).uniform(-1e-3, 1e-3, size=(n, 3))
    done = 0
    for nb in range(1, 260):
        if nb % npd == 0:
            continue
        o.set_nbsteps(nb)
        s.set_nbsteps(nb)
        o.set_fhf(fh)
        s.set_fhf(fh)
        if nb == 1 or nb % 100 == 1:
            o.phase("init_verlet")
            s.build_verlet()
            cum_o, half_o = o.verlet()
            cum_s, half_s = s.verlet()
            assert np.array_equal(half_o, half_s) and np.array_equal(cum_o[:-1], cum_s[:-1])
            for lo, ls in zip(o.wall_lists(), s.wall_lists()):
                assert np.array_equal(lo, ls)
            assert len(half_o) > 0 and any(len(l) for l in o.wall_lists())
        o.step(1)
        s.step(1)
        done += 1
        assert np.array_equal(o.grains()[:, :9], s.grains()[:, :9]), f"DEM step {nb}"
    assert done > 200


def test_film_step_contact_law_bit_exact():
    """nbsteps % stepFilm == 0: LBM step + Verlet rebuild + the alternate in-lined contact law
    (src/main.c:1342-1426) in the same renderScene() call."""
    o, s, n = _pair("f64", 64, 48, seed=8, strict_fp=1)
    _same_start(o, s, n, 9, vmax=0.02)
    for z in (o, s):
        z.set_nbsteps(7998)
        z.step(4)                      # 7998, 7999 (normal law), 8000 (alternate law), 8001
    assert np.array_equal(o.grains()[:, :9], s.grains()[:, :9])
    assert o.scalars()["nbsteps"] == s.scalars()["nbsteps"] == 8002
    assert o.scalars()["nFile"] == s.scalars()["nFile"] == 1
    # and the two laws do differ on this input (the test would be vacuous otherwise)
    o2, s2, _ = _pair("f64", 64, 48, seed=8, strict_fp=1)
    _same_start(o2, s2, n, 9, vmax=0.02)
    o2.set_nbsteps(6998)
    o2.step(4)
    assert not np.array_equal(o2.grains()[:, 3:6], o.grains()[:, 3:6])


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_coupled_trajectory_strict_is_bit_exact(prec):
    o, s, n = _pair(prec, 96, 80, seed=31, n_target=60, strict_fp=1)
    _same_start(o, s, n, 32, vmax=0.02)
    for chunk in range(5):
        o.step(37)
        s.step(37)
        assert np.array_equal(o.grains()[:, :9], s.grains()[:, :9]), f"grains after {(chunk + 1) * 37} DEM steps"
        assert np.array_equal(o.obst(), s.obst())
    assert np.array_equal(o.f(), s.f())
    assert np.array_equal(o.fhf(), s.fhf())
    assert o.scalars()["nbsteps"] == s.scalars()["nbsteps"] == 185


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_coupled_trajectory_default_build_100_dem_steps(prec):
    """north_star: grain trajectories and rho, u within 1e-6 relative (fp64) at the horizon
    where two CPU builds of the reference still agree."""
    lx, ly = 256, 256
    o, s, n = _pair(prec, lx, ly, seed=41, n_target=200)
    _same_start(o, s, n, 42, vmax=0.01)
    horizon = 100 if prec == "f64" else 30
    o.step(horizon)
    s.step(horizon)
    tol = REL_FAST[prec]
    go, gs = o.grains(), s.grains()
    assert np.array_equal(o.obst(), s.obst())                      # node indices bit-exact
    assert _relerr(gs[:, 0:3], go[:, 0:3]) < tol                   # positions
    assert _relerr(gs[:, 3:6], go[:, 3:6]) < tol                   # velocities
    fo, fs = o.f(), s.f()
    rho_o, rho_s = fo.sum(-1), fs.sum(-1)
    ex = np.array([0, -1, -1, -1, 0, 1, 1, 1, 0.0])
    ey = np.array([0, 1, 0, -1, -1, -1, 0, 1, 1.0])
    assert _relerr(rho_s, rho_o) < tol
    for e in (ex, ey):
        jo, js = (fo * e).sum(-1), (fs * e).sum(-1)
        assert np.abs(js - jo).max() < tol * max(np.abs(jo).max(), 1e-3)
    assert _relerr(s.fhf(), o.fhf()) < (1e-6 if prec == "f64" else 5e-2)
    assert abs(s.total_density() - o.total_density()) < (1e-9 if prec == "f64" else 1e-2) * lx * ly


def test_total_density_and_fields():
    o, s, n = _pair("f64", 70, 90, seed=51)
    _same_start(o, s, n, 52)
    o.step(12)
    s.step(12)
    fo = o.f()
    assert abs(s.total_density() - fo.sum()) < 1e-9 * fo.size
    fl = s.fields(grain_p=np.arange(n, dtype=float))
    obst = o.obst()
    fs = s.f()
    # reference formulas (src/main.c:284-323), float accumulation one population at a time
    ex = np.array([0, -1, -1, -1, 0, 1, 1, 1, 0.0])
    acc = np.zeros(obst.shape, dtype=np.float32)
    jx = np.zeros(obst.shape, dtype=np.float32)
    for q in range(9):
        acc = (acc.astype(np.float64) + fs[:, :, q]).astype(np.float32)
        jx = (jx.astype(np.float64) + fs[:, :, q] * ex[q]).astype(np.float32)
    press = ((1.0 / 3.0) * 1000.0 * (acc.astype(np.float64) - 1.0)).astype(np.float32)
    solid = (obst >= 0) & (obst < n)
    assert np.array_equal(fl["fluid_pressure"].T[~solid], press[~solid])
    assert np.array_equal(fl["fluid_velocity"][:, :, 0].T[~solid], jx[~solid])
    assert np.all(fl["fluid_pressure"].T[solid] == 0)
    assert np.array_equal(fl["grain_pressure"].T[solid], obst[solid].astype(np.float32))
    assert np.all(fl["grain_pressure"].T[~solid] == -1)
    g = s.grains()
    assert np.array_equal(fl["grain_velocity"][:, :, 0].T[solid], g[obst[solid], 3].astype(np.float32))


def test_step_host_end_to_end_call():
    o, s, n = _pair("f64", 64, 48, seed=61, strict_fp=1)
    _same_start(o, s, n, 62)
    st = o.grains()[:, :9].copy()
    npd = o.scalars()["npDEM"]
    o.step(npd)
    sout, fh, dens = s.step_host(st, npd)
    assert np.array_equal(sout, o.grains()[:, :9])
    assert np.array_equal(fh, o.fhf())
    assert abs(dens - o.f().sum()) < 1e-9 * o.f().size


def test_step_host_float_rows():
    """lbmdem_step_host_f32: grain rows of float, what a -DSINGLE_PRECISION reference holds; the same bits as the
    double rows rounded to float, chained over several calls through the page-locked buffers"""
    lx, ly = 96, 80
    r, x, y = small_packing(lx, ly, 1.0, seed=63, n_target=40)
    a = G.Solver(lx, ly, 1.0, "f32")
    b = G.Solver(lx, ly, 1.0, "f32")
    n = a.init_arrays(r, x, y)
    assert b.init_arrays(r, x, y) == n
    npd = a.scalars()["npDEM"]
    sa = a.grains()[:, :9].copy()
    sb = sa.astype(np.float32)
    for _ in range(4):
        sa, fa, da = a.step_host(sa, npd)
        sb, fb, db = b.step_host(sb, npd, rows="f32")
        assert sb.dtype == np.float32 and fb.dtype == np.float32
        assert np.array_equal(sa.astype(np.float32), sb) and np.array_equal(fa.astype(np.float32), fb)
        assert da == db
    d = G.Solver(lx, ly, 1.0, "f64")
    d.init_arrays(r, x, y)
    with pytest.raises(G.LbmdemError) as ei:
        d.step_host(sb, npd, rows="f32")
    assert ei.value.code == -1


def test_errors_are_reported():
    s = G.Solver(64, 48)
    with pytest.raises(G.LbmdemError) as ei:
        s.step(1)
    assert ei.value.code == -4
    with pytest.raises(G.LbmdemError):
        s.init("/nonexistent/sample.data")
    # neighbour capacity overflow is an error, not a print (src/main.c:1535)
    t = G.Solver(64, 48, neighbour_capacity=2)
    r, x, y = small_packing(64, 48, 1.0, seed=1)
    t.init_arrays(r, x, y)
    with pytest.raises(G.LbmdemError) as ei:
        t.step(1)
    assert ei.value.code == -6


# ---- against the committed reference fixtures (tests/golden, written by the compiled reference) ----
import hashlib
import os

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_golden_a08d83_strict_build_is_bit_exact():
    """bin/a08d83.data on a 512 x 512 fp64 lattice: the strict CUDA build must give
    the reference's own bits after 15, 100 and 202 renderScene() calls."""
    gold = np.load(os.path.join(GOLD, "a08d83_512_f64.npz"))
    s = G.Solver(512, 512, 1.0, "f64", strict_fp=1)
    n = s.init(os.path.join(GOLD, "a08d83.data"))
    assert n == 726
    sc = s.scalars()
    for k in sc:
        assert sc[k] == gold[f"scalar_{k}"], k
    assert np.array_equal(s.grains(), gold["init_grains"])
    assert _sha(s.obst()) == str(gold["init_obst_sha256"])
    done = 0
    for upto in (15, 100, 202):
        s.step(upto - done)
        done = upto
        tag = f"s{upto}"
        assert np.array_equal(s.grains()[:, :9], gold[f"{tag}_grains"]), tag
        assert np.array_equal(s.fhf(), gold[f"{tag}_fhf"]), tag
        assert _sha(s.obst()) == str(gold[f"{tag}_obst_sha256"]), tag
        assert _sha(s.f()) == str(gold[f"{tag}_f_sha256"]), tag
        assert abs(s.total_density() - float(gold[f"{tag}_density"])) < 1e-9 * 512 * 512


def test_golden_a08d83_default_build_within_tolerance():
    """Same run with the default (contracted, fixed-point force sums) build: node indices bit-exact,
    trajectories and rho, u within 1e-6 relative at the 100-call horizon (north_star)."""
    gold = np.load(os.path.join(GOLD, "a08d83_512_f64.npz"))
    s = G.Solver(512, 512, 1.0, "f64")
    s.init(os.path.join(GOLD, "a08d83.data"))
    s.step(100)
    go, gs = gold["s100_grains"], s.grains()[:, :9]
    assert _sha(s.obst()) == str(gold["s100_obst_sha256"])
    assert _relerr(gs[:, 0:3], go[:, 0:3]) < 1e-6 and _relerr(gs[:, 3:6], go[:, 3:6]) < 1e-6
    assert _relerr(s.fhf(), gold["s100_fhf"]) < 1e-6
    fs, fo = s.f()[5::16, 7::16], gold["s100_f_sample"]
    assert _relerr(fs.sum(-1), fo.sum(-1)) < 1e-6
    ex = np.array([0, -1, -1, -1, 0, 1, 1, 1, 0.0])
    ey = np.array([0, 1, 0, -1, -1, -1, 0, 1, 1.0])
    for e in (ex, ey):
        jo, js = (fo * e).sum(-1), (fs * e).sum(-1)
        assert np.abs(js - jo).max() < 1e-6 * max(np.abs(jo).max(), 1e-3)
    assert abs(s.total_density() - float(gold["s100_density"])) < 1e-9 * 512 * 512


@pytest.mark.parametrize("name,prec", [("pack_64x48", "f64"), ("pack_64x48", "f32"),
                                       ("pack_256x256", "f64"), ("pack_256x256", "f32")])
def test_golden_packings_strict_build_is_bit_exact(name, prec):
    gold = np.load(os.path.join(GOLD, f"{name}_{prec}.npz"))
    lx, ly = (int(v) for v in name.split("_")[1].split("x"))
    s = G.Solver(lx, ly, 1.0, prec, strict_fp=1)
    s.init(os.path.join(GOLD, f"{name}_{prec}.data"))
    f0 = gold["start_f"] if "start_f" in gold else perturbed_f(lx, ly, int(gold["start_f_seed"]))
    s.set_f(f0)
    s.set_grain_state(gold["start_state"])
    s.step(int(gold["steps"]))
    assert np.array_equal(s.grains()[:, :9], gold["end_grains"])
    assert np.array_equal(s.fhf(), gold["end_fhf"])
    assert _sha(s.obst()) == str(gold["end_obst_sha256"])
    assert _sha(s.f()) == str(gold["end_f_sha256"])


def test_lbmdem_executable_writes_the_reference_files(tmp_path):
    """The drop-in binary (2d-lbm-dem_b200/host/lbmdem_main.c): 8000 renderScene() calls of the 64 x 48
    packing from rest, strict build.  stats.data, DEM000000.dat, DEM000001.dat and the five VTK files of
    frame 0 must be byte-identical to what the compiled reference wrote (tests/golden/outputs_64x48)."""
    import importlib.util
    import subprocess
    spec = importlib.util.spec_from_file_location("host_build", os.path.join(os.path.dirname(GOLD), "..", "2d-lbm-dem_b200",
                                                                             "host", "build.py"))
    hb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(hb)
    exe = hb.build_exe()
    ref_dir = os.path.join(GOLD, "outputs_64x48")
    p = subprocess.run([exe, os.path.join(GOLD, "pack_64x48_f64.data"), "--lx", "64", "--ly", "48", "--strict", "--steps", "8000",
                        "--outdir", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "Nb grains 7" in p.stdout and "Iteration Number 0, Total density in the system" in p.stdout
    assert "final_density:" in p.stderr
    for name in sorted(os.listdir(ref_dir)):
        if name.endswith(".npz"):
            continue
        mine, ref = open(tmp_path / name, "rb").read(), open(os.path.join(ref_dir, name), "rb").read()
        assert mine == ref, f"{name} differs from the reference's file"


# ---- edge cases and full-size properties ---------------------------------------------------------
def test_far_grain_lid_driven_cavity():
    """BASELINE configs[1]: no grain inside the lattice (the reference cannot run with zero grains: one
    grain outside the lattice, inside the DEM walls), moving lid.  No contacts, hence no chaos: the
    strict build stays bit-exact and the default build within 1e-9 over 300 renderScene() calls."""
    lx = ly = 128
    for strict in (1, 0):
        o = Oracle(lx, ly, 1.0, "f64")
        s = G.Solver(lx, ly, 1.0, "f64", lid_u=0.05, strict_fp=strict)
        r, x, y = np.array([1.0e-3]), np.array([0.5 * lx * 1e-3]), np.array([1.0e-3])
        assert o.init_arrays(r, x, y) == s.init_arrays(r, x, y) == 1
        o.set_lid(0.05)
        o.step(300)
        s.step(300)
        fo, fs = o.f(), s.f()
        assert (o.obst() >= 0).sum() == 2 * lx + 2 * ly - 4          # only the wall ring is solid
        assert np.array_equal(o.obst(), s.obst())
        jx = (fo * np.array([0, -1, -1, -1, 0, 1, 1, 1, 0.0])).sum(-1)
        assert np.abs(jx).max() > 1e-3                                # the lid does drive a flow
        if strict:
            assert np.array_equal(fo, fs) and np.array_equal(o.grains(), s.grains())
        else:
            assert _relerr(fs, fo) < 1e-9


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_overlapping_discs_and_discs_on_the_ring_strict(prec):
    """Reduced discs that overlap each other (> 15 % interpenetration), discs that reach into the wall
    ring and a disc partly outside the lattice: owner = highest index, links into other grains and into
    the ring count in forces_fluid, deep nodes of overlapping grains are not skipped."""
    lx, ly = 96, 72
    dx = 1e-4 * lx / (lx - 1)
    r = np.array([10, 9, 8, 7, 9, 6.5]) * dx
    x = np.array([30, 39, 34, 3.0, 80, 95.0]) * dx            # 0-1-2 overlap pairwise; 3 on the left ring; 5 sticks out on the right
    y = np.array([30, 31, 38, 40, 2.5, 20]) * dx              # 4 on the bottom ring
    o = Oracle(lx, ly, 1.0, prec)
    s = G.Solver(lx, ly, 1.0, prec, strict_fp=1)
    n = o.init_arrays(r, x, y)
    assert s.init_arrays(r, x, y) == n
    _same_start(o, s, n, 77, vmax=0.03)
    ob = o.obst()
    assert ((ob >= 0) & (ob < n)).sum() > 500 and len(np.unique(ob)) == n + 2
    for step in range(4):
        o.lbm_step()
        s.lbm_step()
        assert np.array_equal(o.obst(), s.obst())
        solid = (o.obst() >= 0) & (o.obst() < n)
        assert np.array_equal(o.act()[solid], s.act()[solid])
        assert np.array_equal(o.f(), s.f()), f"step {step}"
        assert np.array_equal(o.fhf(), s.fhf()), f"step {step}"
        st = o.grains()[:, :9].copy()
        st[:, 0:2] += np.random.default_rng(step).uniform(-0.5, 0.5, size=(n, 2)) * dx
        o.set_grain_state(st)
        s.set_grain_state(st)
    # the default build agrees too (fixed-point force sums, list-driven kernels)
    d = G.Solver(lx, ly, 1.0, prec)
    d.init_arrays(r, x, y)
    d.set_obst(o.obst())          # the map reinit_obst_density will treat as "old"
    d.set_f(o.f())
    d.set_grain_state(o.grains()[:, :9])
    o.lbm_step()
    d.lbm_step()
    assert np.array_equal(o.obst(), d.obst())
    assert _relerr(d.f(), o.f()) < REL_STEP[prec]
    assert _relerr(d.fhf(), o.fhf()) < (1e-9 if prec == "f64" else 5e-3)


@pytest.mark.parametrize("case", ["dense", "overlap_ring"])
def test_incremental_rasteriser_equals_full_rebuild(case):
    """By default the rasteriser rebuilds only the lattice tiles in which some covered node changed since the previous
    step (grains move by a small fraction of a node per LBM step) and carries the rest of the map and of the two link
    lists over; with kernel=2 every tile is rebuilt every step.  Same map, same act, same populations and forces, bit
    for bit -- and the default run must really have skipped tiles."""
    if case == "dense":
        lx, ly = 200, 333
        r, x, y = small_packing(lx, ly, 1.0, seed=5, n_target=300)
    else:
        lx, ly = 96, 72
        dx = 1e-4 * lx / (lx - 1)
        r = np.array([10, 9, 8, 7, 9, 6.5]) * dx
        x = np.array([30, 39, 34, 3.0, 80, 95.0]) * dx
        y = np.array([30, 31, 38, 40, 2.5, 20]) * dx
    a = G.Solver(lx, ly, 1.0, "f64", kernel=0, strict_fp=1)
    b = G.Solver(lx, ly, 1.0, "f64", kernel=2, strict_fp=1)
    n = a.init_arrays(r, x, y)
    b.init_arrays(r, x, y)
    f0 = perturbed_f(lx, ly, 6)
    v, w, acc = random_kinematics(n, 7, vmax=0.3)      # fast enough for nodes to change owner every few LBM steps
    st = a.grains()[:, :9].copy()
    st[:, 3:5], st[:, 5:6] = v, w
    for slv in (a, b):
        slv.set_f(f0)
        slv.set_grain_state(st)
    ntiles = ((lx + 31) // 32) * ((ly + 63) // 64)
    rebuilt, obst0 = [], a.obst()
    npd = a.scalars()["npDEM"]
    for _ in range(12):
        a.step(npd)
        b.step(npd)
        assert np.array_equal(a.obst(), b.obst())
        assert np.array_equal(a.act(), b.act())
        ca, cb = a.list_counts(), b.list_counts()
        assert (ca["links"], ca["boundary_nodes"], ca["deferred"]) == (cb["links"], cb["boundary_nodes"], cb["deferred"])
        assert cb["tiles_rebuilt"] == ntiles
        rebuilt.append(ca["tiles_rebuilt"])
    assert not np.array_equal(a.obst(), obst0), "no node changed owner: the test does not exercise the incremental path"
    assert min(rebuilt[3:]) < ntiles, rebuilt
    assert np.array_equal(a.f(), b.f())
    assert np.array_equal(a.fhf(), b.fhf())
    assert np.array_equal(a.grains(), b.grains())
    # default build: the force sums are fixed point, hence independent of the order of the list entries too
    c = G.Solver(lx, ly, 1.0, "f64", kernel=0)
    d = G.Solver(lx, ly, 1.0, "f64", kernel=2)
    for slv in (c, d):
        slv.init_arrays(r, x, y)
        slv.set_f(f0)
        slv.set_grain_state(st)
        slv.step(8 * npd)
    assert np.array_equal(c.f(), d.f()) and np.array_equal(c.fhf(), d.fhf())


def test_grains_at_rest_leave_every_tile_alone():
    """LBM steps without DEM sub-steps: after the two rebuilds that follow set-up no tile is touched again"""
    lx, ly = 200, 160
    r, x, y = small_packing(lx, ly, 1.0, seed=15, n_target=120)
    o = Oracle(lx, ly, 1.0, "f64")
    s = G.Solver(lx, ly, 1.0, "f64", strict_fp=1)
    n = o.init_arrays(r, x, y)
    assert s.init_arrays(r, x, y) == n
    f0 = perturbed_f(lx, ly, 16)
    for z in (o, s):
        z.set_f(f0)
    for k in range(5):
        o.lbm_step()
        s.lbm_step()
    assert s.list_counts()["tiles_rebuilt"] == 0
    assert np.array_equal(o.f(), s.f()) and np.array_equal(o.obst(), s.obst()) and np.array_equal(o.fhf(), s.fhf())


def test_full_size_row_kernel_equals_plain_kernel_fp64():
    """BASELINE configs[2] size (2048 x 2048, fp64, 726 grains): the TMA row pipeline and the plain
    one-thread-per-node kernel are the same map; every population identical after 3 coupled steps."""
    import make_sample as ms
    lx = ly = 2048
    n, r_min, r_max, width = ms.PRESETS["a08d83"]
    r, x, y = ms.packed_sample(n, r_min, r_max, width, seed=12345)
    a = G.Solver(lx, ly, 1.0, "f64", kernel=0)
    b = G.Solver(lx, ly, 1.0, "f64", kernel=1)
    for slv in (a, b):
        assert slv.init_arrays(r * 1e-3, x * 1e-3, y * 1e-3) == n
        slv.step(3 * slv.scalars()["npDEM"] + 1)
    assert np.array_equal(a.obst(), b.obst())
    assert np.array_equal(a.fhf(), b.fhf()) and np.array_equal(a.grains(), b.grains())
    fa = a.f()
    assert np.array_equal(fa, b.f())
    assert np.isfinite(fa).all() and abs(fa.sum() / (lx * ly) - 1.0) < 1e-6


def test_full_size_fp32_properties():
    """BASELINE configs[3] size (4096 x 4096, fp32, 6355 grains): row kernel == plain kernel through
    the order-free checksums (fixed-shape density reduction, fixed-point force sums), and the lattice
    mass stays put."""
    import make_sample as ms
    lx = ly = 4096
    n, r_min, r_max, width = ms.PRESETS["a08_7000"]
    r, x, y = ms.packed_sample(n, r_min, r_max, width, seed=12345)
    a = G.Solver(lx, ly, 2.7, "f32", kernel=0)
    b = G.Solver(lx, ly, 2.7, "f32", kernel=1)
    dens = []
    for slv in (a, b):
        assert slv.init_arrays(r * 1e-3, x * 1e-3, y * 1e-3) == n
        slv.step(4 * slv.scalars()["npDEM"] + 1)
        dens.append(slv.total_density())
    assert dens[0] == dens[1]
    assert np.array_equal(a.fhf(), b.fhf()) and np.array_equal(a.grains(), b.grains())
    assert abs(dens[0] / (lx * ly) - 1.0) < 1e-5
    assert np.abs(a.fhf()).max() > 0


def test_vibrating_walls_strict_is_bit_exact():
    """int vib = 1 (src/main.c:162, :1701-1706)"""
    o = Oracle(96, 80, 1.0, "f64")
    s = G.Solver(96, 80, 1.0, "f64", strict_fp=1, vib=1)
    r, x, y = small_packing(96, 80, 1.0, 91, n_target=60)
    n = o.init_arrays(r, x, y)
    assert s.init_arrays(r, x, y) == n
    o.set_vib(1)
    _same_start(o, s, n, 92, vmax=0.02)
    for chunk in range(3):
        o.step(41)
        s.step(41)
        assert o.scalars() == s.scalars()
        assert np.array_equal(o.grains()[:, :9], s.grains()[:, :9])
        assert np.array_equal(o.obst(), s.obst())
    assert np.array_equal(o.f(), s.f()) and np.array_equal(o.fhf(), s.fhf())
    assert o.scalars()["Mgx"] != 0.0


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_checkpoint_restart_is_bit_exact(prec, tmp_path):
    """lbmdem_save_state / lbmdem_load_state: a run continued from the file equals the uninterrupted
    one bit for bit, in the default build, also when the file is written between two LBM steps and
    between two Verlet rebuilds."""
    lx, ly = 120, 96
    r, x, y = small_packing(lx, ly, 1.0, seed=101, n_target=70)
    a = G.Solver(lx, ly, 1.0, prec)
    n = a.init_arrays(r, x, y)
    v, w, acc = random_kinematics(n, 102, vmax=0.02)
    st = a.grains()[:, :9].copy()
    st[:, 3:5], st[:, 5:6] = v, w
    a.set_f(perturbed_f(lx, ly, 103))
    a.set_grain_state(st)
    a.step(137)                                   # not a multiple of npDEM nor of UpdateVerlet
    ck = str(tmp_path / "state.ck")
    a.save_state(ck)
    a.step(150)
    b = G.Solver(lx, ly, 1.0, prec)
    assert b.load_state(ck) == n
    assert b.scalars()["nbsteps"] == 137
    b.step(150)
    assert a.scalars() == b.scalars()
    assert np.array_equal(a.grains(), b.grains())
    assert np.array_equal(a.fhf(), b.fhf())
    assert np.array_equal(a.obst(), b.obst())
    assert np.array_equal(a.f(), b.f())
    with pytest.raises(G.LbmdemError):
        G.Solver(lx + 1, ly, 1.0, prec).load_state(ck)


def test_reference_default_build_size_strict(tmp_path):
    """BASELINE configs[0]: the reference's DEFAULT build (7826 x 2325, fp64) on the 47 980-grain synthetic
    stand-in for bin/50000-test.data, 37 renderScene() calls.  The strict build reproduces the compiled
    reference's bits (hashes in tests/golden/default_build_50000.npz); grains beyond the lattice are legal."""
    import make_sample as ms
    gold = np.load(os.path.join(GOLD, "default_build_50000.npz"))
    n, r_min, r_max, width = ms.PRESETS["50000-test"]
    r, x, y = ms.packed_sample(n, r_min, r_max, width, seed=12345)
    path = str(tmp_path / "s50k.data")
    ms.write_sample(path, r, x, y, comment="# synthetic 50000-test seed=12345")
    assert hashlib.sha256(open(path, "rb").read()).hexdigest() == str(gold["sample_sha256"])
    s = G.Solver(7826, 2325, 1.0, "f64", strict_fp=1)
    assert s.init(path) == n
    sc = s.scalars()
    for k in sc:
        assert sc[k] == gold[f"scalar_{k}"], k
    s.step(int(gold["steps"]))
    g = s.grains()
    assert np.array_equal(g[::97, :9], gold["grains_sample"])
    assert _sha(g[:, :9]) == str(gold["grains_sha256"])
    assert _sha(s.fhf()) == str(gold["fhf_sha256"])
    assert _sha(s.obst()) == str(gold["obst_sha256"])
    assert _sha(s.f()) == str(gold["f_sha256"])
    # the reference adds 1.6e8 terms serially; the device reduces pairwise: the SUMS differ in the 9th digit
    # although every addend is identical (f hash above)
    assert abs(s.total_density() - float(gold["density"])) < 1e-8 * 7826 * 2325


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_one_launch_dem_equals_three_launch_dem(prec):
    """The DEM sub-steps between two LBM steps run as ONE launch (a thread-block cluster up to 1024 grains, a cooperative grid
    above: grid barriers where the reference's loops end) by default and as three launches per sub-step with
    kernel=4.  Same arithmetic in the same order: same bits, and the oracle's bits in the strict build -- here with
    more than 1024 grains, across a Verlet rebuild, and across the film step (alternate contact law)."""
    lx, ly = 400, 300
    r, x, y = small_packing(lx, ly, 1.0, seed=31, n_target=1500)
    o = Oracle(lx, ly, 1.0, prec)
    a = G.Solver(lx, ly, 1.0, prec, strict_fp=1)
    b = G.Solver(lx, ly, 1.0, prec, strict_fp=1, kernel=4)
    n = o.init_arrays(r, x, y)
    assert n > 1024 and a.init_arrays(r, x, y) == n and b.init_arrays(r, x, y) == n
    v, w, acc = random_kinematics(n, 32, vmax=0.02)
    st = o.grains()[:, :9].copy()
    st[:, 3:5], st[:, 5:6] = v, w
    f0 = perturbed_f(lx, ly, 33)
    for z in (o, a, b):
        z.set_f(f0)
        z.set_grain_state(st)
        z.set_nbsteps(7950)                 # the film step (8000) and a Verlet rebuild fall inside the run
    for chunk in range(3):
        for z in (o, a, b):
            z.step(37)
        assert np.array_equal(a.grains(), b.grains()) and np.array_equal(a.fhf(), b.fhf())
        assert np.array_equal(o.grains()[:, :9], a.grains()[:, :9]), chunk
        assert np.array_equal(o.fhf(), a.fhf())
    assert np.array_equal(o.f(), a.f()) and np.array_equal(a.f(), b.f())
    # default build: fixed-point sums turned into fhf inside the DEM launch
    c = G.Solver(lx, ly, 1.0, prec)
    d = G.Solver(lx, ly, 1.0, prec, kernel=4)
    for z in (c, d):
        z.init_arrays(r, x, y)
        z.set_f(f0)
        z.set_grain_state(st)
        z.step(45)
    assert np.array_equal(c.grains(), d.grains()) and np.array_equal(c.fhf(), d.fhf()) and np.array_equal(c.f(), d.f())


def test_checkpoint_is_validated(tmp_path):
    """a checkpoint continued with other physics, or a truncated / doctored file, is refused (no silent drift, no
    allocation sized by an untrusted header)"""
    lx, ly = 96, 72
    r, x, y = small_packing(lx, ly, 1.0, seed=11, n_target=30)
    a = G.Solver(lx, ly, 1.0, "f64")
    a.init_arrays(r, x, y)
    a.step(25)
    ck = str(tmp_path / "s.ck")
    a.save_state(ck)
    with pytest.raises(G.LbmdemError) as ei:
        G.Solver(lx, ly, 1.0, "f64", tau=0.51).load_state(ck)
    assert ei.value.code == -1 and "physical parameters" in str(ei.value)
    raw = open(ck, "rb").read()
    open(ck, "wb").write(raw[:-1000])
    with pytest.raises(G.LbmdemError) as ei:
        G.Solver(lx, ly, 1.0, "f64").load_state(ck)
    assert ei.value.code == -5
    import struct
    bad = bytearray(raw)
    struct.pack_into("<i", bad, 8 + 7 * 4, 1 << 30)      # the neighbour capacity field of the header
    open(ck, "wb").write(bytes(bad))
    with pytest.raises(G.LbmdemError) as ei:
        G.Solver(lx, ly, 1.0, "f64").load_state(ck)
    assert ei.value.code == -5
    open(ck, "wb").write(raw)
    b = G.Solver(lx, ly, 1.0, "f64")
    assert b.load_state(ck) == a.n
