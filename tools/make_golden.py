"""tools/make_golden.py -- writes tests/golden/*.npz from the COMPILED REFERENCE.

Runs only where /root/reference exists (the authoring container): every vector below is
produced by the unmodified src/main.c of cb-geo/2d-lbm-dem, compiled in place by
oracle/build.py (gcc -std=c99 -O2 -ffp-contract=off, serial) and driven through
oracle/ref_shim.c.  The fixtures travel to the GPU box, the reference does not.

    python tools/make_golden.py            # regenerates every fixture

Cases
  a08d83_512_f64   bin/a08d83.data (726 grains; copied to tests/golden/a08d83.data as the input
                   fixture), 512 x 512, scale 1, fp64 -- SURVEY.md 4.4's known-answer configuration.
                   State after 15, 100 and 202 renderScene() calls.
  pack_64x48_<p>   seeded synthetic packing + perturbed populations + random grain kinematics on
                   a 64 x 48 lattice, 30 renderScene() calls, p in f64 / f32; full state stored.
  pack_256_<p>     same on 256 x 256 with ~200 grains, 100 (f64) / 30 (f32) calls; sampled state.
  cfg3/cfg4/cfg5   (--baseline[=name]) BASELINE.json configs[2..4] on the reference's own input files
                   (bin/a08d83.data at 2048^2 fp64; bin/a08_a4b4r18_7000.data at 4096^2 scale 2.7 fp32;
                   bin/50000.data at 8192^2 scale 2.6 fp64, 12 calls): hashes, node samples, 64 x 64 block sums
                   of rho and momentum, and the drift between two builds of the reference itself.
Large arrays are stored as a SHA-256 of their bytes (bit-exact check for the oracle and the
strict CUDA build) plus a strided sample (tolerance check for the default CUDA build).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle.refwrap import Reference  # noqa: E402
from util import perturbed_f, random_kinematics, small_packing  # noqa: E402
import make_sample  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SAMPLE_STRIDE = (16, 5, 7)   # nodes with x % 16 == 5 and y % 16 == 7


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def sample_nodes(a):
    s, ox, oy = SAMPLE_STRIDE
    return np.ascontiguousarray(a[ox::s, oy::s])


def snapshot(ref, tag, out, full=False):
    f, obst = ref.f(), ref.obst()
    out[f"{tag}_grains"] = ref.grains()[:, :9]
    out[f"{tag}_fhf"] = ref.fhf()
    out[f"{tag}_density"] = np.float64(ref.total_density())
    out[f"{tag}_f_sha256"] = np.array(sha(f))
    out[f"{tag}_obst_sha256"] = np.array(sha(obst))
    out[f"{tag}_act_sha256"] = np.array(sha(ref.act()))
    out[f"{tag}_solid_nodes"] = np.int64(((obst >= 0) & (obst < ref.n)).sum())
    if full:
        out[f"{tag}_f"] = f
        out[f"{tag}_obst"] = obst
        out[f"{tag}_act"] = ref.act()
    else:
        out[f"{tag}_f_sample"] = sample_nodes(f)
        out[f"{tag}_obst_sample"] = sample_nodes(obst)


def scalars_of(ref, out):
    sc = ref.scalars()
    for k, v in sc.items():
        out[f"scalar_{k}"] = np.float64(v) if isinstance(v, float) else np.int64(v)


def case_a08d83():
    src = "/root/reference/bin/a08d83.data"
    dst = os.path.join(GOLD, "a08d83.data")
    if not os.path.exists(dst):
        shutil.copyfile(src, dst)
    ref = Reference(512, 512, "1.", "f64")
    out = {}
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix="golden_"))
    try:
        n = ref.init(dst)
        assert n == 726
        scalars_of(ref, out)
        out["init_grains"] = ref.grains()
        out["init_obst_sha256"] = np.array(sha(ref.obst()))
        done = 0
        for upto in (15, 100, 202):
            ref.step(upto - done)
            done = upto
            snapshot(ref, f"s{upto}", out)
        cum, half = ref.verlet()
        out["verlet_cumul"], out["verlet_half"] = cum, half
        for name, lst in zip("BTLR", ref.wall_lists()):
            out[f"wall_{name}"] = lst
    finally:
        os.chdir(cwd)
    np.savez_compressed(os.path.join(GOLD, "a08d83_512_f64.npz"), **out)
    print("a08d83_512_f64: density after 15 / 202 steps", out["s15_density"], out["s202_density"])


def case_packing(lx, ly, prec, seed, steps, n_target, full):
    ref = Reference(lx, ly, "1.", prec)
    r, x, y = small_packing(lx, ly, 1.0, seed, n_target=n_target)
    tmp = tempfile.mkdtemp(prefix="golden_")
    path = os.path.join(tmp, "pack.data")
    # the reference only reads files: write the packing in its format (mm), read it back the
    # same way the tests will (values in the file are the fixture, not the numpy arrays)
    make_sample.write_sample(path, r * 1e3, x * 1e3, y * 1e3, comment=f"# small_packing seed={seed}")
    name = f"pack_{lx}x{ly}_{prec}"
    shutil.copyfile(path, os.path.join(GOLD, name + ".data"))
    out = {}
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        n = ref.init(path)
        scalars_of(ref, out)
        f0 = perturbed_f(lx, ly, seed + 1)
        v, w, a = random_kinematics(n, seed + 2, vmax=0.02)
        st = ref.grains()[:, :9].copy()
        st[:, 3:5], st[:, 5:6], st[:, 6:9] = v, w, a * 0.1
        ref.set_f(f0)
        ref.set_grain_state(st)
        # what the reference actually holds after rounding to its `real`
        out["start_state"] = ref.grains()[:, :9]
        if full:
            out["start_f"] = ref.f()
        else:
            out["start_f_seed"] = np.int64(seed + 1)
        ref.step(steps)
        out["steps"] = np.int64(steps)
        snapshot(ref, "end", out, full=full)
    finally:
        os.chdir(cwd)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, "n =", n, "density", out["end_density"])


def case_outputs():
    """The reference's own output files for 8000 renderScene() calls of the 64 x 48 packing from
    rest: DEM000000.dat (call 4000), DEM000001.dat + five VTK files (call 8000), stats.data.
    Stored verbatim under tests/golden/outputs_64x48/ with the state the writers saw."""
    ref = Reference(64, 48, "1.", "f64")
    out_dir = os.path.join(GOLD, "outputs_64x48")
    os.makedirs(out_dir, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="golden_out_")
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        n = ref.init(os.path.join(GOLD, "pack_64x48_f64.data"))
        # the reference's main() truncates stats.data and writes the header before the loop (:1867-1877)
        with open("stats.data", "w") as fh:
            fh.write("#1_t 2_xfront 3_xgrainmax 4_height 5_zmean 6_energie_x 7_energie_y "
                     "8_energie_teta 9_energie_cin 10_N0 11_N1 12_N2 13_N3 14_N4 15_N5 "
                     "16_energy_Potential 17_Strain_Energy 18_Frictional_Work "
                     "19_Internal_Friction 20_Inelastic_Collision 21_Slip "
                     "22_Rotational_Work\n")
        replay = {}
        done = 0
        for mark in (3998, 3999, 7998, 7999):
            ref.step(mark - done)
            done = mark
            replay[f"grains_{mark}"] = ref.grains()
            replay[f"fhf_{mark}"] = ref.fhf()
            cum, half = ref.verlet()
            replay[f"cumul_{mark}"], replay[f"half_{mark}"] = cum, half
            for nm, lst in zip("BTLR", ref.wall_lists()):
                replay[f"wall{nm}_{mark}"] = lst
            sc = ref.scalars()
            replay[f"d11_{mark}"] = np.array([sc[k] for k in ("dx", "dtLB", "dt", "dt2", "c", "Mgx", "Mdx", "Mby", "Mhy", "xG", "yG")])
            if mark in (3999, 7999):
                ref.step(1)
                done = mark + 1
                replay[f"grains_{done}"] = ref.grains()
                replay[f"fhf_{done}"] = ref.fhf()
                replay[f"diag_{done}"] = ref.grain_diag()
        np.savez_compressed(os.path.join(out_dir, "replay_states.npz"), **replay)
        assert done == 8000
        state = dict(f=ref.f(), obst=ref.obst(), grains=ref.grains(), diag=ref.grain_diag(), fhf=ref.fhf(),
                     density=np.float64(ref.total_density()))
        scalars_of(ref, state)
    finally:
        os.chdir(cwd)
    for name in sorted(os.listdir(tmp)):
        if name.endswith((".vtk", ".dat", ".data")) and not name.startswith("pack"):
            shutil.copyfile(os.path.join(tmp, name), os.path.join(out_dir, name))
    np.savez_compressed(os.path.join(out_dir, "state_8000.npz"), **state)
    print("outputs_64x48:", sorted(os.listdir(out_dir)))


def case_default_build():
    """The reference's DEFAULT build (lx = 7826, ly = 2325, fp64) on the synthetic 47 980-grain packing that
    stands in for bin/50000-test.data (BASELINE configs[0]; tools/make_sample.py, seed 12345): 37 calls of
    renderScene() = 4 LBM steps, one Verlet build.  Hashes only (the populations alone are 1.3 GB)."""
    n, r_min, r_max, width = make_sample.PRESETS["50000-test"]
    r, x, y = make_sample.packed_sample(n, r_min, r_max, width, seed=12345)
    tmp = tempfile.mkdtemp(prefix="golden_def_")
    path = os.path.join(tmp, "s50k.data")
    make_sample.write_sample(path, r, x, y, comment="# synthetic 50000-test seed=12345")
    ref = Reference(7826, 2325, "1.", "f64")
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        assert ref.init(path) == n
        out = {}
        scalars_of(ref, out)
        ref.step(37)
        g = ref.grains()
        out.update(steps=np.int64(37), grains_sha256=np.array(sha(g[:, :9])), fhf_sha256=np.array(sha(ref.fhf())),
                   obst_sha256=np.array(sha(ref.obst())), f_sha256=np.array(sha(ref.f())),
                   density=np.float64(ref.total_density()), grains_sample=g[::97, :9], fhf_sample=ref.fhf()[::97],
                   sample_sha256=np.array(hashlib.sha256(open(path, "rb").read()).hexdigest()))
    finally:
        os.chdir(cwd)
    np.savez_compressed(os.path.join(GOLD, "default_build_50000.npz"), **out)
    print("default_build_50000: density", out["density"])


# ---- BASELINE.json configs[2..4] on the reference's OWN input files, at the benchmarked lattice sizes ----
# name -> (input fixture, lx, ly, scale tag, precision, renderScene() call counts at which state is recorded,
#          stride of the node sample, whether the reference's own build-to-build spread is recorded too)
BASELINE_CASES = {
    # BASELINE configs[0]: the reference's DEFAULT build on its own 47 980-grain sample (grains above y = 232 mm lie outside
    # the lattice: legal, the loops clamp); 37 calls = 4 LBM steps and the O(N^2) Verlet build
    "cfg1_50000test_default_f64": ("50000-test.data", 7826, 2325, "1.", "f64", (13, 37), 64, False),
    "cfg3_a08d83_2048_f64": ("a08d83.data", 2048, 2048, "1.", "f64", (10, 100), 16, True),
    "cfg4_a08_7000_4096_f32": ("a08_a4b4r18_7000.data", 4096, 4096, "2.7", "f32", (2, 10, 30), 32, True),
    "cfg5_50000_8192_f64": ("50000.data", 8192, 8192, "2.6", "f64", (2, 12), 64, True),
}
BLOCK = 64   # rho / momentum are also stored as sums over BLOCK x BLOCK node blocks (covers every node)


def moments(f):
    """rho, jx, jy per node of a [lx][ly][9] array (src/main.c:1082-1090 without the collision)."""
    ex = np.array([0, -1, -1, -1, 0, 1, 1, 1, 0.0])
    ey = np.array([0, 1, 0, -1, -1, -1, 0, 1, 1.0])
    return f.sum(-1), f @ ex, f @ ey


def block_sums(a, b=BLOCK):
    lx, ly = a.shape
    return a[: lx // b * b, : ly // b * b].reshape(lx // b, b, ly // b, b).sum(axis=(1, 3))


def baseline_snapshot(ref, tag, out, stride):
    f, obst = ref.f(), ref.obst()
    g, fh = ref.grains()[:, :9], ref.fhf()
    out[f"{tag}_grains_sha256"], out[f"{tag}_fhf_sha256"] = np.array(sha(g)), np.array(sha(fh))
    gs = max(1, ref.n // 7000)   # every grain up to 7000, then a regular subset
    out["grain_stride"] = np.int64(gs)
    out[f"{tag}_grains"], out[f"{tag}_fhf"] = g[::gs].copy(), fh[::gs].copy()
    out[f"{tag}_density"] = np.float64(ref.total_density())   # serial sum in `real` (saturates in fp32, App. B #17)
    out[f"{tag}_density_f64"] = np.float64(f.sum(dtype=np.float64))
    out[f"{tag}_f_sha256"] = np.array(sha(f))
    out[f"{tag}_obst_sha256"] = np.array(sha(obst))
    out[f"{tag}_act_sha256"] = np.array(sha(ref.act()))
    out[f"{tag}_solid_nodes"] = np.int64(((obst >= 0) & (obst < ref.n)).sum())
    out[f"{tag}_f_sample"] = np.ascontiguousarray(f[5::stride, 7::stride])
    out[f"{tag}_obst_sample"] = np.ascontiguousarray(obst[5::stride, 7::stride])
    rho, jx, jy = moments(f)
    out[f"{tag}_rho_blocks"], out[f"{tag}_jx_blocks"], out[f"{tag}_jy_blocks"] = (block_sums(m) for m in (rho, jx, jy))
    return dict(g=g, fh=fh, rho=rho, jx=jx, jy=jy, obst=obst)


def case_baseline(name):
    """One BASELINE config on the reference's own input: the compiled reference (-O2 -ffp-contract=off, serial)
    from rest; optionally the same run with the reference's release flags (-Ofast ...) to record how far two
    builds of the reference itself drift apart at each mark (the honest bound for a differently-rounded build)."""
    fixture, lx, ly, scale, prec, marks, stride, spread = BASELINE_CASES[name]
    src = os.path.join(GOLD, fixture)
    if not os.path.exists(src):
        shutil.copyfile(os.path.join("/root/reference/bin", fixture), src)
    out = {"marks": np.array(marks, dtype=np.int64), "sample_stride": np.int64(stride), "block": np.int64(BLOCK),
           "fixture_sha256": np.array(hashlib.sha256(open(src, "rb").read()).hexdigest())}
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix="golden_"))
    try:
        ref = Reference(lx, ly, scale, prec)
        n = ref.init(src)
        scalars_of(ref, out)
        out["init_grains_sha256"] = np.array(sha(ref.grains()))
        out["init_obst_sha256"] = np.array(sha(ref.obst()))
        states, done = {}, 0
        for m in marks:
            ref.step(m - done)
            done = m
            states[m] = baseline_snapshot(ref, f"s{m}", out, stride)
            print(name, "mark", m, "density", out[f"s{m}_density_f64"], flush=True)
        cum, half = ref.verlet()
        out["verlet_pairs"] = np.int64(len(half))
        out["verlet_sha256"] = np.array(sha(np.concatenate([cum, half])))
        for nm, lst in zip("BTLR", ref.wall_lists()):
            out[f"wall_{nm}"] = lst
        del ref
        if spread:
            rel = Reference(lx, ly, scale, prec, omp=False, release=True)
            assert rel.init(src) == n
            done = 0
            for m in marks:
                rel.step(m - done)
                done = m
                a = states[m]
                g, fh = rel.grains()[:, :9], rel.fhf()
                rho, jx, jy = moments(rel.f())
                rel_err = lambda x, y: float(np.abs(x - y).max() / max(np.abs(y).max(), 1e-300))
                out[f"s{m}_spread"] = np.array([rel_err(g[:, 0:3], a["g"][:, 0:3]), rel_err(g[:, 3:6], a["g"][:, 3:6]),
                                                rel_err(fh, a["fh"]), rel_err(rho, a["rho"]),
                                                float(np.abs(jx - a["jx"]).max() / max(np.abs(a["jx"]).max(), 1e-3)),
                                                float(np.abs(jy - a["jy"]).max() / max(np.abs(a["jy"]).max(), 1e-3)),
                                                float((rel.obst() != a["obst"]).sum())])
                print(name, "mark", m, "release-vs-O2 spread [x v fhf rho jx jy obst]", out[f"s{m}_spread"], flush=True)
            del rel
    finally:
        os.chdir(cwd)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, "n =", n, "written", os.path.getsize(os.path.join(GOLD, name + ".npz")) // 1024, "KB")


def main():
    os.makedirs(GOLD, exist_ok=True)
    if "--default-build" in sys.argv:
        case_default_build()
        return
    picked = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--baseline=")]
    if picked or "--baseline" in sys.argv:
        for name in (picked or BASELINE_CASES):
            case_baseline(name)
        return
    case_outputs()
    case_a08d83()
    case_packing(64, 48, "f64", 3, 30, None, True)
    case_packing(64, 48, "f32", 3, 30, None, True)
    case_packing(256, 256, "f64", 41, 100, 200, False)
    case_packing(256, 256, "f32", 41, 30, 200, False)


if __name__ == "__main__":
    main()
