"""Shared helpers for the tests (CPU and GPU)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W = np.array([4 / 9, 1 / 36, 1 / 9, 1 / 36, 1 / 9, 1 / 36, 1 / 9, 1 / 36, 1 / 9])
EX = np.array([0, -1, -1, -1, 0, 1, 1, 1, 0])
EY = np.array([0, 1, 0, -1, -1, -1, 0, 1, 1])

# constants of src/main.c:74-118 that the tests need
RHO_MOY, TAU, NU = 1000.0, 0.504, 1e-6
DEM_CONST = dict(kg=1.6e6, kt=1.0e6, km=3e6, ktm=2e6, nug=6.4e1, num=8.7e1, numb=8.7e1, nugt=5e-1,
                 mu=.5317, mum=.466, mumb=.466, murf=0.01, freq=5.0, amp=4.e-4, t=0.0, distVerlet=5e-4)


def small_packing(lx, ly, scale=1.0, seed=0, n_target=None, overlap=0.02, wall_touch=True):
    """A few dozen discs (metres) inside the lattice extent 1e-4*lx/scale x 1e-4*ly/scale:
    jittered triangular packing with slight overlaps (contacts from step 0), some discs
    touching the bottom/left DEM walls and the wall ring of the lattice."""
    rng = np.random.default_rng(seed)
    dx = (1.0 / scale) * (1e-3 * lx / 10) / (lx - 1)
    W_, H_ = dx * (lx - 1), dx * (ly - 1)
    rmean = 7.0 * dx if n_target is None else max(4.0 * dx, 0.5 * np.sqrt(0.5 * W_ * H_ / n_target))
    pitch = 2 * rmean * (1 - overlap)
    xs, ys, rs = [], [], []
    row = 0
    y = rmean * (0.98 if wall_touch else 1.3)
    while y + rmean < 0.8 * H_:
        x = rmean * (0.98 if wall_touch else 1.3) + (row % 2) * 0.5 * pitch
        while x + rmean < 0.9 * W_:
            xs.append(x + rng.uniform(-0.01, 0.01) * rmean)
            ys.append(y + rng.uniform(-0.01, 0.01) * rmean)
            rs.append(rmean * rng.uniform(0.93, 1.0))
            x += pitch
        y += pitch * np.sqrt(3) / 2
        row += 1
    return np.array(rs), np.array(xs), np.array(ys)


def random_kinematics(n, seed=1, vmax=0.05, wmax=20.0, amax=50.0):
    rng = np.random.default_rng(seed)
    v = rng.uniform(-vmax, vmax, size=(n, 2))
    w = rng.uniform(-wmax, wmax, size=(n, 1))
    a = rng.uniform(-amax, amax, size=(n, 3))
    return v, w, a


def perturbed_f(lx, ly, seed=2, amp=0.02):
    rng = np.random.default_rng(seed)
    return W[None, None, :] * (1.0 + amp * rng.uniform(-1, 1, size=(lx, ly, 9)))


def scale_fhf(fhf_unscaled, dx, prec="f64"):
    """fhf scaling of src/main.c:1329-1331 with the reference's promotions: the numerator is a
    `real` product chain, `(tau - 0.5)` is double, so the factor and the final product are
    evaluated in double and rounded to `real` on store."""
    real = np.float64 if prec == "f64" else np.float32
    r = real
    num = r(r(r(r(RHO_MOY) * r(9)) * r(NU)) * r(NU))
    den = np.float64(r(dx)) * (np.float64(r(TAU)) - 0.5) * (np.float64(r(TAU)) - 0.5)
    k12 = np.float64(num) / den
    num3 = r(r(r(r(r(dx) * r(RHO_MOY)) * r(9)) * r(NU)) * r(NU))
    k3 = np.float64(num3) / den
    out = np.asarray(fhf_unscaled, dtype=real).astype(np.float64)
    out[:, 0] *= k12
    out[:, 1] *= k12
    out[:, 2] *= k3
    return out.astype(real).astype(np.float64)


def load_hostcheck():
    import importlib.util
    spec = importlib.util.spec_from_file_location("hostcheck_build", os.path.join(ROOT, "tests", "hostcheck", "build.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    lib = C.CDLL(m.build())
    dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    for sfx in ("f64", "f32"):
        fn = getattr(lib, f"hc_lbm_step_{sfx}")
        fn.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp, ip, dp, ip, ip, dp]
        fn.restype = C.c_int
        fn = getattr(lib, f"hc_raster_tiles_check_{sfx}")
        fn.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, C.c_int, C.c_int, np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")]
        fn.restype = C.c_int
        fn = getattr(lib, f"hc_dem_step_{sfx}")
        fn.argtypes = [C.c_int, dp, C.c_int, dp, dp, dp, C.c_int, ip, ip, C.c_int, ip]
        fn.restype = C.c_int
    return lib


def mirrored_state(st, Mdx, Mhy):
    """The kinematic state [n][9] point-mirrored through the centre of the box (0, Mdx) x (0, Mhy): a packing that
    leans on the bottom / left DEM walls then leans on the TOP / RIGHT ones (force_WallT / force_WallR,
    src/main.c:846-887, :923-951), with the same relative geometry between grains."""
    out = np.array(st, dtype=np.float64, copy=True)
    out[:, 0] = Mdx - out[:, 0]
    out[:, 1] = Mhy - out[:, 1]
    out[:, 3:5] = -out[:, 3:5]
    out[:, 6:8] = -out[:, 6:8]
    return out
