"""CPU tests of the plain-C host side (2d-lbm-dem_b200/host): the VTK writer and the replay of the
contact diagnostics + write_DEM, against FILES WRITTEN BY THE COMPILED REFERENCE
(tests/golden/outputs_64x48, tools/make_golden.py case_outputs)."""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "outputs_64x48")

import lbmdem_gpu as G

dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def host():
    spec = importlib.util.spec_from_file_location("host_build", os.path.join(ROOT, "2d-lbm-dem_b200", "host", "build.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    L = C.CDLL(m.build_host_lib())
    L.lbmdem_write_vtk_frame.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, fp, fp, fp, fp, fp]
    L.lbmdem_diag_create.argtypes = [C.c_int, C.c_void_p]
    L.lbmdem_diag_create.restype = C.c_void_p
    L.lbmdem_diag_destroy.argtypes = [C.c_void_p]
    L.lbmdem_diag_pass.argtypes = [C.c_void_p, dp, dp, dp, ip, ip, C.c_int, ip, dp]
    L.lbmdem_diag_get.argtypes = [C.c_void_p, dp]
    L.lbmdem_write_dem.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_long, dp, dp, dp, C.c_void_p]
    L.lbmdem_write_stats_header.argtypes = [C.c_char_p]
    return L


def test_vtk_frame_is_byte_identical_to_the_reference(host, tmp_path):
    st = np.load(os.path.join(OUT, "state_8000.npz"))
    f, obst, g, diag = st["f"], st["obst"], st["grains"], st["diag"]
    lx, ly = obst.shape
    n = len(g)
    # write_vtk's field definitions (src/main.c:284-323), [y][x] order
    solid = ((obst >= 0) & (obst < n)).T
    idx = np.where(solid, obst.T, 0)
    gp = np.where(solid, diag[idx, 0], -1.0).astype(np.float32)
    gv = np.zeros((ly, lx, 3), dtype=np.float32)
    ga = np.zeros((ly, lx, 3), dtype=np.float32)
    gv[..., 0], gv[..., 1] = np.where(solid, g[idx, 3], 0), np.where(solid, g[idx, 4], 0)
    ga[..., 0], ga[..., 1] = np.where(solid, g[idx, 6], 0), np.where(solid, g[idx, 7], 0)
    ex = np.array([0, -1, -1, -1, 0, 1, 1, 1, 0.0])
    ey = np.array([0, 1, 0, -1, -1, -1, 0, 1, 1.0])
    acc, jx, jy = (np.zeros((lx, ly), dtype=np.float32) for _ in range(3))
    for q in range(9):   # the reference accumulates into float fields one population at a time
        acc = (acc.astype(np.float64) + f[:, :, q]).astype(np.float32)
        jx = (jx.astype(np.float64) + f[:, :, q] * ex[q]).astype(np.float32)
        jy = (jy.astype(np.float64) + f[:, :, q] * ey[q]).astype(np.float32)
    fpv = np.where(solid, 0.0, ((1.0 / 3.0) * 1000.0 * (acc.astype(np.float64) - 1.0)).astype(np.float32).T).astype(np.float32)
    fv = np.zeros((ly, lx, 3), dtype=np.float32)
    fv[..., 0], fv[..., 1] = np.where(solid, 0, jx.T), np.where(solid, 0, jy.T)
    arrs = [np.ascontiguousarray(a) for a in (gp, gv, ga, fpv, fv)]
    assert host.lbmdem_write_vtk_frame(os.fsencode(str(tmp_path)), 0, lx, ly, *arrs) == 0
    for name in ("grain_pressure", "grain_velocity", "grain_acceleration", "fluid_pressure", "fluid_velocity"):
        mine = open(tmp_path / f"{name}_000000.vtk", "rb").read()
        ref = open(os.path.join(OUT, f"{name}_000000.vtk"), "rb").read()
        assert len(mine) == len(ref), name
        assert mine == ref, f"{name}: first differing byte {next(i for i, (a, b) in enumerate(zip(mine, ref)) if a != b)}"


def _full_lists(cumul, half, n, cap=32):
    cnt = np.zeros(n, dtype=np.int32)
    nbr = np.full((n, cap), -1, dtype=np.int32)
    pairs, start = [], 0
    for i in range(n):
        end = cumul[i] if i < n - 1 else start     # the reference never writes cumul[n-1]
        pairs += [(i, int(j)) for j in half[start:end]]
        start = end
    adj = [[] for _ in range(n)]
    for i, j in pairs:
        adj[i].append(j)
        adj[j].append(i)
    for i in range(n):
        row = sorted(adj[i])
        cnt[i] = len(row)
        nbr[i, :len(row)] = row
    return cnt, nbr


def _mid_state(g, dt, dt2):
    """kick-drift of src/main.c:1748-1753 with its association: x = (x + dt*v) + (dt2*a)/2"""
    x, v, a = g[:, 0:3], g[:, 3:6], g[:, 6:9]
    return np.ascontiguousarray(np.hstack([(x + dt * v) + (dt2 * a) / 2.0, v + (dt * a) / 2.0]))


def test_contact_diagnostics_and_dem_files_match_the_reference(host, tmp_path):
    rp = np.load(os.path.join(OUT, "replay_states.npz"))
    n = len(rp["grains_3998"])
    params = G.default_params(lx=64, ly=48)
    d = host.lbmdem_diag_create(n, C.addressof(params))
    assert host.lbmdem_write_stats_header(os.fsencode(str(tmp_path))) == 0
    for nfile, first in ((0, 3998), (1, 7998)):
        for mark in (first, first + 1):          # the two calls before the output: carry-over of pf, pft, pff
            g = rp[f"grains_{mark}"]
            d11 = np.ascontiguousarray(rp[f"d11_{mark}"])
            cnt, nbr = _full_lists(rp[f"cumul_{mark}"], rp[f"half_{mark}"], n)
            wf = np.zeros(n, dtype=np.int32)
            for bit, nm in enumerate("BTLR"):
                wf[rp[f"wall{nm}_{mark}"]] |= 1 << bit
            # VerletWall has moved the right/top walls by then (dtt = 0): use the scalars stored with the state
            host.lbmdem_diag_pass(d, _mid_state(g, d11[2], d11[3]), np.ascontiguousarray(g),
                                  np.ascontiguousarray(rp[f"fhf_{mark}"]), cnt, np.ascontiguousarray(nbr), 32, wf, d11)
        after = first + 2
        diag = np.empty((n, 17))
        host.lbmdem_diag_get(d, diag)
        ref = rp[f"diag_{after}"]
        for col in (0, 1, 2, 3, 4, 6, 8, 9, 10, 11, 12, 13, 14, 15, 16):   # fm, ifr are set by write_DEM itself
            assert np.array_equal(diag[:, col], ref[:, col]), f"diagnostic column {col} before file {nfile}"
        d11 = np.ascontiguousarray(rp[f"d11_{first + 1}"])
        assert host.lbmdem_write_dem(d, os.fsencode(str(tmp_path)), nfile, after, np.ascontiguousarray(rp[f"grains_{after}"]),
                                     np.ascontiguousarray(rp[f"fhf_{after}"]), d11, None) == 0
        assert open(tmp_path / f"DEM{nfile:06d}.dat").read() == open(os.path.join(OUT, f"DEM{nfile:06d}.dat")).read()
    assert open(tmp_path / "stats.data").read() == open(os.path.join(OUT, "stats.data")).read()
    host.lbmdem_diag_destroy(d)
