"""ctypes front-end for the compiled reference (oracle/_ref/libref_*.so).

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's CPU legs).
See oracle/ref_shim.c for what each entry point calls in the reference.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class Reference:
    """One compiled configuration of the reference.  State is global inside the
    library (the reference is written that way), so use one instance per .so."""

    def __init__(self, lx, ly, scale="1.", prec="f64", omp=False, release=False):
        path = _build.build_ref(lx, ly, scale, prec, omp, release)
        if path is None or not os.path.exists(path):
            raise FileNotFoundError(
                f"reference library for {lx}x{ly} scale={scale} {prec} is not built and "
                f"{_build.REF_ROOT} is absent")
        self.path = path
        self.lib = L = C.CDLL(path)
        L.ref_init.argtypes = [C.c_char_p]
        L.ref_init.restype = C.c_int
        L.ref_step.argtypes = [C.c_long]
        L.ref_total_density.restype = C.c_double
        L.ref_time_coupled.argtypes = [C.c_long, C.POINTER(C.c_long)]
        L.ref_time_coupled.restype = C.c_double
        L.ref_time_lbm.argtypes = [C.c_long]
        L.ref_time_lbm.restype = C.c_double
        L.ref_set_nbsteps.argtypes = [C.c_long]
        for name in ("ref_get_f", "ref_set_f", "ref_get_delta", "ref_get_grains",
                     "ref_set_grain_state", "ref_get_grain_diag", "ref_get_fhf", "ref_set_fhf"):
            getattr(L, name).argtypes = [_dp]
        for name in ("ref_get_obst", "ref_set_obst", "ref_get_act"):
            getattr(L, name).argtypes = [_ip]
        L.ref_get_verlet.argtypes = [_ip, _ip, C.c_int]
        L.ref_get_verlet.restype = C.c_int
        L.ref_get_wall_lists.argtypes = [_ip, _ip, _ip, _ip, _ip]
        lx_, ly_, sc, rb = C.c_int(), C.c_int(), C.c_double(), C.c_int()
        L.ref_config(C.byref(lx_), C.byref(ly_), C.byref(sc), C.byref(rb))
        self.lx, self.ly, self.scale, self.real_bytes = lx_.value, ly_.value, sc.value, rb.value
        assert (self.lx, self.ly) == (lx, ly)
        self.n = 0

    # -- life cycle -------------------------------------------------------------------
    def init(self, sample_path: str) -> int:
        self.n = self.lib.ref_init(os.fsencode(sample_path))
        if self.n < 0:
            raise MemoryError("reference allocation failed")
        return self.n

    def set_vib(self, vib):
        self.lib.ref_set_vib(int(vib))

    def set_dtt(self, dtt):
        self.lib.ref_set_dtt.argtypes = [C.c_double]
        self.lib.ref_set_dtt(float(dtt))

    def set_angleG(self, a):
        self.lib.ref_set_angleG.argtypes = [C.c_double]
        self.lib.ref_set_angleG(float(a))

    def step(self, n=1):
        self.lib.ref_step(n)

    def lbm_step(self):
        self.lib.ref_lbm_step()

    # -- scalars ----------------------------------------------------------------------
    def scalars(self) -> dict:
        d = (C.c_double * 11)()
        l = (C.c_long * 4)()
        self.lib.ref_get_scalars(d, l)
        keys = ["dx", "dtLB", "dt", "dt2", "c", "Mgx", "Mdx", "Mby", "Mhy", "xG", "yG"]
        out = dict(zip(keys, list(d)))
        out.update(npDEM=l[0], nbsteps=l[1], nFile=l[2], nbgrains=l[3])
        return out

    def total_density(self) -> float:
        return self.lib.ref_total_density()

    # -- arrays -----------------------------------------------------------------------
    def f(self) -> np.ndarray:
        a = np.empty((self.lx, self.ly, 9))
        self.lib.ref_get_f(a)
        return a

    def set_f(self, a):
        self.lib.ref_set_f(np.ascontiguousarray(a, dtype=np.float64))

    def delta(self) -> np.ndarray:
        a = np.empty((self.lx, self.ly, 9))
        self.lib.ref_get_delta(a)
        return a

    def obst(self) -> np.ndarray:
        a = np.empty((self.lx, self.ly), dtype=np.int32)
        self.lib.ref_get_obst(a)
        return a

    def set_obst(self, a):
        self.lib.ref_set_obst(np.ascontiguousarray(a, dtype=np.int32))

    def act(self) -> np.ndarray:
        a = np.empty((self.lx, self.ly), dtype=np.int32)
        self.lib.ref_get_act(a)
        return a

    def grains(self) -> np.ndarray:
        """[N,13]: x1 x2 x3 v1 v2 v3 a1 a2 a3 r m It rLB"""
        a = np.empty((self.n, 13))
        self.lib.ref_get_grains(a)
        return a

    def set_grain_state(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == (self.n, 9)
        self.lib.ref_set_grain_state(a)

    def grain_diag(self) -> np.ndarray:
        a = np.empty((self.n, 17))
        self.lib.ref_get_grain_diag(a)
        return a

    def fhf(self) -> np.ndarray:
        a = np.empty((self.n, 3))
        self.lib.ref_get_fhf(a)
        return a

    def set_fhf(self, a):
        self.lib.ref_set_fhf(np.ascontiguousarray(a, dtype=np.float64))

    def verlet(self):
        """(cumul[N] end offsets, neighbours[total]) of the half list (src/main.c:1519-1543)."""
        cumul = np.zeros(self.n, dtype=np.int32)
        neigh = np.zeros(max(6 * self.n, 1), dtype=np.int32)
        tot = self.lib.ref_get_verlet(cumul, neigh, neigh.size)
        assert tot >= 0
        return cumul, neigh[:tot].copy()

    def wall_lists(self):
        cnt = np.zeros(4, dtype=np.int32)
        ls = [np.zeros(max(self.n, 1), dtype=np.int32) for _ in range(4)]
        self.lib.ref_get_wall_lists(cnt, *ls)
        return [l[:c].copy() for l, c in zip(ls, cnt)]  # B T L R

    # -- timers -----------------------------------------------------------------------
    def time_coupled(self, n_dem_steps):
        nl = C.c_long()
        t = self.lib.ref_time_coupled(n_dem_steps, C.byref(nl))
        return t, nl.value

    def time_lbm(self, n_lbm_steps):
        return self.lib.ref_time_lbm(n_lbm_steps)

    def set_omp_threads(self, n):
        self.lib.ref_set_omp_threads(int(n))

    def omp_threads(self):
        return self.lib.ref_omp_threads()
