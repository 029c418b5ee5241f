"""ctypes front-end for the plain-C restatement (oracle/lbmdem_oracle.c).

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's CPU legs).
Same method names as oracle.refwrap.Reference so that tests can treat both alike.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build_oracle())
    return _lib


class Oracle:
    def __init__(self, lx, ly, scale=1.0, prec="f64"):
        assert prec in ("f64", "f32")
        self.L = _load()
        self.sfx = prec
        self.lx, self.ly, self.scale = lx, ly, float(scale)
        self.real_bytes = 8 if prec == "f64" else 4
        fn = self._fn("oracle_create", [C.c_int, C.c_int, C.c_double], C.c_void_p)
        self.o = fn(lx, ly, float(scale))
        if not self.o:
            raise MemoryError("oracle_create failed")
        self.n = 0

    def _fn(self, name, argtypes, restype=None):
        f = getattr(self.L, f"{name}_{self.sfx}")
        f.argtypes = argtypes
        f.restype = restype
        return f

    def close(self):
        if self.o:
            self._fn("oracle_destroy", [C.c_void_p])(self.o)
            self.o = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- life cycle -------------------------------------------------------------------
    def init(self, sample_path: str) -> int:
        n = self._fn("oracle_read_sample", [C.c_void_p, C.c_char_p], C.c_int)(self.o, os.fsencode(sample_path))
        if n < 0:
            raise IOError(f"oracle_read_sample({sample_path}) -> {n}")
        self.n = n
        self._fn("oracle_init", [C.c_void_p])(self.o)
        return n

    def init_arrays(self, r, x1, x2) -> int:
        r, x1, x2 = (np.ascontiguousarray(a, dtype=np.float64) for a in (r, x1, x2))
        n = self._fn("oracle_set_grains", [C.c_void_p, C.c_int, _dp, _dp, _dp], C.c_int)(self.o, len(r), r, x1, x2)
        self.n = n
        self._fn("oracle_init", [C.c_void_p])(self.o)
        return n

    def set_lid(self, uw):
        self._fn("oracle_set_lid", [C.c_void_p, C.c_double])(self.o, float(uw))

    def set_vib(self, vib):
        self._fn("oracle_set_vib", [C.c_void_p, C.c_int])(self.o, int(vib))

    def set_dtt(self, dtt):
        self._fn("oracle_set_dtt", [C.c_void_p, C.c_double])(self.o, float(dtt))

    def set_angleG(self, a):
        self._fn("oracle_set_angleG", [C.c_void_p, C.c_double])(self.o, float(a))

    def step(self, n=1):
        self._fn("oracle_step", [C.c_void_p, C.c_long])(self.o, n)

    def lbm_step(self):
        self._fn("oracle_lbm_step", [C.c_void_p])(self.o)

    def phase(self, name):
        """name in reinit_obst_density, obst_construction, collision_streaming, forces_fluid, init_verlet"""
        self._fn(f"oracle_{name}", [C.c_void_p])(self.o)

    # -- scalars ----------------------------------------------------------------------
    def scalars(self) -> dict:
        d = (C.c_double * 11)()
        l = (C.c_long * 4)()
        self._fn("oracle_get_scalars", [C.c_void_p, C.c_void_p, C.c_void_p])(self.o, d, l)
        keys = ["dx", "dtLB", "dt", "dt2", "c", "Mgx", "Mdx", "Mby", "Mhy", "xG", "yG"]
        out = dict(zip(keys, list(d)))
        out.update(npDEM=l[0], nbsteps=l[1], nFile=l[2], nbgrains=l[3])
        return out

    def set_nbsteps(self, n):
        self._fn("oracle_set_nbsteps", [C.c_void_p, C.c_long])(self.o, n)

    def total_density(self) -> float:
        return self._fn("oracle_total_density", [C.c_void_p], C.c_double)(self.o)

    # -- arrays -----------------------------------------------------------------------
    def _get(self, name, shape, dtype=np.float64):
        a = np.empty(shape, dtype=dtype)
        self._fn(name, [C.c_void_p, _dp if dtype == np.float64 else _ip])(self.o, a)
        return a

    def _set(self, name, a, dtype=np.float64):
        a = np.ascontiguousarray(a, dtype=dtype)
        self._fn(name, [C.c_void_p, _dp if dtype == np.float64 else _ip])(self.o, a)

    def f(self):
        return self._get("oracle_get_f", (self.lx, self.ly, 9))

    def set_f(self, a):
        assert a.shape == (self.lx, self.ly, 9)
        self._set("oracle_set_f", a)

    def delta(self):
        return self._get("oracle_get_delta", (self.lx, self.ly, 9))

    def obst(self):
        return self._get("oracle_get_obst", (self.lx, self.ly), np.int32)

    def set_obst(self, a):
        self._set("oracle_set_obst", a, np.int32)

    def act(self):
        return self._get("oracle_get_act", (self.lx, self.ly), np.int32)

    def grains(self):
        return self._get("oracle_get_grains", (self.n, 13))

    def set_grain_state(self, a):
        a = np.asarray(a)
        assert a.shape == (self.n, 9)
        self._set("oracle_set_grain_state", a)

    def grain_diag(self):
        return self._get("oracle_get_grain_diag", (self.n, 17))

    def fhf(self):
        return self._get("oracle_get_fhf", (self.n, 3))

    def set_fhf(self, a):
        self._set("oracle_set_fhf", a)

    def verlet(self):
        cumul = np.zeros(self.n, dtype=np.int32)
        neigh = np.zeros(64 * self.n + 64, dtype=np.int32)
        tot = self._fn("oracle_get_verlet", [C.c_void_p, _ip, _ip, C.c_int], C.c_int)(self.o, cumul, neigh, neigh.size)
        assert tot >= 0
        return cumul, neigh[:tot].copy()

    def wall_lists(self):
        cnt = np.zeros(4, dtype=np.int32)
        ls = [np.zeros(max(self.n, 1), dtype=np.int32) for _ in range(4)]
        self._fn("oracle_get_wall_lists", [C.c_void_p, _ip, _ip, _ip, _ip, _ip])(self.o, cnt, *ls)
        return [l[:c].copy() for l, c in zip(ls, cnt)]

    # -- timers -----------------------------------------------------------------------
    def time_coupled(self, n_dem_steps):
        nl = C.c_long()
        t = self._fn("oracle_time_coupled", [C.c_void_p, C.c_long, C.c_void_p], C.c_double)(self.o, n_dem_steps, C.byref(nl))
        return t, nl.value

    def time_lbm(self, n_lbm_steps):
        return self._fn("oracle_time_lbm", [C.c_void_p, C.c_long], C.c_double)(self.o, n_lbm_steps)
