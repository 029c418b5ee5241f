"""ctypes front-end of liblbmdem_gpu.so (include/lbmdem_gpu.h).

Host-side mirror of the reference's driver for the hot path: the method names follow the
reference's functions (renderScene -> step, the LBM part of it -> lbm_step, initVerlet +
VerletWall -> build_verlet, read_sample + main()'s set-up -> init) and are the same as those of
the test oracles (oracle/oraclewrap.py, oracle/refwrap.py), so parity tests read alike.

There is no CPU path: the library fails at create() when no sm_100 GPU is visible, and this
module raises if the shared library has not been built (python 2d-lbm-dem_b200/build.py).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LBMDEM_LIB") or os.path.join(HERE, "liblbmdem_gpu.so")   # LBMDEM_LIB: tuning variants

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")

ERRORS = {-1: "EINVAL", -2: "ECUDA", -3: "ENOMEM", -4: "ESTATE", -5: "EIO", -6: "ECAP", -7: "ENCCL", -8: "ERANGE"}


class LbmdemError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"lbmdem {ERRORS.get(code, code)}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [
        ("lx", C.c_int), ("ly", C.c_int), ("scale", C.c_double), ("single_precision", C.c_int),
        ("device", C.c_int), ("rank", C.c_int), ("nranks", C.c_int),
        ("tau", C.c_double), ("nu", C.c_double), ("rho_moy", C.c_double), ("reductionR", C.c_double),
        ("s2", C.c_double), ("s3", C.c_double), ("s5", C.c_double), ("s7", C.c_double), ("s8", C.c_double),
        ("s9", C.c_double),
        ("G", C.c_double), ("angleG", C.c_double), ("kg", C.c_double), ("kt", C.c_double), ("km", C.c_double),
        ("ktm", C.c_double), ("nug", C.c_double), ("num", C.c_double), ("numb", C.c_double), ("nugt", C.c_double),
        ("mu", C.c_double), ("mum", C.c_double), ("mumb", C.c_double), ("murf", C.c_double),
        ("rscale", C.c_double), ("distVerlet", C.c_double), ("dtt", C.c_double), ("iterDEM", C.c_double),
        ("freq", C.c_double), ("amp", C.c_double), ("rhoS", C.c_double),
        ("UpdateVerlet", C.c_long), ("stepFilm", C.c_long),
        ("lid_u", C.c_double), ("strict_fp", C.c_int), ("kernel", C.c_int), ("neighbour_capacity", C.c_int), ("vib", C.c_int),
    ]


_lib = None


def load_library():
    """Loads liblbmdem_gpu.so and declares every entry point of include/lbmdem_gpu.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} is not built: run `python 2d-lbm-dem_b200/build.py` (needs nvcc)")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    sig = {
        "lbmdem_default_params": ([C.POINTER(Params)], C.c_int),
        "lbmdem_sizeof_params": ([], C.c_int),
        "lbmdem_create": ([C.POINTER(Params), C.POINTER(vp)], C.c_int),
        "lbmdem_destroy": ([vp], None),
        "lbmdem_last_error": ([vp], C.c_char_p),
        "lbmdem_load_sample": ([vp, C.c_char_p], C.c_int),
        "lbmdem_set_grains": ([vp, C.c_int, _dp, _dp, _dp], C.c_int),
        "lbmdem_step": ([vp, C.c_long], C.c_int),
        "lbmdem_step_capture": ([vp, _dp], C.c_int),
        "lbmdem_lbm_step": ([vp], C.c_int),
        "lbmdem_lbm_steps": ([vp, C.c_long], C.c_int),
        "lbmdem_build_verlet": ([vp], C.c_int),
        "lbmdem_get_scalars": ([vp, vp, vp], C.c_int),
        "lbmdem_set_nbsteps": ([vp, C.c_long], C.c_int),
        "lbmdem_get_strip": ([vp, C.POINTER(C.c_int), C.POINTER(C.c_int)], C.c_int),
        "lbmdem_total_density": ([vp, C.POINTER(C.c_double)], C.c_int),
        "lbmdem_get_f": ([vp, _dp], C.c_int),
        "lbmdem_set_f": ([vp, _dp], C.c_int),
        "lbmdem_get_obst": ([vp, _ip], C.c_int),
        "lbmdem_set_obst": ([vp, _ip], C.c_int),
        "lbmdem_get_act": ([vp, _ip], C.c_int),
        "lbmdem_get_grains": ([vp, _dp], C.c_int),
        "lbmdem_set_grain_state": ([vp, _dp], C.c_int),
        "lbmdem_get_fhf": ([vp, _dp], C.c_int),
        "lbmdem_set_fhf": ([vp, _dp], C.c_int),
        "lbmdem_get_verlet": ([vp, _ip, _ip, C.c_int, _ip], C.c_int),
        "lbmdem_get_fields": ([vp, vp, _fp, _fp, _fp, _fp, _fp], C.c_int),
        "lbmdem_save_state": ([vp, C.c_char_p], C.c_int),
        "lbmdem_load_state": ([vp, C.c_char_p], C.c_int),
        "lbmdem_step_host": ([vp, vp, C.c_long, vp, vp, vp], C.c_int),
        "lbmdem_step_host_f32": ([vp, vp, C.c_long, vp, vp, vp], C.c_int),
        "lbmdem_host_alloc": ([C.c_size_t, C.POINTER(vp)], C.c_int),
        "lbmdem_host_free": ([vp], C.c_int),
        "lbmdem_nccl_unique_id": ([vp], C.c_int),
        "lbmdem_attach_nccl": ([vp, vp], C.c_int),
        "lbmdem_get_kernel_timer": ([vp, C.POINTER(C.c_double), C.POINTER(C.c_long), C.POINTER(C.c_long)], C.c_int),
        "lbmdem_reset_kernel_timer": ([vp, C.c_int], C.c_int),
        "lbmdem_get_list_counts": ([vp, C.POINTER(C.c_long)], C.c_int),
        "lbmdem_get_share": ([vp, C.POINTER(C.c_int), C.POINTER(C.c_int)], C.c_int),
        "lbmdem_step_host_share": ([vp, vp, C.c_long, vp, vp, vp], C.c_int),
        "lbmdem_step_host_share_f32": ([vp, vp, C.c_long, vp, vp, vp], C.c_int),
        "lbmdem_local_group_create": ([C.c_int, C.POINTER(vp)], C.c_int),
        "lbmdem_attach_local": ([vp, vp], C.c_int),
        "lbmdem_local_group_destroy": ([vp], None),
        "lbmdem_state_checksum": ([vp, C.POINTER(C.c_ulonglong)], C.c_int),
        "lbmdem_stream": ([vp], vp),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res
    L._declared = sorted(sig)
    if L.lbmdem_sizeof_params() != C.sizeof(Params):
        raise ImportError("lbmdem_params layout mismatch between lbmdem_gpu.py and liblbmdem_gpu.so")
    _lib = L
    return L


def default_params(**over) -> Params:
    L = load_library()
    p = Params()
    L.lbmdem_default_params(C.byref(p))
    for k, v in over.items():
        if not hasattr(p, k):
            raise AttributeError(f"lbmdem_params has no field {k}")
        setattr(p, k, v)
    return p


def nccl_unique_id() -> bytes:
    L = load_library()
    buf = C.create_string_buffer(128)
    rc = L.lbmdem_nccl_unique_id(buf)
    if rc:
        raise LbmdemError(rc, L.lbmdem_last_error(None).decode())
    return buf.raw


class Solver:
    """One context = one GPU (one x-strip of the lattice when nranks > 1)."""

    def __init__(self, lx, ly, scale=1.0, prec="f64", **over):
        assert prec in ("f64", "f32")
        self.L = load_library()
        self.params = default_params(lx=lx, ly=ly, scale=float(scale), single_precision=int(prec == "f32"), **over)
        self.lx, self.ly, self.scale, self.prec = lx, ly, float(scale), prec
        self.real_bytes = 8 if prec == "f64" else 4
        h = C.c_void_p()
        rc = self.L.lbmdem_create(C.byref(self.params), C.byref(h))
        if rc:
            raise LbmdemError(rc, self.L.lbmdem_last_error(None).decode())
        self.h = h
        self.n = 0
        a, b = C.c_int(), C.c_int()
        self._ck(self.L.lbmdem_get_strip(self.h, C.byref(a), C.byref(b)))
        self.xlo, self.xhi = a.value, b.value
        self.nx = self.xhi - self.xlo

    def _ck(self, rc):
        if rc < 0:
            raise LbmdemError(rc, self.L.lbmdem_last_error(self.h).decode())
        return rc

    def close(self):
        for k in [k for k in self.__dict__ if k.startswith("_hostbufs")]:
            self.__dict__.pop(k, None)
        for p, _ in self.__dict__.pop("_pinned_bufs", {}).values():
            self.L.lbmdem_host_free(p)
        if getattr(self, "h", None):
            self.L.lbmdem_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- life cycle (read_sample + main():1834-1861) ---------------------------------------
    def init(self, sample_path: str) -> int:
        self.n = self._ck(self.L.lbmdem_load_sample(self.h, os.fsencode(sample_path)))
        return self.n

    def init_arrays(self, r, x1, x2) -> int:
        r, x1, x2 = (np.ascontiguousarray(a, dtype=np.float64) for a in (r, x1, x2))
        self.n = self._ck(self.L.lbmdem_set_grains(self.h, len(r), r, x1, x2))
        return self.n

    def save_state(self, path: str):
        self._ck(self.L.lbmdem_save_state(self.h, os.fsencode(path)))

    def load_state(self, path: str) -> int:
        self.n = self._ck(self.L.lbmdem_load_state(self.h, os.fsencode(path)))
        return self.n

    def attach_local(self, group):
        """joins an in-process strip group (lbmdem_local_group_create): peer copies instead of NCCL"""
        self._ck(self.L.lbmdem_attach_local(self.h, group))

    def attach_nccl(self, uid: bytes):
        buf = C.create_string_buffer(uid, 128)
        self._ck(self.L.lbmdem_attach_nccl(self.h, buf))

    # -- stepping -------------------------------------------------------------------------
    def step(self, n=1):
        self._ck(self.L.lbmdem_step(self.h, n))

    def lbm_step(self):
        self._ck(self.L.lbmdem_lbm_step(self.h))

    def lbm_steps(self, n):
        self._ck(self.L.lbmdem_lbm_steps(self.h, n))

    def build_verlet(self):
        self._ck(self.L.lbmdem_build_verlet(self.h))

    def step_capture(self):
        """one renderScene() call; returns [n][6] x1 x2 x3 v1 v2 v3 as acceleration_grains() saw them"""
        mid = np.empty((self.n, 6))
        self._ck(self.L.lbmdem_step_capture(self.h, mid))
        return mid

    def _pinned(self, name, shape, f32=False):
        """a float64 / float32 array in page-locked memory (lbmdem_host_alloc), kept for the life of the solver"""
        bufs = self.__dict__.setdefault("_pinned_bufs", {})
        if name not in bufs:
            count = int(np.prod(shape))
            p = C.c_void_p()
            self._ck(self.L.lbmdem_host_alloc(8 * count, C.byref(p)))
            ctype = C.c_float if f32 else C.c_double
            arr = np.ctypeslib.as_array((ctype * count).from_address(p.value)).reshape(shape)
            bufs[name] = (p, arr)
        return bufs[name][1]

    def share(self):
        """[i0, i1): the grain rows this rank moves across the host boundary in step_host(..., share=True)"""
        a, b = C.c_int(), C.c_int()
        self._ck(self.L.lbmdem_get_share(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def step_host(self, state_in, n_dem_steps, want_state=True, want_fhf=True, want_density=True, rows="f64", share=False):
        """lbmdem_step_host (rows="f64") / lbmdem_step_host_f32 (rows="f32", single-precision solvers): host arrays in,
        host arrays out (the end-to-end call).  share=True: the *_share forms -- this rank's rows [i0, i1) of the grains
        only (see share()).  The buffers are page-locked and reused: the returned arrays are views that the next call
        overwrites; passing the returned state back in costs no host copy."""
        f32 = rows == "f32"
        key = "_hostbufs" + ("32" if f32 else "") + ("s" if share else "")
        hb = self.__dict__.get(key)
        if hb is None:
            # one input slot and two output blocks, so that the state returned by the previous call can be the input of
            # this one; an output block holds [n][9] state followed by [n][3] fhf: one device-to-host copy
            sfx = key[9:]
            i0, i1 = self.share() if share else (0, self.n)
            n = i1 - i0
            a_in = self._pinned("in" + sfx, (n, 9), f32)
            hb = [(a_in, a_in.ctypes.data_as(C.c_void_p))]
            for k in ("blk0", "blk1"):
                blk = self._pinned(k + sfx, (12 * n,), f32)
                st, fhv = blk[:9 * n].reshape(n, 9), blk[9 * n:].reshape(n, 3)
                hb.append((st, st.ctypes.data_as(C.c_void_p), fhv, fhv.ctypes.data_as(C.c_void_p)))
            self.__dict__[key] = hb
        (a_in, p_in), (a_o0, p_o0, a_f0, p_f0), (a_o1, p_o1, a_f1, p_f1) = hb
        sout = fh = None
        psin = psout = pfh = None
        use_o1 = False
        if state_in is not None:
            if state_in is a_o0:
                psin, use_o1 = p_o0, True
            elif state_in is a_o1:
                psin = p_o1
            else:
                np.copyto(a_in, state_in)
                psin = p_in
        if want_state:
            sout, psout = (a_o1, p_o1) if use_o1 else (a_o0, p_o0)
        if want_fhf:
            fh, pfh = (a_f1, p_f1) if use_o1 else (a_f0, p_f0)
        dens = C.c_double() if want_density else None
        if share:
            call = self.L.lbmdem_step_host_share_f32 if f32 else self.L.lbmdem_step_host_share
        else:
            call = self.L.lbmdem_step_host_f32 if f32 else self.L.lbmdem_step_host
        self._ck(call(self.h, psin, n_dem_steps, psout, pfh, C.cast(C.byref(dens), C.c_void_p) if dens is not None else None))
        return sout, fh, (dens.value if dens is not None else None)

    # -- scalars --------------------------------------------------------------------------
    def scalars(self) -> dict:
        d = (C.c_double * 11)()
        l = (C.c_long * 4)()
        self._ck(self.L.lbmdem_get_scalars(self.h, C.cast(d, C.c_void_p), C.cast(l, C.c_void_p)))
        keys = ["dx", "dtLB", "dt", "dt2", "c", "Mgx", "Mdx", "Mby", "Mhy", "xG", "yG"]
        out = dict(zip(keys, list(d)))
        out.update(npDEM=l[0], nbsteps=l[1], nFile=l[2], nbgrains=l[3])
        return out

    def set_nbsteps(self, n):
        self._ck(self.L.lbmdem_set_nbsteps(self.h, n))

    def total_density(self) -> float:
        s = C.c_double()
        self._ck(self.L.lbmdem_total_density(self.h, C.byref(s)))
        return s.value

    # -- arrays (owned rows [xlo, xhi)) -------------------------------------------------------
    def f(self):
        a = np.empty((self.nx, self.ly, 9))
        self._ck(self.L.lbmdem_get_f(self.h, a))
        return a

    def set_f(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == (self.nx, self.ly, 9)
        self._ck(self.L.lbmdem_set_f(self.h, a))

    def obst(self):
        a = np.empty((self.nx, self.ly), dtype=np.int32)
        self._ck(self.L.lbmdem_get_obst(self.h, a))
        return a

    def set_obst(self, a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        assert a.shape == (self.nx, self.ly)
        self._ck(self.L.lbmdem_set_obst(self.h, a))

    def act(self):
        a = np.empty((self.nx, self.ly), dtype=np.int32)
        self._ck(self.L.lbmdem_get_act(self.h, a))
        return a

    def grains(self):
        """[N,13]: x1 x2 x3 v1 v2 v3 a1 a2 a3 r m It rLB"""
        a = np.empty((self.n, 13))
        self._ck(self.L.lbmdem_get_grains(self.h, a))
        return a

    def set_grain_state(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == (self.n, 9)
        self._ck(self.L.lbmdem_set_grain_state(self.h, a))

    def fhf(self):
        a = np.empty((self.n, 3))
        self._ck(self.L.lbmdem_get_fhf(self.h, a))
        return a

    def set_fhf(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == (self.n, 3)
        self._ck(self.L.lbmdem_set_fhf(self.h, a))

    def verlet_full(self):
        cap = self.params.neighbour_capacity
        cnt = np.zeros(self.n, dtype=np.int32)
        nbr = np.zeros((self.n, cap), dtype=np.int32)
        wf = np.zeros(self.n, dtype=np.int32)
        self._ck(self.L.lbmdem_get_verlet(self.h, cnt, nbr, cap, wf))
        return cnt, nbr, wf

    def verlet(self):
        """(cumul, neighbours) of the reference's HALF list (src/main.c:1519-1543), derived
        from the device's full lists; cumul[N-1] = 0 like the reference leaves it."""
        cnt, nbr, _ = self.verlet_full()
        half, cumul = [], np.zeros(self.n, dtype=np.int32)
        for i in range(self.n):
            row = nbr[i, :cnt[i]]
            half.extend(int(j) for j in row if j > i)
            cumul[i] = len(half)
        if self.n:
            cumul[self.n - 1] = 0
        return cumul, np.array(half, dtype=np.int32)

    def wall_lists(self):
        _, _, wf = self.verlet_full()
        return [np.nonzero(wf & (1 << k))[0].astype(np.int32) for k in range(4)]  # B T L R

    def fields(self, grain_p=None):
        """write_vtk's five point fields in VTK order [y][x] (x fastest)."""
        nn = self.nx * self.ly
        gp = None if grain_p is None else np.ascontiguousarray(grain_p, dtype=np.float64)
        out = [np.empty(nn, dtype=np.float32), np.empty(nn * 3, dtype=np.float32), np.empty(nn * 3, dtype=np.float32),
               np.empty(nn, dtype=np.float32), np.empty(nn * 3, dtype=np.float32)]
        self._ck(self.L.lbmdem_get_fields(self.h, None if gp is None else gp.ctypes.data_as(C.c_void_p), *out))
        names = ["grain_pressure", "grain_velocity", "grain_acceleration", "fluid_pressure", "fluid_velocity"]
        shp = [(self.ly, self.nx), (self.ly, self.nx, 3), (self.ly, self.nx, 3), (self.ly, self.nx), (self.ly, self.nx, 3)]
        return {k: a.reshape(s) for k, a, s in zip(names, out, shp)}

    # -- instrumentation --------------------------------------------------------------------
    def reset_kernel_timer(self, enable=True):
        self._ck(self.L.lbmdem_reset_kernel_timer(self.h, int(enable)))

    def kernel_timer(self):
        ms, k1, al = C.c_double(), C.c_long(), C.c_long()
        self._ck(self.L.lbmdem_get_kernel_timer(self.h, C.byref(ms), C.byref(k1), C.byref(al)))
        return ms.value, k1.value, al.value

    def state_checksum(self):
        """(f fingerprint, obst fingerprint) of the owned rows; strips add up mod 2^64 to the one-GPU value"""
        c = (C.c_ulonglong * 2)()
        self._ck(self.L.lbmdem_state_checksum(self.h, c))
        return int(c[0]), int(c[1])

    def list_counts(self) -> dict:
        """sizes of the sparse work lists of the last LBM step (bounce-back links, boundary nodes, deferred links)"""
        c = (C.c_long * 4)()
        self._ck(self.L.lbmdem_get_list_counts(self.h, c))
        return {"links": c[0], "boundary_nodes": c[1], "deferred": c[2], "tiles_rebuilt": c[3]}

    def stream(self) -> int:
        return int(self.L.lbmdem_stream(self.h) or 0)
