"""Host-side logic of the ctypes mirror's end-to-end call (no GPU): Solver.step_host keeps page-locked buffers, hands
the state it returned back in without a host copy, and lays state + fhf out as ONE block so that the library downloads
them with one copy (csrc/sim.cu: step_host).  The library is replaced by a stand-in that records the pointers."""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mirror():
    spec = importlib.util.spec_from_file_location("lbmdem_gpu_mirror", os.path.join(ROOT, "2d-lbm-dem_b200", "lbmdem_gpu.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


class _FakeLib:
    """malloc instead of cudaHostAlloc; the step calls write recognisable values through the pointers they get"""

    def __init__(self, n):
        self.n, self.calls, self.freed = n, [], []
        self.libc = C.CDLL(None)
        self.libc.malloc.restype = C.c_void_p
        self.libc.malloc.argtypes = [C.c_size_t]

    def lbmdem_host_alloc(self, nbytes, pp):
        C.cast(pp, C.POINTER(C.c_void_p))[0] = self.libc.malloc(nbytes)
        return 0

    def lbmdem_host_free(self, p):
        self.freed.append(p.value)
        return 0

    def lbmdem_destroy(self, h):
        return None

    def _step(self, ctype, pin, pout, pfh, pd):
        n = self.n
        v = None if pin is None or pin.value is None else np.ctypeslib.as_array((ctype * (9 * n)).from_address(pin.value)).copy()
        self.calls.append((ctype, pin.value if pin else None, pout.value if pout else None, pfh.value if pfh else None, v))
        if pout is not None and pout.value:
            out = np.ctypeslib.as_array((ctype * (9 * n)).from_address(pout.value))
            out[:] = (v if v is not None else 0) + 1          # "one step": every value + 1
        if pfh is not None and pfh.value:
            np.ctypeslib.as_array((ctype * (3 * n)).from_address(pfh.value))[:] = len(self.calls)
        if pd is not None and pd.value:
            C.cast(pd, C.POINTER(C.c_double))[0] = 42.0
        return 0

    def lbmdem_step_host(self, h, pin, nsteps, pout, pfh, pd):
        return self._step(C.c_double, pin, pout, pfh, pd)

    def lbmdem_step_host_f32(self, h, pin, nsteps, pout, pfh, pd):
        return self._step(C.c_float, pin, pout, pfh, pd)


@pytest.mark.parametrize("rows", ["f64", "f32"])
def test_step_host_buffers(rows):
    G = _mirror()
    n = 7
    s = G.Solver.__new__(G.Solver)
    s.L, s.h, s.n = _FakeLib(n), 1, n
    s._ck = lambda rc: rc
    eb = 8 if rows == "f64" else 4
    st0 = np.arange(9.0 * n).reshape(n, 9)
    a, fa, d = s.step_host(st0, 2, rows=rows)
    assert d == 42.0 and a.dtype == (np.float64 if rows == "f64" else np.float32)
    assert np.array_equal(a, st0 + 1) and np.all(fa == 1)
    b, fb, _ = s.step_host(a, 2, rows=rows, want_density=False)
    c, fc, _ = s.step_host(b, 2, rows=rows)
    assert np.array_equal(c, st0 + 3) and np.all(fc == 3)
    calls = s.L.calls
    for k, (ctype, pin, pout, pfh, _) in enumerate(calls):
        assert pfh == pout + eb * 9 * n                      # state and fhf are one block: one download
        assert pout != pin                                   # never in place
        if k:
            assert pin == calls[k - 1][2]                    # the returned state goes back in without a copy
    assert c is a                                            # two output blocks, alternating
    # a foreign array is copied into the input slot; outputs can be declined
    e, fe, de = s.step_host(np.zeros((n, 9)), 1, want_state=False, want_fhf=False, want_density=False, rows=rows)
    assert e is None and fe is None and de is None and calls[-1][2] is None and calls[-1][3] is None
    assert calls[-1][1] not in (calls[0][2], calls[1][2])
    s.close()
    assert len(s.L.freed) == 3                               # input slot + two blocks
