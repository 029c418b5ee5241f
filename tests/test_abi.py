"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/lbmdem_gpu.h declares, the FFI struct matches, and -- there being no CPU path -- the
library refuses loudly to create a context when no sm_100 GPU is visible."""
import ctypes as C
import os
import re
import subprocess

import pytest

import lbmdem_gpu as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_in_header():
    src = open(os.path.join(ROOT, "include", "lbmdem_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lbmdem_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = G.load_library()
    names = _declared_in_header()
    assert len(names) >= 30
    out = subprocess.run(["nm", "-D", "--defined-only", G.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (lbmdem_[a-z0-9_]+)", out))
    missing = [n for n in names if n not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"
    for n in names:
        assert getattr(L, n) is not None
    # the python mirror binds the same set
    assert sorted(L._declared) == names


def test_default_params_are_the_reference_defaults():
    p = G.default_params()
    assert (p.lx, p.ly, p.scale, p.single_precision) == (7826, 2325, 1.0, 0)          # src/main.c:24-40
    assert (p.tau, p.nu, p.rho_moy, p.reductionR) == (0.504, 1e-6, 1000.0, 0.85)        # :74-94
    assert (p.s2, p.s3, p.s5, p.s7, p.s8, p.s9) == (1.5, 1.4, 1.5, 1.5, 1.9841, 1.9841)  # :79-80
    assert (p.kg, p.kt, p.km, p.ktm) == (1.6e6, 1.0e6, 3e6, 2e6)                          # :100-103
    assert (p.mu, p.mum, p.mumb, p.murf) == (.5317, .466, .466, 0.01)
    assert (p.UpdateVerlet, p.stepFilm, p.distVerlet, p.iterDEM) == (100, 8000, 5e-4, 100.0)
    assert G.load_library().lbmdem_sizeof_params() == C.sizeof(G.Params)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    with pytest.raises(G.LbmdemError) as ei:
        G.Solver(64, 48)
    assert ei.value.code == -2 and "no CUDA device" in str(ei.value)


def test_bad_arguments_are_rejected_before_any_device_work():
    L = G.load_library()
    assert L.lbmdem_default_params(None) == -1
    assert L.lbmdem_create(None, None) == -1
    assert L.lbmdem_step(None, 1) == -1
    assert L.lbmdem_get_f(None, __import__("numpy").zeros(1)) == -1
