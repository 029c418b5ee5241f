"""The INTEGRATION.md binding, compiled INTO the reference.

tools/integration/build.py patches a scratch copy of the reference's src/main.c at three anchors (tools/integration/
lbmdem_glue.h is the code a maintainer would add) and links it against liblbmdem_gpu.so: the reference's own main(),
read_sample, console lines and output writers, with the body of renderScene() -- the coupled LBM + DEM step -- behind
the C ABI.  CPU: it builds, links, and refuses to run without a GPU.  GPU (strict build): 8000 renderScene() calls
write the very files the unmodified reference wrote (tests/golden/outputs_64x48)."""
import importlib.util
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _builder():
    spec = importlib.util.spec_from_file_location("integration_build", os.path.join(ROOT, "tools", "integration", "build.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _exe():
    b = _builder()
    dt = float(np.load(os.path.join(GOLD, "outputs_64x48", "state_8000.npz"))["scalar_dt"])
    exe = b.build(64, 48, 7999.5 * dt)      # main()'s do-while leaves after exactly 8000 calls
    if exe is None or not os.path.exists(exe):
        pytest.skip("the reference sources are not here and no prebuilt integrated executable travelled along")
    return exe


def test_binding_compiles_into_the_reference_and_links_the_c_abi():
    exe = _exe()
    out = subprocess.run(["ldd", exe], capture_output=True, text=True, check=True).stdout
    assert "liblbmdem_gpu.so" in out and "not found" not in out
    syms = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True, check=True).stdout
    used = set(re.findall(r"\bU (lbmdem_[a-z0-9_]+)", syms))
    assert {"lbmdem_create", "lbmdem_load_sample", "lbmdem_step", "lbmdem_get_f", "lbmdem_get_obst", "lbmdem_get_grains",
            "lbmdem_get_fhf", "lbmdem_total_density", "lbmdem_get_scalars", "lbmdem_default_params"} <= used


def test_integrated_reference_has_no_cpu_path(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    p = subprocess.run([_exe(), os.path.join(GOLD, "pack_64x48_f64.data")], cwd=tmp_path, capture_output=True, text=True)
    assert p.returncode != 0 and "no CUDA device visible" in p.stderr
    assert "Nb grains 7" in p.stdout      # the reference's own read_sample ran before the device was asked for


@pytest.mark.gpu
def test_integrated_reference_writes_the_reference_files(tmp_path):
    gold = os.path.join(GOLD, "outputs_64x48")
    env = dict(os.environ, LBMDEM_STRICT="1")
    p = subprocess.run([_exe(), os.path.join(GOLD, "pack_64x48_f64.data")], cwd=tmp_path, capture_output=True, text=True,
                       env=env, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    # the VTK fields the device state feeds (grain_pressure shows g[i].p, a contact diagnostic the reference
    # accumulates inside force_grains: not part of the hot path's state)
    for name in ("fluid_pressure_000000.vtk", "fluid_velocity_000000.vtk", "grain_velocity_000000.vtk",
                 "grain_acceleration_000000.vtk"):
        assert open(tmp_path / name, "rb").read() == open(os.path.join(gold, name), "rb").read(), name
    # the grain table: kinematics and hydrodynamic forces (columns i r x1 x2 x3 v1 v2 v3 a1 a2 a3 fhf1 fhf2 fhf3)
    for name in ("DEM000000.dat", "DEM000001.dat"):
        mine = [l.split("\t")[:14] for l in open(tmp_path / name).read().splitlines()]
        ref = [l.split("\t")[:14] for l in open(os.path.join(gold, name)).read().splitlines()]
        assert mine == ref, name
    # final_density(): the reference's own serial sum over the populations that came back through the C ABI
    st = np.load(os.path.join(gold, "state_8000.npz"))
    assert f"final_density: {float(st['density']):f}" in p.stderr
    assert "Iteration Number 8000" in p.stdout
