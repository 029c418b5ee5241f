"""Builds the plain-C host side: liblbmdem_host.so (VTK / DEM writers, no CUDA, testable on a CPU box)
and the `lbmdem` executable (lbmdem_main.c + the writers, linked against ../liblbmdem_gpu.so).

    python 2d-lbm-dem_b200/host/build.py
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
HOST_LIB = os.path.join(PKG, "liblbmdem_host.so")
EXE = os.path.join(PKG, "lbmdem")
CFLAGS = ["-std=c99", "-O2", "-ffp-contract=off", "-Wall", "-Wextra", "-D_DEFAULT_SOURCE", "-D_POSIX_C_SOURCE=200809L"]
WRITERS = ["vtk_writer.c", "dem_output.c"]


def _stale(target, deps):
    return not os.path.exists(target) or any(os.path.getmtime(d) > os.path.getmtime(target) for d in deps)


def build_host_lib() -> str:
    srcs = [os.path.join(HERE, f) for f in WRITERS]
    deps = srcs + [os.path.join(HERE, h) for h in ("vtk_writer.h", "dem_output.h", "dem_output_impl.h")]
    if _stale(HOST_LIB, deps):
        subprocess.run(["gcc", *CFLAGS, "-fPIC", "-shared", *srcs, "-lm", "-o", HOST_LIB], check=True)
    return HOST_LIB


def build_exe() -> str:
    gpu_lib = os.path.join(PKG, "liblbmdem_gpu.so")
    srcs = [os.path.join(HERE, f) for f in ["lbmdem_main.c", *WRITERS]]
    deps = srcs + [gpu_lib, os.path.join(PKG, "..", "include", "lbmdem_gpu.h")]
    if _stale(EXE, deps):
        subprocess.run(["gcc", *CFLAGS, *srcs, "-L", PKG, "-llbmdem_gpu", "-Wl,-rpath,$ORIGIN", "-lm", "-o", EXE], check=True)
    return EXE


def build() -> tuple[str, str]:
    return build_host_lib(), build_exe()


if __name__ == "__main__":
    print(*build(), sep="\n")
