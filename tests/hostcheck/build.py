"""tests/hostcheck/build.py -- builds the host-compiled check of the product's node headers.
TEST INFRASTRUCTURE ONLY (see hostcheck.cpp)."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "2d-lbm-dem_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libhostcheck.so")


def build() -> str:
    src = os.path.join(HERE, "hostcheck.cpp")
    deps = [src] + [os.path.join(CSRC, h) for h in ("lbm_node.cuh", "raster_node.cuh", "dem_node.cuh")]
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-fvisibility=hidden",
           "-x", "c++", "-I", CSRC, src, "-o", OUT]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build())
