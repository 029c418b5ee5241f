"""Builds liblbmdem_gpu.so (sm_100a only) in-tree with nvcc.

    python 2d-lbm-dem_b200/build.py [-v] [--force]

Translation units:
  csrc/lbm_kernels.cu   twice: -DK1_NS=k1_fast -DLBM_RELAXED (contraction on, shared reciprocals) and
                        -DK1_NS=k1_strict (-fmad=false, the reference's expressions verbatim)
  csrc/aux_kernels.cu   -fmad=false (bit-exact obstacle map and DEM step)
  csrc/sim.cu           host orchestration + C ABI (include/lbmdem_gpu.h)
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "liblbmdem_gpu.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", *ARCH, "-Xcompiler", "-fPIC,-fvisibility=hidden", "-I", CSRC]
HEADERS = ["kernels.h", "lbm_node.cuh", "raster_node.cuh", "dem_node.cuh"]

UNITS = [
    ("lbm_kernels_fast.o", "lbm_kernels.cu", ["-DK1_NS=k1_fast", "-DLBM_RELAXED"]),
    ("lbm_kernels_strict.o", "lbm_kernels.cu", ["-DK1_NS=k1_strict", "-fmad=false"]),
    ("aux_kernels.o", "aux_kernels.cu", ["-fmad=false"]),
    ("sim.o", "sim.cu", ["-fmad=false"]),
]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose=False, force=False, tag="", defines=()) -> str:
    """tag / defines: experimental variants (tools/k1_variants.sh): objects and library get the tag as a suffix"""
    global OBJ, LIB
    if tag:
        OBJ = os.path.join(HERE, "_obj_" + tag)
        LIB = os.path.join(HERE, f"liblbmdem_gpu_{tag}.so")
        force = force or not os.path.exists(LIB)
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.join(HERE, "..", "include", "lbmdem_gpu.h"),
                                                         os.path.abspath(__file__)]
    jobs = []
    for obj, src, flags in UNITS:
        o, s = os.path.join(OBJ, obj), os.path.join(CSRC, src)
        if force or _stale(o, [s] + hdrs):
            jobs.append(["nvcc", *COMMON, *flags, *defines, "-c", s, "-o", o])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)

    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(OBJ, u[0]) for u in UNITS]
    if jobs or force or _stale(LIB, objs):
        run(["nvcc", *ARCH, "-shared", "-o", LIB, *objs, "-ldl"])
    return LIB


if __name__ == "__main__":
    tag = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--tag=")), "")
    print(build(verbose="-v" in sys.argv, force="--force" in sys.argv, tag=tag, defines=[a for a in sys.argv if a.startswith("-D")]))
