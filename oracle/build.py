"""oracle/build.py -- builds the CHECKERS (test infrastructure, not product).

Two things are built here, both into git-ignored locations:

1. ``oracle/_ref/libref_<lx>x<ly>_s<scale>_<f64|f32>[_omp].so`` -- the UNMODIFIED
   reference (``/root/reference/src/main.c`` + ``visit_writer.c``) compiled in
   place through ``oracle/ref_shim.c``; one library per compile-time
   configuration, because the reference fixes lx/ly/scale/precision with -D
   macros (src/main.c:24-40).  Only possible where ``/root/reference`` exists
   (the authoring container); on the GPU box the prebuilt files that travelled
   with the snapshot are used.
2. ``oracle/_build/liboracle.so`` -- the plain-C restatement ``oracle/lbmdem_oracle.c``
   (runtime lx/ly/scale, both precisions), pinned against (1) by
   ``tests/test_oracle_pin.py``.

Flags.  The *correctness* oracle is ``-std=c99 -O2 -ffp-contract=off``, serial:
the reference's own release flags (``-Ofast -march=native ...``, CMakeLists.txt:17)
perturb results at 1e-11 per step.  The *timing* baseline
(``release=True``) uses the reference's GNU release flag set, except that
``-march=native`` becomes ``-march=x86-64-v3`` because the library is built in
this container and executed on the GPU box's host CPU.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("LBMDEM_REFERENCE", "/root/reference")
REF_MAIN = os.path.join(REF_ROOT, "src", "main.c")
REF_VISIT = os.path.join(REF_ROOT, "src", "visit_writer.c")
REF_DIR = os.path.join(HERE, "_ref")
BUILD_DIR = os.path.join(HERE, "_build")

ORACLE_FLAGS = ["-std=c99", "-O2", "-ffp-contract=off"]
# CMakeLists.txt:12,17 (GNU), -march=native replaced (see module docstring)
RELEASE_FLAGS = ["-std=c99", "-Ofast", "-march=x86-64-v3", "-flto", "-fipa-pta",
                 "-funsafe-math-optimizations", "-fno-math-errno", "-fno-trapping-math"]

# Configurations that must exist on the GPU box (built by __graft_entry__.build()).
# (lx, ly, scale, precision, omp, release)
PREBUILT = [
    (64, 48, "1.", "f64", False, False),      # tiny unit cases
    (64, 48, "1.", "f32", False, False),
    (256, 256, "1.", "f64", False, False),    # synthetic packing, parity
    (256, 256, "1.", "f32", False, False),
    (512, 512, "1.", "f64", False, False),    # SURVEY 4.4 known-answer configuration (a08d83)
    (4096, 4096, "2.7", "f32", True, True),   # BASELINE cfg 4: timed CPU baseline (all cores)
    (4096, 4096, "2.7", "f64", True, True),   # same lattice in fp64
    (4096, 4096, "2.7", "f32", False, True),  # serial variants of the timed baseline
    (4096, 4096, "2.7", "f64", False, True),
    (2048, 2048, "1.", "f64", True, True),    # BASELINE cfg 3 / cfg 2 timed baselines (bench.py --workload)
    (1024, 1024, "1.", "f64", True, True),
    (8192, 8192, "2.6", "f64", True, True),   # BASELINE cfg 5 as stated (bench.py --workload cfg5 --impl reference)
]


def ref_available() -> bool:
    return os.path.isfile(REF_MAIN) and os.path.isfile(REF_VISIT)


def ref_lib_path(lx, ly, scale="1.", prec="f64", omp=False, release=False) -> str:
    tag = f"{lx}x{ly}_s{str(scale).rstrip('.')}_{prec}"
    if release:
        tag += "_rel"
    if omp:
        tag += "_omp"
    return os.path.join(REF_DIR, f"libref_{tag}.so")


def _newer(target: str, *deps: str) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps if os.path.exists(d))


def build_ref(lx, ly, scale="1.", prec="f64", omp=False, release=False, verbose=False):
    """Return the path of the reference library for this configuration, building it
    if the reference sources are present; None if it is neither built nor buildable."""
    out = ref_lib_path(lx, ly, scale, prec, omp, release)
    shim = os.path.join(HERE, "ref_shim.c")
    if _newer(out, shim):
        return out
    if not ref_available():
        return out if os.path.exists(out) else None
    os.makedirs(REF_DIR, exist_ok=True)
    flags = list(RELEASE_FLAGS if release else ORACLE_FLAGS)
    if omp:
        flags.append("-fopenmp")
    if prec == "f32":
        flags.append("-DSINGLE_PRECISION")
    cmd = ["gcc", *flags, "-w", "-shared", "-fPIC", "-fvisibility=hidden", "-Wl,-Bsymbolic",
           f"-Dlx={lx}", f"-Dly={ly}", f"-Dscale={scale}",
           f'-DLBMDEM_REF_MAIN="{REF_MAIN}"', shim, REF_VISIT, "-lm", "-o", out]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return out


def oracle_lib_path() -> str:
    return os.path.join(BUILD_DIR, "liboracle.so")


def build_oracle(verbose=False) -> str:
    """Compile the plain-C restatement (both precisions in one library)."""
    out = oracle_lib_path()
    src = os.path.join(HERE, "lbmdem_oracle.c")
    hdr = os.path.join(HERE, "lbmdem_oracle.h")
    if _newer(out, src, hdr):
        return out
    os.makedirs(BUILD_DIR, exist_ok=True)
    objs = []
    for prec, define in (("f64", []), ("f32", ["-DORACLE_SINGLE"])):
        obj = os.path.join(BUILD_DIR, f"oracle_{prec}.o")
        cmd = ["gcc", *ORACLE_FLAGS, "-Wall", "-fPIC", "-fvisibility=hidden", *define,
               "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
        objs.append(obj)
    subprocess.run(["gcc", "-shared", "-Wl,-Bsymbolic", *objs, "-lm", "-o", out], check=True)
    return out


def build_all(verbose=False):
    build_oracle(verbose)
    built = []
    for cfg in PREBUILT:
        p = build_ref(*cfg, verbose=verbose)
        built.append((cfg, p))
    return built


if __name__ == "__main__":
    for cfg, p in build_all(verbose="-v" in sys.argv):
        print(cfg, "->", p)
