"""tools/integration/build.py -- compiles the INTEGRATION.md binding INTO the reference and links it against
liblbmdem_gpu.so: the proof that the C ABI is the drop-in boundary it claims to be.

A scratch copy of /root/reference/src/main.c (never committed) is patched at three anchors -- see lbmdem_glue.h -- and
built with the reference's own visit_writer.c into  oracle/_ref/lbmdem_integrated_<lx>x<ly>  (git-ignored, travels to
the GPU box like the other built checkers).  Only possible where /root/reference exists; elsewhere the prebuilt
executable is used.

    python tools/integration/build.py [lx ly [duration]]
"""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_ROOT = os.environ.get("LBMDEM_REFERENCE", "/root/reference")
OUT_DIR = os.path.join(ROOT, "oracle", "_ref")
LIB_DIR = os.path.join(ROOT, "2d-lbm-dem_b200")

ANCHORS = [
    # (text that must occur exactly once, replacement)
    ("#define duration 1.5", "#ifndef duration\n#define duration 1.5\n#endif //"),
    ("void renderScene(void) {\n  long i;\n", '#include "lbmdem_glue.h"\nvoid renderScene(void) {\n  long i;\n'
                                             "#ifdef LBMDEM_GPU\n  if (gpu_render()) return;\n#endif\n"),
    ("  init_obst();\n\n  //\tVerletWall();", "  init_obst();\n#ifdef LBMDEM_GPU\n  gpu_setup(argv[1]);\n#endif\n\n  //\tVerletWall();"),
]


def exe_path(lx, ly):
    return os.path.join(OUT_DIR, f"lbmdem_integrated_{lx}x{ly}")


def patched_source() -> str:
    src = open(os.path.join(REF_ROOT, "src", "main.c")).read()
    # the renderScene prototype near the top of the file also matches the second anchor's first line: anchor on the body
    for old, new in ANCHORS:
        if src.count(old) != 1:
            raise RuntimeError(f"anchor {old!r} occurs {src.count(old)} times in the reference's main.c")
        src = src.replace(old, new)
    return src


def build(lx=64, ly=48, duration=None, verbose=False):
    out = exe_path(lx, ly)
    if not os.path.isfile(os.path.join(REF_ROOT, "src", "main.c")):
        return out if os.path.exists(out) else None
    os.makedirs(OUT_DIR, exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="lbmdem_integ_") as tmp:
        path = os.path.join(tmp, "main_patched.c")
        with open(path, "w") as fh:
            fh.write(patched_source())
        cmd = ["gcc", "-std=c99", "-O2", "-ffp-contract=off", "-w", "-DLBMDEM_GPU", f"-Dlx={lx}", f"-Dly={ly}", "-Dscale=1.",
               "-I", os.path.join(ROOT, "include"), "-I", HERE, "-I", os.path.join(REF_ROOT, "src"),
               path, os.path.join(REF_ROOT, "src", "visit_writer.c"),
               "-L", LIB_DIR, "-llbmdem_gpu", "-Wl,-rpath,$ORIGIN/../../2d-lbm-dem_b200", "-lm", "-o", out]
        if duration is not None:
            cmd.insert(6, f"-Dduration={duration!r}")
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    a = sys.argv[1:]
    print(build(int(a[0]) if a else 64, int(a[1]) if len(a) > 1 else 48, float(a[2]) if len(a) > 2 else None, verbose=True))
