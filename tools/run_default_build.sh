#!/bin/bash
# Application-level timing of the drop-in executable at the reference's default build size
# (7826 x 2325 fp64, 47 980 synthetic grains): 1200 renderScene() calls = 100 LBM steps, 12 Verlet builds.
cd "$(dirname "$0")/.."
mkdir -p /tmp/lbmdem_out gpurun_out
python - <<'PY'
import subprocess, sys, time
sys.path.insert(0, "tools")
import make_sample as ms
n, a, b, w = ms.PRESETS["50000-test"]
r, x, y = ms.packed_sample(n, a, b, w, seed=12345)
ms.write_sample("/tmp/s50k.data", r, x, y, comment="# synthetic 50000-test seed=12345")
for steps in (1200, 12000):
    t0 = time.time()
    p = subprocess.run(["./2d-lbm-dem_b200/lbmdem", "/tmp/s50k.data", "--steps", str(steps), "--outdir", "/tmp/lbmdem_out"],
                       capture_output=True, text=True)
    dt = time.time() - t0
    line = (f"lbmdem default build (7826x2325 fp64, {n} grains): {steps} renderScene() calls, {steps // 12} LBM steps: "
            f"{dt:.2f} s wall from process start to exit (rc {p.returncode}); {p.stderr.strip().splitlines()[-1] if p.stderr.strip() else ''}")
    print(line)
    open("gpurun_out/default_build_run.log", "a").write(line + "\n")
PY
