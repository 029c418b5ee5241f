#!/bin/bash
# Application-level timing of the drop-in executable at the reference's default build size (7826 x 2325 fp64) on the
# reference's own bin/50000-test.data (47 980 grains; BASELINE configs[0]): 1200 renderScene() calls = 100 LBM steps,
# 12 Verlet builds; then 12000 calls.  Grains above y = 232 mm lie outside the lattice: legal, as in the reference.
cd "$(dirname "$0")/.."
mkdir -p /tmp/lbmdem_out gpurun_out
python - <<'PY'
import subprocess, time
sample = "tests/golden/50000-test.data"
for steps in (1200, 12000):
    t0 = time.time()
    p = subprocess.run(["./2d-lbm-dem_b200/lbmdem", sample, "--steps", str(steps), "--outdir", "/tmp/lbmdem_out"],
                       capture_output=True, text=True)
    dt = time.time() - t0
    line = (f"lbmdem default build (7826x2325 fp64, bin/50000-test.data, 47980 grains): {steps} renderScene() calls, {steps // 12} LBM steps: "
            f"{dt:.2f} s wall from process start to exit (rc {p.returncode}); {p.stderr.strip().splitlines()[-1] if p.stderr.strip() else ''}")
    print(line)
    open("gpurun_out/default_build_run.log", "a").write(line + "\n")
PY
