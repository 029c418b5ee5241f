#!/usr/bin/env python
"""tools/make_profiles.py <tag> [round] -- turns one tools/gpu_check.sh visit (gpurun_out/<tag>_*) into the
tracked summaries under profiles/: launch list shares, key ncu metrics of K1 (and of the sparse
kernels when captured), the bench line, and profiles/k1_traffic.json (DRAM bytes per K1 launch,
read back by bench.py as roofline.traffic)."""
import contextlib
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary  # noqa: E402


def capture(fn, *a):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        fn(*a)
    return buf.getvalue()


def main():
    tag = sys.argv[1]
    rnd = sys.argv[2] if len(sys.argv) > 2 else "r01"
    work = sys.argv[3] if len(sys.argv) > 3 else "cfg4"
    src = os.path.join(ROOT, "gpurun_out")
    dst = os.path.join(ROOT, "profiles")
    os.makedirs(dst, exist_ok=True)
    g = lambda name: os.path.join(src, f"{tag}_{name}")
    if os.path.exists(g("launches.csv")):
        open(os.path.join(dst, f"{rnd}_{work}_launches.txt"), "w").write(capture(ncu_summary.launches, g("launches.csv")))
    for part in ("k1", "aux"):
        rep = g(f"{part}.ncu-rep")
        if os.path.exists(rep):
            open(os.path.join(dst, f"{rnd}_{work}_{part}_ncu.txt"), "w").write(
                f"# ncu --set full --clock-control none of `python bench.py --workload {work} --steps 3 --warmup 3` ({tag})\n"
                + capture(ncu_summary.report, rep))
    rep = g("k1.ncu-rep")
    if os.path.exists(rep):
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        tot = []
        for r in rows[2:]:
            if "lbm_rows" not in r[hdr.index("Kernel Name")]:
                continue
            b = 0.0
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                v, u = float(r[hdr.index(k)].replace(",", "")), units[hdr.index(k)]
                b += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            tot.append(b)
        if tot:
            path = os.path.join(dst, "k1_traffic.json")
            d = json.load(open(path)) if os.path.exists(path) else {}
            d[f"{work}_bytes_per_launch"] = sum(tot) / len(tot)
            d[f"{work}_source"] = f"{rnd}_{work}_k1_ncu.txt: dram__bytes_read.sum + dram__bytes_write.sum, mean of {len(tot)} launches"
            json.dump(d, open(path, "w"), indent=1)
    if os.path.exists(g("strips_launches.csv")):
        open(os.path.join(dst, f"{rnd}_strips_launches.txt"), "w").write(
            "# python tools/strip_launches.py 2 6: two strips of the benchmarked sample in ONE process (in-process strip group)\n"
            + capture(ncu_summary.launches, g("strips_launches.csv")))
    for name in ("bench.json", "smi.txt", "host.txt", "pytest.log", "smoke.log"):
        if os.path.exists(g(name)):
            shutil.copyfile(g(name), os.path.join(dst, f"{rnd}_{work}_{name}"))
    for name, out in (("cfg2.json", "cfg2_bench.json"), ("cfg3.json", "cfg3_bench.json"), ("cfg5.json", "cfg5_bench.json"),
                      ("strict.json", "cfg4_bench_strict_build.json"), ("ref.json", "cfg4_bench_reference_arm.json"),
                      ("memcheck.log", "sanitizer_memcheck_smoke.log"), ("racecheck.log", "sanitizer_racecheck_smoke.log")):
        if os.path.exists(g(name)):
            shutil.copyfile(g(name), os.path.join(dst, f"{rnd}_{out}"))


if __name__ == "__main__":
    main()
