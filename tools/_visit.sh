OUT=gpurun_out; RUN=r02B; mkdir -p $OUT
for t in ns5 ns6 stcs; do
  LBMDEM_LIB=$PWD/2d-lbm-dem_b200/liblbmdem_gpu_$t.so timeout 200 python bench.py --workload cfg4 --steps 300 --no-cpu-baseline > $OUT/${RUN}_$t.json 2> $OUT/${RUN}_$t.err
done
timeout 200 python bench.py --workload cfg4 --steps 100 --no-cpu-baseline --kernel 1 > $OUT/${RUN}_kernel1.json 2> $OUT/${RUN}_kernel1.err
python - <<PY
import json,glob
for p in sorted(glob.glob("$OUT/${RUN}_*.json")):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        print(p.split("${RUN}_")[1][:-5].ljust(10), "MLUPS %.0f  ms/step %.4f  K1 ms %.4f frac %.3f  e2e %.0f" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"]))
    except Exception as e:
        print(p, "unreadable", e)
PY
