OUT=gpurun_out; RUN=r02W; mkdir -p $OUT
for t in mb7; do
  LBMDEM_LIB=$PWD/2d-lbm-dem_b200/liblbmdem_gpu_$t.so timeout 200 python bench.py --steps 300 --no-cpu-baseline --no-cfg5 > $OUT/${RUN}_$t.json 2> $OUT/${RUN}_$t.err
done
timeout 200 python bench.py --steps 300 --no-cpu-baseline --no-cfg5 > $OUT/${RUN}_default.json 2> $OUT/${RUN}_default.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lbm_rows -s 6 -c 1 -f -o $OUT/${RUN}_k1 python bench.py --steps 3 --warmup 8 --no-cpu-baseline --no-cfg5 > $OUT/${RUN}_k1.log 2>&1
python - <<PY
import json,glob
for p in sorted(glob.glob("$OUT/${RUN}_*.json")):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        print(p.split("${RUN}_")[1][:-5].ljust(14), "MLUPS %.0f  ms/step %.4f  K1 ms %.4f frac %.3f  e2e %.0f" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"]))
    except Exception as e:
        print(p, "unreadable", e)
PY
