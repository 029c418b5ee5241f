OUT=gpurun_out; RUN=r02Y; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/${RUN}_pytest.log 2>&1; tail -5 $OUT/${RUN}_pytest.log
timeout 200 python bench.py --steps 300 --no-cpu-baseline --no-cfg5 > $OUT/${RUN}_default.json 2> $OUT/${RUN}_default.err
for t in sw5 sw6; do
  LBMDEM_LIB=$PWD/2d-lbm-dem_b200/liblbmdem_gpu_$t.so timeout 200 python bench.py --steps 300 --no-cpu-baseline --no-cfg5 > $OUT/${RUN}_$t.json 2> $OUT/${RUN}_$t.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${RUN}_launches.csv python bench.py --steps 3 --warmup 8 --no-cpu-baseline --no-cfg5 > $OUT/${RUN}_launches.log 2>&1
python - <<PY
import json,glob
for p in sorted(glob.glob("$OUT/${RUN}_*.json")):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        print(p.split("${RUN}_")[1][:-5].ljust(14), "MLUPS %.0f  ms/step %.4f  K1 ms %.4f frac %.3f  e2e %.0f" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"]))
    except Exception as e:
        print(p, "unreadable", e)
PY
python tools/ncu_summary.py launches $OUT/${RUN}_launches.csv | cut -c1-120 | sed -n 3,10p
