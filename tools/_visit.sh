OUT=gpurun_out; RUN=r02w; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/${RUN}_pytest.log 2>&1; tail -3 $OUT/${RUN}_pytest.log
for v in "" _old; do
LBMDEM_LIB=$PWD/2d-lbm-dem_b200/liblbmdem_gpu$v.so timeout 300 python bench.py --steps 300 --no-cpu-baseline --no-cfg5 > $OUT/${RUN}_bench$v.json 2> $OUT/${RUN}_bench$v.err
LBMDEM_LIB=$PWD/2d-lbm-dem_b200/liblbmdem_gpu$v.so timeout 300 python tools/_prof_default.py 2>&1 | grep -v Warn | head -4
done
python - <<PY
import json
for nm in ("bench","bench_old"):
    try:
        d=json.loads(open("$OUT/${RUN}_%s.json"%nm).read().strip().splitlines()[-1])
        print(nm, "MLUPS %.0f ms/step %.4f K1 %.4f frac %.3f e2e %.0f launches %d" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"]))
    except Exception as e: print(nm, e); print(open("$OUT/${RUN}_%s.err"%nm).read()[-800:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dem_coop -c 12 --csv --log-file $OUT/${RUN}_dem.csv python bench.py --steps 3 --warmup 8 --no-cpu-baseline --no-cfg5 > /dev/null 2>&1
grep dem_coop $OUT/${RUN}_dem.csv | tail -3 | cut -d, -f5,15- 
