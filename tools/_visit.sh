OUT=gpurun_out; RUN=r02T; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/${RUN}_pytest.log 2>&1; tail -12 $OUT/${RUN}_pytest.log
timeout 300 python bench.py --steps 300 --no-cpu-baseline --no-cfg5 > $OUT/${RUN}_bench.json 2> $OUT/${RUN}_bench.err; echo "rc $?"
timeout 300 python bench.py --steps 300 --no-cpu-baseline --no-cfg5 --kernel 8 > $OUT/${RUN}_bench_k8.json 2> $OUT/${RUN}_bench_k8.err; echo "rc $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${RUN}_launches.csv python bench.py --steps 3 --warmup 8 --no-cpu-baseline --no-cfg5 > $OUT/${RUN}_launches.log 2>&1
python - <<PY
import json
for nm in ("bench","bench_k8"):
    try:
        d=json.loads(open("$OUT/${RUN}_%s.json"%nm).read().strip().splitlines()[-1])
        print(nm, "MLUPS %.0f ms/step %.4f K1 %.4f frac %.3f e2e %.0f launches %d" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"]))
    except Exception as e: print(nm, "unreadable", e)
PY
python tools/ncu_summary.py launches $OUT/${RUN}_launches.csv | cut -c1-130 | head -12
