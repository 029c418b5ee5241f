OUT=gpurun_out; RUN=r02b; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/${RUN}_pytest.log 2>&1; tail -5 $OUT/${RUN}_pytest.log
timeout 300 python bench.py --steps 300 --no-cpu-baseline --no-cfg5 --strict 1 > $OUT/${RUN}_strict.json 2> $OUT/${RUN}_strict.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${RUN}_launches.csv python bench.py --steps 3 --warmup 8 --no-cpu-baseline --no-cfg5 --strict 1 > $OUT/${RUN}_launches.log 2>&1
python - <<PY
import json
d=json.loads(open("$OUT/${RUN}_strict.json").read().strip().splitlines()[-1])
print("strict MLUPS %.0f ms/step %.4f K1 %.4f frac %.3f e2e %.0f" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"]))
PY
python tools/ncu_summary.py launches $OUT/${RUN}_launches.csv | cut -c1-120 | sed -n 3,12p
