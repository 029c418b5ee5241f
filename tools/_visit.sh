OUT=gpurun_out; RUN=r02k; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/${RUN}_pytest.log 2>&1; tail -4 $OUT/${RUN}_pytest.log
timeout 300 python bench.py --steps 300 --no-cpu-baseline > $OUT/${RUN}_bench.json 2> $OUT/${RUN}_bench.err
for w in cfg2 cfg3; do timeout 300 python bench.py --workload $w --steps 300 --no-cpu-baseline > $OUT/${RUN}_$w.json 2> $OUT/${RUN}_$w.err; done
python - <<PY
import json
for nm in ("bench","cfg2","cfg3"):
    try:
        d=json.loads(open("$OUT/${RUN}_%s.json"%nm).read().strip().splitlines()[-1])
        print(nm, "MLUPS %.0f ms/step %.4f K1 %.4f frac %.3f e2e %.0f launches %d" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"]), "cfg5", (d.get("cfg5") or {}).get("value"), ((d.get("cfg5") or {}).get("roofline") or {}).get("frac"))
    except Exception as e: print(nm, e); print(open("$OUT/${RUN}_%s.err"%nm).read()[-600:])
PY
