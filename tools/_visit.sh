OUT=gpurun_out; RUN=r02t; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/${RUN}_pytest.log 2>&1; tail -3 $OUT/${RUN}_pytest.log
timeout 300 python bench.py --steps 1000 --no-cpu-baseline --no-cfg5 > $OUT/${RUN}_bench.json 2> $OUT/${RUN}_bench.err
python - <<PY
import json
d=json.loads(open("$OUT/${RUN}_bench.json").read().strip().splitlines()[-1])
print("MLUPS %.0f ms/step %.4f K1 %.4f frac %.3f e2e %.0f (%.4f ms) launches %d" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"]))
PY
