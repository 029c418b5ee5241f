OUT=gpurun_out; RUN=r02S; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/${RUN}_pytest.log 2>&1; tail -6 $OUT/${RUN}_pytest.log
timeout 300 python bench.py --steps 300 --no-cpu-baseline --no-cfg5 > $OUT/${RUN}_bench.json 2> $OUT/${RUN}_bench.err; echo "rc $?"
timeout 900 compute-sanitizer --tool memcheck --log-file $OUT/${RUN}_memcheck.log python __graft_entry__.py --smoke > $OUT/${RUN}_memcheck.out 2>&1; echo "memcheck rc $?"; tail -3 $OUT/${RUN}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --log-file $OUT/${RUN}_racecheck.log python __graft_entry__.py --smoke > $OUT/${RUN}_racecheck.out 2>&1; echo "racecheck rc $?"; tail -3 $OUT/${RUN}_racecheck.log
python - <<PY
import json
d=json.loads(open("$OUT/${RUN}_bench.json").read().strip().splitlines()[-1])
print("MLUPS %.0f ms/step %.4f K1 %.4f frac %.3f e2e %.0f launches %d" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"]))
PY
