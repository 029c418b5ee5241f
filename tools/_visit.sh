OUT=gpurun_out; RUN=r02X; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/${RUN}_pytest.log 2>&1; tail -8 $OUT/${RUN}_pytest.log
for n in 2 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 100 --warmup 10 > $OUT/${RUN}_bench$n.json 2> $OUT/${RUN}_bench$n.err; echo "bench$n rc $?"
done
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${RUN}_bench1.json 2> $OUT/${RUN}_bench1.err
python - <<PY
import json
for nm in ("bench1","bench2","bench4"):
    try:
        d=json.loads(open("$OUT/${RUN}_%s.json"%nm).read().strip().splitlines()[-1])
        print(nm, "value %.0f ms %.4f e2e %.0f K1 %.4f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["avg_launch_ms"]), "check", d.get("strip_check",{}).get("ok"), "cfg5 %.0f ms %.4f e2e %.0f check %s" % (d["cfg5"]["value"], d["cfg5"]["ms_per_step"], d["cfg5"]["e2e"]["value"], d["cfg5"].get("strip_check",{}).get("ok")))
    except Exception as e:
        print(nm, "unreadable", e)
PY
