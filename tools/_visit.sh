OUT=gpurun_out; RUN=r02E; mkdir -p $OUT
nvidia-smi -L | head -3
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${RUN}_pytest.log 2>&1; tail -4 $OUT/${RUN}_pytest.log
timeout 600 python bench.py --steps 100 --warmup 10 > $OUT/${RUN}_bench1.json 2> $OUT/${RUN}_bench1.err; echo "bench1 rc $?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/${RUN}_ref1.json 2> $OUT/${RUN}_ref1.err; echo "ref1 rc $?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 10 > $OUT/${RUN}_bench2.json 2> $OUT/${RUN}_bench2.err; echo "bench2 rc $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --impl reference --steps 20 --warmup 5 > $OUT/${RUN}_ref2.json 2> $OUT/${RUN}_ref2.err; echo "ref2 rc $?"
cut -c1-1500 $OUT/${RUN}_bench1.json; tail -3 $OUT/${RUN}_bench1.err
python - <<PY
import json
for nm in ("bench1","ref1","bench2","ref2"):
    try:
        d=json.loads(open("$OUT/${RUN}_%s.json"%nm).read().strip().splitlines()[-1])
        print(nm, "value %.0f e2e %.0f" % (d["value"], d["e2e"]["value"]), "cores", d.get("cpu_baseline",{}).get("cores"), "check", d.get("strip_check"), "cfg5", {k:d.get("cfg5",{}).get(k) for k in ("value","ms_per_step","strip_check","failed")}, d["config"]["lattice"], d["data"])
    except Exception as e:
        print(nm, "unreadable", e)
PY
