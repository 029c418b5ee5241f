OUT=gpurun_out; RUN=r02z; mkdir -p $OUT
export LBMDEM_PEER_SUMS=1 LBMDEM_VERBOSE=1
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 300 --warmup 20 --no-cfg5 > $OUT/${RUN}_bench_ipc_${N}gpu.json 2> $OUT/${RUN}_bench_ipc_${N}gpu.err
grep -m1 "peer-memory" $OUT/${RUN}_bench_ipc_${N}gpu.err
python - <<PY
import json
for nm in ("bench_ipc_${N}gpu",):
    try:
        d = json.loads(open("$OUT/${RUN}_%s.json"%nm).read().strip().splitlines()[-1])
        print(nm, "N=%d MLUPS %.0f ms/step %.4f e2e %.0f launches %d strip_check %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d.get("strip_check", {}).get("ok")))
    except Exception as e: print(nm, e); print(open("$OUT/${RUN}_%s.err"%nm).read()[-1500:])
PY
