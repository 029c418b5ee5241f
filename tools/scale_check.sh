#!/bin/bash
# One multi-GPU box visit: the multi-GPU parity tests (N >= 2 GPUs visible: NCCL strips == one GPU, the executable
# on two GPUs) and the bench line at N GPUs as the driver launches it (torchrun, one rank per GPU), which carries
# strip_check and the cfg5 key (BASELINE configs[4]).
# usage: gpurun --gpus N --timeout 1500 -- 'bash tools/scale_check.sh <tag> N [pytest]'
TAG=${1:-r02}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > $OUT/${TAG}_smi_${N}gpu.txt 2>&1
if [ "$3" = pytest ]; then
  timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_${N}gpu_box.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_${N}gpu_box.log
  tail -4 $OUT/${TAG}_pytest_${N}gpu_box.log
fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 300 --warmup 20 > $OUT/${TAG}_bench_${N}gpu.json 2> $OUT/${TAG}_bench_${N}gpu.err
echo "bench exit $?"; tail -3 $OUT/${TAG}_bench_${N}gpu.err
python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_bench_${N}gpu.json").read().strip().splitlines()[-1])
print("N=%d MLUPS %.0f ms/step %.4f e2e %.0f K1 %.4f strip_check %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["avg_launch_ms"], d.get("strip_check", {}).get("ok")))
c = d.get("cfg5")
if c: print("cfg5: MLUPS %.0f ms/step %.4f strip_check %s" % (c["value"], c["ms_per_step"], (c.get("strip_check") or {}).get("ok")))
PY
