#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, one full capture of K1.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag]'
TAG=${1:-r01}
KREGEX=${2:-lbm_rows}
WORK=${3:-cfg4}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
nproc > $OUT/${TAG}_host.txt; grep -m1 'model name' /proc/cpuinfo >> $OUT/${TAG}_host.txt
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.log; tail -2 $OUT/${TAG}_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --workload $WORK > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
cat $OUT/${TAG}_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --workload $WORK --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lbm_rows -s 4 -c 2 -f -o $OUT/${TAG}_k1 \
    python bench.py --workload $WORK --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_k1.log 2>&1
if [ "$KREGEX" != "lbm_rows" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 8 -c 4 -f -o $OUT/${TAG}_aux \
    python bench.py --workload $WORK --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_aux.log 2>&1
fi
ls -la $OUT
