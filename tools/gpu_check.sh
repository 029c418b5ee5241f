#!/bin/bash
# One GPU-box visit: smoke, parity tests, bench line, ncu launch list, full captures of K1 and of the rim kernel.
# usage: gpurun --timeout 1800 -- 'bash tools/gpu_check.sh [tag]'
TAG=${1:-r02}
WORK=${2:-cfg4}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
nproc > $OUT/${TAG}_host.txt; grep -m1 'model name' /proc/cpuinfo >> $OUT/${TAG}_host.txt
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.log; tail -2 $OUT/${TAG}_smoke.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
timeout 900 python bench.py --workload $WORK > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
cut -c1-400 $OUT/${TAG}_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --workload $WORK --steps 3 --warmup 8 --no-cpu-baseline --no-cfg5 > $OUT/${TAG}_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lbm_rows -s 6 -c 2 -f -o $OUT/${TAG}_k1 \
    python bench.py --workload $WORK --steps 3 --warmup 8 --no-cpu-baseline --no-cfg5 > $OUT/${TAG}_k1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rim_kernel|raster_tile|grain_bin|dem_coop' -s 16 -c 8 -f -o $OUT/${TAG}_aux \
    python bench.py --workload $WORK --steps 3 --warmup 8 --no-cpu-baseline --no-cfg5 > $OUT/${TAG}_aux.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_strips_launches.csv \
    python tools/strip_launches.py 2 6 > $OUT/${TAG}_strips_launches.log 2>&1
ls -la $OUT | tail -20
