#!/bin/bash
# The round's evidence run on one B200: tools/gpu_check.sh (smoke, parity tests, bench line with CPU baseline and the
# cfg5 key, ncu launch lists, full ncu captures of K1 and of the sparse kernels), then the other single-GPU workloads,
# the reference arm and the compute-sanitizer passes.
# usage: gpurun --timeout 2400 -- 'bash tools/final_check.sh <tag>'
TAG=${1:-r02Z}
OUT=gpurun_out
bash tools/gpu_check.sh $TAG cfg4
for w in cfg2 cfg3 cfg5; do
  timeout 300 python bench.py --workload $w --steps 300 --no-cpu-baseline > $OUT/${TAG}_$w.json 2> $OUT/${TAG}_$w.err
  cut -c1-160 $OUT/${TAG}_$w.json
done
timeout 300 python bench.py --steps 300 --no-cpu-baseline --no-cfg5 --strict 1 > $OUT/${TAG}_strict.json 2> $OUT/${TAG}_strict.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/${TAG}_ref.json 2> $OUT/${TAG}_ref.err
cut -c1-300 $OUT/${TAG}_ref.json
timeout 600 compute-sanitizer --tool memcheck --log-file $OUT/${TAG}_memcheck.log python __graft_entry__.py --smoke > /dev/null 2>&1; tail -1 $OUT/${TAG}_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --log-file $OUT/${TAG}_racecheck.log python __graft_entry__.py --smoke > /dev/null 2>&1; tail -1 $OUT/${TAG}_racecheck.log
