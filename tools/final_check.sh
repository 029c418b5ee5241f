#!/bin/bash
# The round's evidence run on one B200: tools/gpu_check.sh (smoke, parity tests, bench line with CPU baseline, ncu launch
# list, full ncu capture of K1 and of the sparse kernels), then the other single-GPU workloads and the reference arm.
# usage: gpurun --timeout 1500 -- 'bash tools/final_check.sh <tag>'
TAG=${1:-r01Z}
OUT=gpurun_out
bash tools/gpu_check.sh $TAG 'raster_tile|bounce_sweep' cfg4
for w in cfg2 cfg3 cfg5; do
  timeout 300 python bench.py --workload $w --steps 300 --no-cpu-baseline > $OUT/${TAG}_$w.json 2> $OUT/${TAG}_$w.err
  cut -c1-160 $OUT/${TAG}_$w.json
done
timeout 400 python bench.py --impl reference --steps 20 --warmup 1 > $OUT/${TAG}_ref.json 2> $OUT/${TAG}_ref.err
cut -c1-300 $OUT/${TAG}_ref.json
