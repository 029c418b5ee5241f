#!/bin/bash
# One GPU-box visit for tuning: parity tests on the default build, then the same short bench on every
# tagged variant library (2d-lbm-dem_b200/liblbmdem_gpu_<tag>.so, built by `build.py --tag=<tag> -D...`).
# usage: gpurun --timeout 900 -- 'bash tools/variants.sh <runtag> <workload> <tag> [<tag> ...]'
RUN=${1:-r01V}; WORK=${2:-cfg4}; shift 2
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${RUN}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${RUN}_pytest.log
tail -5 $OUT/${RUN}_pytest.log
timeout 200 python bench.py --workload $WORK --steps 300 --no-cpu-baseline > $OUT/${RUN}_default.json 2> $OUT/${RUN}_default.err
# cross-check paths of the default library on request: KERNELS="2" -> bench.py --kernel 2 (per-grain rasteriser)
for k in $KERNELS; do
  timeout 200 python bench.py --workload $WORK --steps 300 --no-cpu-baseline --kernel $k > $OUT/${RUN}_kernel$k.json 2> $OUT/${RUN}_kernel$k.err
done
for t in "$@"; do
  LBMDEM_LIB=$PWD/2d-lbm-dem_b200/liblbmdem_gpu_$t.so timeout 200 python bench.py --workload $WORK --steps 300 --no-cpu-baseline \
      > $OUT/${RUN}_$t.json 2> $OUT/${RUN}_$t.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${RUN}_launches.csv \
    python bench.py --workload $WORK --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${RUN}_launches.log 2>&1
# optional: other workloads ($EXTRA_WORK) on the default library and on the tags in $EXTRA_TAGS
for w in $EXTRA_WORK; do
  timeout 200 python bench.py --workload $w --steps 200 --no-cpu-baseline > $OUT/${RUN}_${w}_default.json 2> $OUT/${RUN}_${w}_default.err
  for t in $EXTRA_TAGS; do
    LBMDEM_LIB=$PWD/2d-lbm-dem_b200/liblbmdem_gpu_$t.so timeout 200 python bench.py --workload $w --steps 200 --no-cpu-baseline \
        > $OUT/${RUN}_${w}_$t.json 2> $OUT/${RUN}_${w}_$t.err
  done
done
# optional: one full ncu capture of K1 per tag in $NCU_TAGS ("default" = the shipped library)
for t in $NCU_TAGS; do
  L=$PWD/2d-lbm-dem_b200/liblbmdem_gpu_$t.so; [ "$t" = default ] && L=$PWD/2d-lbm-dem_b200/liblbmdem_gpu.so
  LBMDEM_LIB=$L timeout 300 ncu --set full --clock-control none --import-source on -k regex:${NCU_K:-lbm_rows} -s 4 -c 1 -f -o $OUT/${RUN}_k1_$t \
      python bench.py --workload $WORK --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${RUN}_k1_$t.log 2>&1
done
python - <<PY
import json,glob
for p in sorted(glob.glob("$OUT/${RUN}_*.json")):
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        print(p.split("${RUN}_")[1][:-5].ljust(10), "MLUPS %.0f  ms/step %.4f  K1 ms %.4f frac %.3f  e2e %.0f" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"]))
    except Exception as e:
        print(p, "unreadable", e)
PY
