OUT=gpurun_out; RUN=r02r; mkdir -p $OUT
for g in 0 32 64 128; do
  if [ $g = 0 ]; then export LBMDEM_VERBOSE=1; unset LBMDEM_L2_FETCH; else export LBMDEM_L2_FETCH=$g; fi
  timeout 300 python bench.py --steps 300 --no-cpu-baseline --no-cfg5 > $OUT/${RUN}_g$g.json 2> $OUT/${RUN}_g$g.err
  grep -m1 "L2 fetch" $OUT/${RUN}_g$g.err
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:"rim_kernel|lbm_rows|raster_tile" -c 30 --csv --log-file $OUT/${RUN}_launches_g$g.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-cfg5 > $OUT/${RUN}_ncu_g$g.log 2>&1
done
python - <<PY
import json, csv, collections, re
for g in (0,32,64,128):
    try:
        d=json.loads(open("$OUT/${RUN}_g%d.json"%g).read().strip().splitlines()[-1])
        print(g, "MLUPS %.0f ms/step %.4f K1 %.4f frac %.3f e2e %.0f" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"]))
    except Exception as e: print(g, e); print(open("$OUT/${RUN}_g%d.err"%g).read()[-600:])
    rows=[l for l in open("$OUT/${RUN}_launches_g%d.csv"%g) if l.startswith('"')]
    agg=collections.defaultdict(list)
    for x in csv.DictReader(rows):
        try: agg[(re.sub(r"\\(.*","",x["Kernel Name"])[:40], x["Metric Name"])].append(float(x["Metric Value"].replace(",","")))
        except Exception: pass
    for k,v in sorted(agg.items()): print("   %-40s %-28s n=%3d avg %.1f"%(k[0],k[1],len(v),sum(v)/len(v)))
PY
