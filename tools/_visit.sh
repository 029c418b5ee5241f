OUT=gpurun_out; RUN=r02s; mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${RUN}_strips8.csv python tools/strip_launches.py 8 4 > $OUT/${RUN}_strips8.log 2>&1
tail -3 $OUT/${RUN}_strips8.log
python - <<PY
import csv, collections, re
rows=[l for l in open("$OUT/${RUN}_strips8.csv") if l.startswith('"')]
agg=collections.defaultdict(list)
for x in csv.DictReader(rows):
    try: agg[re.sub(r"\\(.*","",x["Kernel Name"])[:50]].append(float(x["Metric Value"].replace(",","")))
    except Exception: pass
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1]))[:24]: print("  %-50s n=%4d avg %9.1f total %10.1f"%(k,len(v),sum(v)/len(v),sum(v)))
PY
