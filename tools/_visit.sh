OUT=gpurun_out; RUN=r02v; mkdir -p $OUT
for v in lds5 lds6 ldsn; do
LBMDEM_LIB=$PWD/2d-lbm-dem_b200/liblbmdem_gpu_$v.so timeout 300 python bench.py --steps 300 --no-cpu-baseline --no-cfg5 > $OUT/${RUN}_$v.json 2> $OUT/${RUN}_$v.err
done
python - <<PY
import json
for nm in ("lds5","lds6","ldsn"):
    try:
        d=json.loads(open("$OUT/${RUN}_%s.json"%nm).read().strip().splitlines()[-1])
        print(nm, "MLUPS %.0f ms/step %.4f K1 %.4f frac %.3f e2e %.0f launches %d" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"]))
    except Exception as e: print(nm, e); print(open("$OUT/${RUN}_%s.err"%nm).read()[-800:])
PY
