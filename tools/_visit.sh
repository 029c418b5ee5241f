OUT=gpurun_out; RUN=r02F; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_local_strips.py -m gpu -x -q > $OUT/${RUN}_local.log 2>&1; tail -15 $OUT/${RUN}_local.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/${RUN}_pytest.log 2>&1; tail -6 $OUT/${RUN}_pytest.log
