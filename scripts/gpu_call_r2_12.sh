mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q) > gpurun_out/r2_12_pytest.log 2>&1
tail -6 gpurun_out/r2_12_pytest.log
TRPA_DEBUG_TIMING=1 python bench.py --no-cpu-baseline --extras none --workload c3 --steps 5 --warmup 3 > gpurun_out/r2_12_c3.json 2> gpurun_out/r2_12_c3.err
grep predict_batch gpurun_out/r2_12_c3.err | tail -2
