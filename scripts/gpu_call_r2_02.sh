mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2_02_pytest.log 2>&1
tail -5 gpurun_out/r2_02_pytest.log
(time python bench.py --steps 3 --warmup 3 --big-scale 0.02) > gpurun_out/r2_02_bench_small.json 2> gpurun_out/r2_02_bench_small.err
tail -c 600 gpurun_out/r2_02_bench_small.err
