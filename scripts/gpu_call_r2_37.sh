mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2_37_pytest.log 2>&1; tail -3 gpurun_out/r2_37_pytest.log
python bench.py --workload c3 --steps 5 --warmup 3 > gpurun_out/r2_37_c3.json 2> gpurun_out/r2_37_c3.err; tail -c 400 gpurun_out/r2_37_c3.err
(time python bench.py --steps 5 --warmup 3) > gpurun_out/r2_37_bench_n1.json 2> gpurun_out/r2_37_bench_n1.err; tail -4 gpurun_out/r2_37_bench_n1.err
