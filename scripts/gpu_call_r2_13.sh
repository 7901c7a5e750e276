mkdir -p gpurun_out
(time python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/r2_13_smoke.log 2>&1; tail -2 gpurun_out/r2_13_smoke.log
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2_13_pytest.log 2>&1; tail -4 gpurun_out/r2_13_pytest.log
(time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5) > gpurun_out/r2_13_reference_n1.json 2> gpurun_out/r2_13_reference_n1.err
(time python bench.py --gpus 1 --steps 20 --warmup 5) > gpurun_out/r2_13_bench_n1.json 2> gpurun_out/r2_13_bench_n1.err
tail -c 300 gpurun_out/r2_13_bench_n1.err
