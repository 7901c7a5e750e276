mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/r2_58_pytest.log 2>&1; tail -3 gpurun_out/r2_58_pytest.log
(time python bench.py --gpus 1 --steps 20 --warmup 5) > gpurun_out/r2_58_bench_n1.json 2> gpurun_out/r2_58_bench_n1.err; tail -3 gpurun_out/r2_58_bench_n1.err
(time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5) > gpurun_out/r2_58_reference_n1.json 2> gpurun_out/r2_58_reference_n1.err; tail -3 gpurun_out/r2_58_reference_n1.err
python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/r2_58_c3.json 2> gpurun_out/r2_58_c3.err
python bench.py --workload c1 --steps 20 --warmup 5 > gpurun_out/r2_58_c1.json 2> gpurun_out/r2_58_c1.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
