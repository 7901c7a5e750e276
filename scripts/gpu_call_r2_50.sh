mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_50_pytest.log 2>&1; tail -3 gpurun_out/r2_50_pytest.log
(time python bench.py --gpus 1 --steps 20 --warmup 5) > gpurun_out/r2_50_bench_n1.json 2> gpurun_out/r2_50_bench_n1.err; tail -3 gpurun_out/r2_50_bench_n1.err
(time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5) > gpurun_out/r2_50_reference_n1.json 2> gpurun_out/r2_50_reference_n1.err; tail -3 gpurun_out/r2_50_reference_n1.err
python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/r2_50_c3.json 2> gpurun_out/r2_50_c3.err
python bench.py --workload c1 --steps 20 --warmup 5 > gpurun_out/r2_50_c1.json 2> gpurun_out/r2_50_c1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_50_launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --extras none > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
