mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2_01_pytest.log 2>&1
for p in 1 2; do python bench.py --steps 5 --warmup 3 --no-cpu-baseline --tune pipes=$p > gpurun_out/r2_01_c2_p$p.json 2> gpurun_out/r2_01_c2_p$p.err; done
python bench.py --workload c4 --segments 60000 --steps 2 --warmup 3 --no-cpu-baseline --tune pipes=1 > gpurun_out/r2_01_c4_p1.json 2> gpurun_out/r2_01_c4_p1.err
python bench.py --workload c4 --segments 60000 --steps 2 --warmup 3 --no-cpu-baseline --tune pipes=2 > gpurun_out/r2_01_c4_p2.json 2> gpurun_out/r2_01_c4_p2.err
python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_01_c1.json 2> gpurun_out/r2_01_c1.err
tail -3 gpurun_out/r2_01_pytest.log
