mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_61_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --extras none > gpurun_out/r2_61_ncu_bench.log 2>&1
tail -c 300 gpurun_out/r2_61_ncu_bench.log
