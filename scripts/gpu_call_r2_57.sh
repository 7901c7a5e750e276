mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_57_launches_c1.csv python bench.py --workload c1 --steps 2 --warmup 1 --no-cpu-baseline --extras none > /dev/null 2>&1
TRPA_DEBUG=1 python bench.py --workload c1 --steps 1 --warmup 0 --no-cpu-baseline --extras none 2>&1 | grep "trpa\]" | head -40
