mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:plan_kernel -s 6 -c 1 -o gpurun_out/r2_44_plan python bench.py --workload c2 --steps 1 --warmup 0 --no-cpu-baseline --extras none 2>&1 | tail -1
ncu -i gpurun_out/r2_44_plan.ncu-rep --page raw --csv > gpurun_out/r2_44_plan_raw.csv
ncu -i gpurun_out/r2_44_plan.ncu-rep --page source --csv > gpurun_out/r2_44_plan_cuda.csv
rm -f gpurun_out/r2_44_plan.ncu-rep
