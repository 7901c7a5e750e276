mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:decide_kernel -s 9 -c 1 -o gpurun_out/r2_40_decide python bench.py --workload c2 --steps 1 --warmup 0 --no-cpu-baseline --extras none 2>&1 | tail -2
ncu -i gpurun_out/r2_40_decide.ncu-rep --page raw --csv > gpurun_out/r2_40_decide_raw.csv
ncu -i gpurun_out/r2_40_decide.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_40_decide_src.csv
ncu -i gpurun_out/r2_40_decide.ncu-rep --page source --csv > gpurun_out/r2_40_decide_cuda.csv
rm -f gpurun_out/r2_40_decide.ncu-rep
