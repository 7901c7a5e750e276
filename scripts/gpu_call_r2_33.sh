mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:protein3_kernel -s 0 -c 2 -o gpurun_out/r2_33_protein3_c3 python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --extras none 2>&1 | tail -2
ncu -i gpurun_out/r2_33_protein3_c3.ncu-rep --page raw --csv > gpurun_out/r2_33_protein3_c3_raw.csv
ncu -i gpurun_out/r2_33_protein3_c3.ncu-rep --page source --csv --print-source sass --launch-skip 0 --launch-count 1 > gpurun_out/r2_33_protein3_c3_src.csv
rm -f gpurun_out/r2_33_protein3_c3.ncu-rep
