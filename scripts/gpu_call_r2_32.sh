mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:protein3_kernel -s 0 -c 1 -o gpurun_out/r2_32_protein3 python scripts/probe_aa300.py 1 2>&1 | tail -3
ncu -i gpurun_out/r2_32_protein3.ncu-rep --page raw --csv > gpurun_out/r2_32_protein3_raw.csv
ncu -i gpurun_out/r2_32_protein3.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_32_protein3_src.csv
ls -la gpurun_out/r2_32*
