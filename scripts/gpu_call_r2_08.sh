mkdir -p gpurun_out
(time python bench.py --steps 5 --warmup 3) > gpurun_out/r2_08_bench_full_n1.json 2> gpurun_out/r2_08_bench_full_n1.err
tail -c 400 gpurun_out/r2_08_bench_full_n1.err
# ncu launch list of the same command (headline only), then a full capture of the dominant kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_08_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --extras none > gpurun_out/r2_08_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:myers3_kernel -s 120 -c 4 -o gpurun_out/r2_08_myers3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --extras none > gpurun_out/r2_08_ncu_full.log 2>&1
ls -la gpurun_out | grep r2_08
