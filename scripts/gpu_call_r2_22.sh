mkdir -p gpurun_out
(time python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --extras c4,c5) > gpurun_out/r2_22_bench_n1.json 2> gpurun_out/r2_22_bench_n1.err
python -m pytest tests -m gpu -x -q -k "fullsize or golden" 2>&1 | tail -2
