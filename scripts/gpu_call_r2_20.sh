mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --extras none"
run() { name=$1; shift; $B "$@" > gpurun_out/r2_20_$name.json 2> gpurun_out/r2_20_$name.err; }
run c2
run c3 --workload c3
run c1 --workload c1 --steps 20 --warmup 5
run c4_60k --workload c4 --segments 60000
run c4_250k --workload c4 --segments 250000 --steps 2
run c5 --workload c5
python -m pytest tests -m gpu -x -q -k "fullsize or golden or pipeline" 2>&1 | tail -2
