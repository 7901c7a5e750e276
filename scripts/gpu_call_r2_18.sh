mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --extras none"
run() { name=$1; shift; $B "$@" > gpurun_out/r2_18_$name.json 2> gpurun_out/r2_18_$name.err; }
run c4_la150 --workload c4 --segments 60000 --tune la_cap=150000
run c4_la75 --workload c4 --segments 60000 --tune la_cap=75000
run c4_la300 --workload c4 --segments 60000
run c5_la75 --workload c5 --tune la_cap=75000
run c5_la600 --workload c5 --tune la_cap=600000
run c2_la450 --tune la_cap=450000
