mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline --extras none"
run() { name=$1; shift; $B "$@" > gpurun_out/r2_19_$name.json 2> gpurun_out/r2_19_$name.err; }
run c4_250k_la300 --workload c4 --segments 250000
run c4_250k_la1250 --workload c4 --segments 250000 --tune la_cap=1250000
run c4_250k_la750 --workload c4 --segments 250000 --tune la_cap=750000
run c2_la550 --tune la_cap=550000
run c3_la150 --workload c3 --tune la_cap=150000
run c3_la600 --workload c3 --tune la_cap=600000
run c3_la300 --workload c3
