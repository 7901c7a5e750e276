mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --extras none"
run() { name=$1; shift; $B "$@" > gpurun_out/r2_07_$name.json 2> gpurun_out/r2_07_$name.err; }
run c5_base --workload c5
run c5_h4 --workload c5 --tune wedge_hint_s8=4
run c5_h6 --workload c5 --tune wedge_hint_s8=6
run c5_h7 --workload c5 --tune wedge_hint_s8=7
run c4_base --workload c4 --segments 60000
run c4_k4096 --workload c4 --segments 60000 --tune wedge_max_k=4096
run c4_k4096s8 --workload c4 --segments 60000 --tune wedge_max_k=4096 --tune wedge_s8=8
run c2_base
run c2_s8 --tune wedge_s8=8
