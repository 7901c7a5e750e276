mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --extras none"
run() { name=$1; shift; $B "$@" > gpurun_out/r2_06_$name.json 2> gpurun_out/r2_06_$name.err; }
run c2_base
run c2_s7 --tune wedge_s8=7
run c2_s8 --tune wedge_s8=8
run c2_e15 --tune wedge_e0=15
run c2_s7c8 --tune wedge_s8=7 --tune wedge_cushion=8
run c4_base --workload c4 --segments 60000
run c4_s7 --workload c4 --segments 60000 --tune wedge_s8=7
run c4_s8 --workload c4 --segments 60000 --tune wedge_s8=8
run c5_nonoise --workload c5 --tune hint_noise=0
run c5_noise --workload c5
run c4_nonoise --workload c4 --segments 60000 --tune hint_noise=0
ls gpurun_out | grep r2_06 | wc -l
