mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --extras none"
run() { name=$1; shift; $B "$@" > gpurun_out/r2_16_$name.json 2> gpurun_out/r2_16_$name.err; }
run c2_base
run c2_setup1000 --tune cost_setup=1000
run c2_setup3000 --tune cost_setup=3000
run c2_step400 --tune cost_step=400
run c2_step400_setup1500 --tune cost_step=400 --tune cost_setup=1500
run c2_la150 --tune la_cap=150000
run c2_la600 --tune la_cap=600000
run c2_lamax64 --tune la_max=64 --tune la_cap=600000
run c4_base --workload c4 --segments 60000
run c4_setup1500 --workload c4 --segments 60000 --tune cost_setup=1500 --tune cost_step=400
run c4_la600 --workload c4 --segments 60000 --tune la_cap=600000
