mkdir -p gpurun_out
run() {
  env $1 python bench.py --workload c3 --steps 3 --warmup 2 --no-cpu-baseline $2 > gpurun_out/r2_36_c3.json 2> gpurun_out/r2_36_c3.err || tail -3 gpurun_out/r2_36_c3.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_36_c3.json"))
print("$1 $2", round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), round(d["roofline"]["frac"],4), d["phase_ms_per_step"], d["rounds_per_step"], d["pairs_launched_per_step"])
PY
}
run X=0 "--tune la_cap=100000"
run X=0 "--tune la_cap=75000"
run X=0 "--tune pipes=2"
run X=0 "--tune pipes=2 --tune la_cap=300000"
run X=0 "--tune pipes=3 --tune la_cap=300000"
run X=0 "--tune pipes=2 --tune la_cap=200000"
