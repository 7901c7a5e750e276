mkdir -p gpurun_out
run() {
  python bench.py --workload c3 --steps 5 --warmup 2 --no-cpu-baseline --extras none $1 > gpurun_out/r2_62.json 2> gpurun_out/r2_62.err || tail -3 gpurun_out/r2_62.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_62.json"))
print("$1", round(d["value"]), round(d["ms_per_step"],2), d["rounds_per_step"], d["pairs_launched_per_step"])
PY
}
run "--tune la_cap=100000"
run ""
run "--tune la_cap=200000"
run "--tune la_cap=250000"
