mkdir -p gpurun_out
run() {
  python bench.py --workload $1 --steps 3 --warmup 2 --no-cpu-baseline --extras none $2 > gpurun_out/r2_39.json 2> gpurun_out/r2_39.err || tail -3 gpurun_out/r2_39.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_39.json"))
print("$1 $2", round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), d["rounds_per_step"], d["pairs_launched_per_step"])
PY
}
run c3 "--tune pipes=2 --tune la_cap=150000"
run c3 "--tune pipes=3 --tune la_cap=150000"
run c3 "--tune pipes=4 --tune la_cap=150000"
run c3 "--tune pipes=4 --tune la_cap=250000"
run c1 ""
run c1 "--tune pipes=2"
