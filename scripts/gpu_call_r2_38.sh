mkdir -p gpurun_out
run() {
  python bench.py --workload $1 --steps 3 --warmup 2 --no-cpu-baseline --extras none $2 > gpurun_out/r2_38.json 2> gpurun_out/r2_38.err || tail -3 gpurun_out/r2_38.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_38.json"))
print("$1 $2", round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), d["rounds_per_step"], d["pairs_launched_per_step"])
PY
}
run c3 ""
run c3 "--tune pipes=2"
run c3 "--tune pipes=2 --tune la_cap=300000"
run c3 "--tune pipes=3 --tune la_cap=300000"
run c2 ""
run c2 "--tune pipes=2"
