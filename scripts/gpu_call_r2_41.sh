mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "pipeline or golden or edge or fullsize" 2>&1 | tail -2
run() {
  python bench.py --workload $1 --steps 3 --warmup 2 --no-cpu-baseline --extras none $2 > gpurun_out/r2_41.json 2> gpurun_out/r2_41.err || tail -3 gpurun_out/r2_41.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_41.json"))
print("$1 $2", round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), d["rounds_per_step"], d["phase_ms_per_step"])
PY
}
run c1 ""
run c2 ""
run c3 ""
