mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "pipeline or golden or edge or random_workload" 2>&1 | tail -2
run() {
  python bench.py --workload $1 --steps $2 --warmup 3 --no-cpu-baseline --extras none > gpurun_out/r2_53.json 2> gpurun_out/r2_53.err || tail -3 gpurun_out/r2_53.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_53.json"))
print("$1", round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), d["rounds_per_step"], d["phase_ms_per_step"])
PY
}
run c1 20
run c2 5
run c3 10
