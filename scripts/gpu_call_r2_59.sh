mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "kernels or pipeline or golden or edge or random_workload" 2>&1 | tail -2
run() {
  python bench.py --workload $1 --steps $2 --warmup 3 --no-cpu-baseline --extras none > gpurun_out/r2_59.json 2> gpurun_out/r2_59.err || tail -3 gpurun_out/r2_59.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_59.json"))
print("$1", round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["phase_ms_per_step"])
PY
}
run c1 30
run c2 5
