mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "kernels or pipeline or golden or edge" 2>&1 | tail -2
run() {
  python bench.py --workload $1 --steps $3 --warmup 2 --no-cpu-baseline --extras none $2 > gpurun_out/r2_45.json 2> gpurun_out/r2_45.err || tail -3 gpurun_out/r2_45.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_45.json"))
print("$1 $2", round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), d["rounds_per_step"], d["phase_ms_per_step"], round(d["roofline"]["frac"],4))
PY
}
run c2 "" 5
run c1 "" 20
