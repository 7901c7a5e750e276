mkdir -p gpurun_out
run() {
  env $3 python bench.py --workload $1 --steps $2 --warmup 1 --no-cpu-baseline --extras none > gpurun_out/r2_56.json 2> gpurun_out/r2_56.err || tail -3 gpurun_out/r2_56.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_56.json"))
print("$1 $3", round(d["value"]), round(d["ms_per_step"],2), d["phase_ms_per_step"]["align"], round(d["roofline"]["frac"],4))
PY
}
run c2 5 X=0
run c4 2 X=0
run c4 2 TRPA_SHAPE_ORDER=asc
