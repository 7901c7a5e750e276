mkdir -p gpurun_out
run() {
  env $3 python bench.py --workload $1 --steps $2 --warmup 2 --no-cpu-baseline --extras none > gpurun_out/r2_55.json 2> gpurun_out/r2_55.err || tail -3 gpurun_out/r2_55.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_55.json"))
print("$1 $3", round(d["value"]), round(d["ms_per_step"],2), d["phase_ms_per_step"]["align"])
PY
}
run c2 5 X=0
run c2 5 TRPA_SHAPE_ORDER=asc
run c1 20 X=0
run c1 20 TRPA_SHAPE_ORDER=asc
