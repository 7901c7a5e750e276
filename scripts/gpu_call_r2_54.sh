mkdir -p gpurun_out
run() {
  env TRPA_DECIDE_PER_WARP=$3 python bench.py --workload $1 --steps $2 --warmup 3 --no-cpu-baseline --extras none > gpurun_out/r2_54.json 2> gpurun_out/r2_54.err || tail -3 gpurun_out/r2_54.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_54.json"))
print("$1 per_warp=$3", round(d["value"]), round(d["ms_per_step"],2), d["phase_ms_per_step"]["decide"])
PY
}
for pw in 32 8 4 2 1; do run c1 20 $pw; run c3 5 $pw; run c2 3 $pw; done
