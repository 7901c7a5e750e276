mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "protein or aa or random_workload or edge or pipeline" 2>&1 | tail -2
python scripts/perf_probe_aa.py 2>&1 | tail -4
python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline --extras none > gpurun_out/r2_49_c3.json 2> gpurun_out/r2_49_c3.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_49_c3.json"))
print(round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), round(d["roofline"]["frac"],4), d["phase_ms_per_step"], d["gpu_launches"])
PY
