mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -m gpu -x -q -k "protein or aa" 2>&1 | tail -1
TRPA_PROTEIN_ALT=2 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -m gpu -x -q -k "protein or aa" 2>&1 | tail -1
run() {
  env $1 python bench.py --workload c3 --steps 3 --warmup 2 --no-cpu-baseline $2 > gpurun_out/r2_35_c3.json 2> gpurun_out/r2_35_c3.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_35_c3.json"))
print("$1 $2", round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), round(d["roofline"]["frac"],4), d["phase_ms_per_step"], d["rounds_per_step"], d["pairs_launched_per_step"])
PY
}
run X=0 ""
run TRPA_PROTEIN_ALT=2 ""
run X=0 "--tune la_cap=150000"
run X=0 "--tune la_cap=400000"
run X=0 "--tune la_cap=600000"
run X=0 "--tune la_cap=1000000"
run X=0 "--tune la_cap=600000 --tune la_max=64"
