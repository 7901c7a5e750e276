mkdir -p gpurun_out
for v in "ALT=4" "ALT=3"; do
  env TRPA_PROTEIN_$v python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -m gpu -x -q -k "protein or aa" 2>&1 | tail -1
done
python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q -k "protein or aa" 2>&1 | tail -1
for v in "X=0" "ALT=4" "ALT=3"; do
  echo "== $v"; env TRPA_PROTEIN_$v python scripts/probe_aa300.py 3 2>&1 | tail -1
  env TRPA_PROTEIN_$v python bench.py --workload c3 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_34_c3_$v.json 2> gpurun_out/r2_34_c3_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_34_c3_$v.json"))
print("$v", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["phase_ms_per_step"])
PY
done
