mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -m gpu -x -q -k "protein or aa" 2>&1 | tail -1
for q in 1 2 4 8 16; do
  echo "== QPW=$q"; TRPA_PROTEIN_QPW=$q python scripts/probe_aa300.py 3 2>&1 | tail -1
  TRPA_PROTEIN_QPW=$q python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline --extras none > gpurun_out/r2_47_c3.json 2> gpurun_out/r2_47_c3.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_47_c3.json"))
print(round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), round(d["roofline"]["frac"],4), d["phase_ms_per_step"])
PY
done
