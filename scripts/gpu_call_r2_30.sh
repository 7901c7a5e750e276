mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k protein 2>&1 | tail -3
TRPA_PROTEIN_ROWS=2 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k protein 2>&1 | tail -2
for cfg in "QUARTER=0" "ROWS=2" "ROWS=4"; do
  echo "== $cfg"; env TRPA_PROTEIN_$cfg python scripts/perf_probe_aa.py 2>&1 | tail -3
done
for cfg in "ROWS=2" "ROWS=4"; do
  env TRPA_PROTEIN_$cfg python bench.py --workload c3 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_30_c3_$cfg.json 2> gpurun_out/r2_30_c3_$cfg.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_30_c3_$cfg.json"))
print("$cfg", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["phase_ms_per_step"])
PY
done
python -m pytest tests -m gpu -x -q -k "golden or pipeline" 2>&1 | tail -2
