mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py tests/test_gpu_fuzz.py tests/test_gpu_pipeline.py -m gpu -x -q -k "protein or seqan_vectors or random_workload or aa" 2>&1 | tail -2
python scripts/perf_probe_aa.py 2>&1 | tail -4
python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/r2_52_c3.json 2> gpurun_out/r2_52_c3.err
python - <<PY
import json
c=json.load(open('gpurun_out/r2_52_c3.json')); print('c3', c['value'], c['ms_per_step'], c['e2e']['value'], c['roofline']['frac'], c['roofline'].get('frac_r01_constant'), c['phase_ms_per_step'], c['cpu_baseline']['value'], c['cpu_baseline']['gff3_identical_to_gpu'])
PY
