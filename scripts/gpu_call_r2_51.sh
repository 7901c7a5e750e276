mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/r2_51_pytest.log 2>&1; tail -3 gpurun_out/r2_51_pytest.log
(time compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py tests/test_gpu_fuzz.py -m gpu -x -q -k "protein or seqan_vectors or random_workload") > gpurun_out/r2_51_sanitizer_memcheck.log 2>&1
tail -5 gpurun_out/r2_51_sanitizer_memcheck.log
python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/r2_51_c3.json 2> gpurun_out/r2_51_c3.err
python - <<PY
import json
c=json.load(open('gpurun_out/r2_51_c3.json')); print('c3', c['value'], c['ms_per_step'], c['e2e']['value'], c['roofline']['frac'], c['roofline'].get('frac_r01_constant'), c['phase_ms_per_step'], c['cpu_baseline']['value'], c['cpu_baseline']['gff3_identical_to_gpu'])
PY
