mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2_11_pytest.log 2>&1
tail -4 gpurun_out/r2_11_pytest.log
B="python bench.py --no-cpu-baseline --extras none"
$B --workload c1 --steps 20 --warmup 5 > gpurun_out/r2_11_c1.json 2> gpurun_out/r2_11_c1.err
$B --workload c1 --steps 20 --warmup 5 --tune tax_smem=0 > gpurun_out/r2_11_c1_nosmem.json 2> gpurun_out/r2_11_c1_nosmem.err
$B --steps 3 --warmup 2 > gpurun_out/r2_11_c2.json 2> gpurun_out/r2_11_c2.err
TRPA_DEBUG_TIMING=1 $B --workload c3 --steps 5 --warmup 3 > gpurun_out/r2_11_c3.json 2> gpurun_out/r2_11_c3.err
$B --workload c3 --steps 5 --warmup 3 --tune tax_smem=0 > gpurun_out/r2_11_c3_nosmem.json 2> gpurun_out/r2_11_c3_nosmem.err
(time compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_binner.py tests/test_verbose_log.py tests/test_gpu_edge_cases.py -m gpu -x -q -k "random or rejects or verbose or edge or single") > gpurun_out/r2_11_sanitizer.log 2>&1
tail -6 gpurun_out/r2_11_sanitizer.log
