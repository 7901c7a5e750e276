mkdir -p gpurun_out
(time compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -m gpu -x -q -k "protein or seed3aa") > gpurun_out/r2_43_sanitizer_memcheck.log 2>&1
tail -6 gpurun_out/r2_43_sanitizer_memcheck.log
(time compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q -k "seed3aa or full_alphabet") > gpurun_out/r2_43_sanitizer_racecheck.log 2>&1
tail -6 gpurun_out/r2_43_sanitizer_racecheck.log
(time compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q -k "full_alphabet") > gpurun_out/r2_43_sanitizer_initcheck.log 2>&1
tail -6 gpurun_out/r2_43_sanitizer_initcheck.log
