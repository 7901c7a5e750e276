mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q) > gpurun_out/r2_15_pytest.log 2>&1; tail -5 gpurun_out/r2_15_pytest.log
