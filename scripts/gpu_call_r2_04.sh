mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2_04_pytest.log 2>&1
tail -15 gpurun_out/r2_04_pytest.log
