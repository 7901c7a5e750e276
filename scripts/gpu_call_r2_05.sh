mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q -k "verbose or golden or pipeline or edge") > gpurun_out/r2_05_pytest.log 2>&1
tail -8 gpurun_out/r2_05_pytest.log
