mkdir -p gpurun_out
python scripts/cli_timing_probe.py 5000 > gpurun_out/r2_14_cli_5k.log 2>&1
python scripts/cli_timing_probe.py 100000 > gpurun_out/r2_14_cli_100k.log 2>&1
cat gpurun_out/r2_14_cli_5k.log | grep -v "^ *$" | head -60
