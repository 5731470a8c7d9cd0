python -m pytest tests -m gpu -x -q -k "baseline_config or radiation_lookahead or long_run" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
