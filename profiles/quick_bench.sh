python -m pytest tests -m gpu -x -q -k "radiation_lookahead or baseline_config" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
