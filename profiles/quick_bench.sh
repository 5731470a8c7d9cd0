python -m pytest tests -m gpu -x -q -k "reset_and_wrong or large_ensemble or smoke" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
