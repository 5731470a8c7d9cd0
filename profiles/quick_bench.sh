python -m pytest tests -m gpu -x -q -k "large_ensemble" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
