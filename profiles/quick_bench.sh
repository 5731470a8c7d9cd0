python -m pytest tests -m gpu -x -q -k "tapered or host_layer or api" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
