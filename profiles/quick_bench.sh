python -m pytest tests -m gpu -x -q -k "mixing_host_and_device" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
