#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "eta_window or restage" 2>&1 | tail -5
for i in 1 2 3 4 5 6 7 8 9 10 11 12 13 14; do timeout 200 python -m pytest tests/test_gpu_parity.py -q -k "wrong_hint or misprediction or long_run" 2>&1 | grep -E "passed|failed|AssertionError:" | cut -c1-600; done ) > gpurun_out/r02e.log 2>&1
cat gpurun_out/r02e.log
