#!/bin/bash
mkdir -p gpurun_out
for ns in 2 1 2; do echo "== HC_RB_STREAMS=$ns"; HC_RB_STREAMS=$ns timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "not benchmark_state" 2>&1 | grep -E "passed|failed|AssertionError|FAILED" ; done > gpurun_out/r02d.log 2>&1
cat gpurun_out/r02d.log
