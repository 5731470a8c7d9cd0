#!/bin/bash
# round 2, last build: memcheck over the kernels added after r02y -- k_finalize_split (B = 70 one-graph step) and the
# row-pair k_rad_block<6> (single-body block path, lag spacing = dt and 2 dt)
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python profiles/b_small_probe.py 70 20 200 > gpurun_out/r02y3_memcheck_b70.log 2>&1; echo "memcheck B=70 rc $?"; tail -2 gpurun_out/r02y3_memcheck_b70.log
timeout 280 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -x -k "radiation_lookahead and 2-1]" > gpurun_out/r02y3_memcheck_d6.log 2>&1; echo "memcheck D=6 block rc $?"; tail -3 gpurun_out/r02y3_memcheck_d6.log
