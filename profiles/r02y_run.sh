#!/bin/bash
# round 2: compute-sanitizer over the compact-graph kernels (k_radiation/k_excitation INLINE, k_finalize_warp), the
# imported-series path and the bracket search: memcheck on a group of small tests, racecheck on the B = 1 step
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -x -k "imported or bracket_search or sphere_decay or regular_waves_two or time_cache or irregular_dt" > gpurun_out/r02y_memcheck_tests.log 2>&1; echo "memcheck tests rc $?"
tail -4 gpurun_out/r02y_memcheck_tests.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python profiles/b_small_probe.py 3 20 200 > gpurun_out/r02y_racecheck_b3.log 2>&1; echo "racecheck B=3 rc $?"
tail -4 gpurun_out/r02y_racecheck_b3.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python profiles/b_small_probe.py 70 20 200 > gpurun_out/r02y_memcheck_b70.log 2>&1; echo "memcheck B=70 rc $?"
tail -4 gpurun_out/r02y_memcheck_b70.log
# the racecheck warnings in detail, and with the graph's excitation branch serialised
timeout 600 compute-sanitizer --tool racecheck --racecheck-report hazard --print-limit 12 python profiles/b_small_probe.py 1 2 10 > gpurun_out/r02y_racecheck_detail.log 2>&1
HC_COMPACT_FORK=0 timeout 900 compute-sanitizer --tool racecheck python profiles/b_small_probe.py 3 20 200 > gpurun_out/r02y_racecheck_b3_fork0.log 2>&1; tail -1 gpurun_out/r02y_racecheck_b3_fork0.log
