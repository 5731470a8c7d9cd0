#!/bin/bash
# round 2: compute-sanitizer (memcheck, then racecheck on shared memory) over the smoke run and two small parity tests
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py smoke > gpurun_out/r02r_memcheck_smoke.log 2>&1; echo "memcheck smoke rc $?"
tail -4 gpurun_out/r02r_memcheck_smoke.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -x -k "multi_device or pinned_and_pageable or eta_window or sphere_decay" > gpurun_out/r02r_memcheck_tests.log 2>&1; echo "memcheck tests rc $?"
tail -4 gpurun_out/r02r_memcheck_tests.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python __graft_entry__.py smoke > gpurun_out/r02r_racecheck_smoke.log 2>&1; echo "racecheck smoke rc $?"
tail -4 gpurun_out/r02r_racecheck_smoke.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-b1 --no-faithful-leg > gpurun_out/r02r_bench_driver.json 2> gpurun_out/r02r_bench_driver.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02r_bench_driver.json').read().strip().splitlines()[-1])
print('value %.3e e2e %.3e parity %s clocks %s' % (d['value'], d['e2e']['value'], d['parity']['worst_rel'], d['clocks']))
P
