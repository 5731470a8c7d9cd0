#!/bin/bash
# round 2: k_rad_block<6> with row pairs (3 exact k-steps instead of 2 x 2 padded) -- parity tests that run the D = 6
# block path, then the sphere workload
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "radiation_lookahead or baseline_config or large_ensemble or benchmark_state or reset_and_wrong or mixing" > gpurun_out/r02x3_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r02x3_pytest.log
timeout 900 python bench.py --workload sphere_irregular_ensemble --steps 480 --warmup 10 --no-cpu --no-b1 > gpurun_out/r02x3_bench_sphere.json 2> gpurun_out/r02x3_bench_sphere.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02x3_bench_sphere.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value %.3e ms %.4f e2e %.3e parity %s faithful %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('parity') or {}).get('worst_rel'), (d.get('faithful_bracketing') or {}).get('value')))
print('rad block frac', r['frac'], 'launch_ms', r['launch_ms'], d['kernel_ms'], 'step fp64', d['step_roofline']['fp64_frac'])
P
