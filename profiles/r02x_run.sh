#!/bin/bash
# round 2: the secondary workload of SURVEY 8(d) -- sphere x 16384 on the reference's real tables (D = 6, L = 1001 at
# lag spacing = dt = 0.015, Le = 8334, nf = 1000)
mkdir -p gpurun_out
timeout 900 python bench.py --workload sphere_irregular_ensemble --steps 480 --warmup 10 > gpurun_out/r02x_bench_sphere.json 2> gpurun_out/r02x_bench_sphere.err
tail -3 gpurun_out/r02x_bench_sphere.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02x_bench_sphere.json').read().strip().splitlines()[-1])
print('value %.3e ms %.4f e2e %.3e parity %s faithful %s cpu %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('parity') or {}), (d.get('faithful_bracketing') or {}).get('value'), (d.get('cpu_baseline') or {})))
print(json.dumps(d['roofline'])[:1500]); print(d['kernel_ms']); print(d.get('step_roofline'))
P
