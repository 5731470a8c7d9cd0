#!/bin/bash
mkdir -p gpurun_out
for M in 2 3; do
timeout 900 python bench.py --workload sphere_irregular_ensemble --steps 480 --warmup 10 --no-cpu --no-b1 --no-faithful-leg --rad-pass-mode $M > gpurun_out/r02x_bench_sphere_pm$M.json 2> gpurun_out/r02x_bench_sphere_pm$M.err
python - <<P
import json
d=json.loads(open('gpurun_out/r02x_bench_sphere_pm$M.json').read().strip().splitlines()[-1])
print('mode $M value %.3e ms %.4f e2e %.3e parity %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('parity') or {}).get('worst_rel')), d['kernel_ms'])
P
done
