#!/bin/bash
# round 2: one B200, the driver's command and a 480-step window, after the B = 1 / compact-graph changes
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02u_bench_driver.json 2> gpurun_out/r02u_bench_driver.err
timeout 600 python bench.py --steps 480 --warmup 10 --no-cpu > gpurun_out/r02u_bench_k480.json 2> gpurun_out/r02u_bench_k480.err
python - <<'P'
import json
for f in ('gpurun_out/r02u_bench_driver.json','gpurun_out/r02u_bench_k480.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'value %.3e ms %.4f e2e %.3e parity %s faithful %s b1 %s cpu %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('parity') or {}).get('worst_rel'), (d.get('faithful_bracketing') or {}).get('value'), json.dumps(d.get('latency_b1_us')), (d.get('cpu_baseline') or {}).get('value')))
P
