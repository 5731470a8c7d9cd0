#!/bin/bash
# round 2: full GPU suite after the compact host path, bench with B=1 latency
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02i_pytest.log
tail -6 gpurun_out/r02i_pytest.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02i_bench_driver.json 2> gpurun_out/r02i_bench_driver.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02i_bench_driver.json').read().strip().splitlines()[-1])
print('value %.3e e2e %.3e parity %s' % (d['value'], d['e2e']['value'], d['parity']['worst_rel']))
print(json.dumps(d['latency_b1_us'], indent=1))
P
