#!/bin/bash
mkdir -p gpurun_out
HC_KSTEP2=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "benchmark_state or large_ensemble or rm3_radiation_lookahead or misprediction or long_run or pinned" > gpurun_out/r02u_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02u_pytest.log
tail -4 gpurun_out/r02u_pytest.log | cut -c1-300
for v in 1 0 1 0; do
  HC_KSTEP2=$v timeout 200 python bench.py --steps 480 --warmup 10 --no-cpu --no-b1 --no-faithful-leg > gpurun_out/r02u_bench_k480_v$v.json 2> gpurun_out/r02u_bench_v$v.err
  python - <<P
import json
d=json.loads(open('gpurun_out/r02u_bench_k480_v$v.json').read().strip().splitlines()[-1])
print('kstep2=$v value %.3e ms %.4f e2e %.3e e2e_ms %.4f parity %s kms %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], (d.get('parity') or {}).get('worst_rel'), d['kernel_ms']['radiation']))
P
done
