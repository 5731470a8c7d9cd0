#!/bin/bash
# round 2: split excitation builds -- full GPU suite, the driver's window, a 480-step window, the sphere workload
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02p_pytest.log
tail -5 gpurun_out/r02p_pytest.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02p_bench_driver.json 2> gpurun_out/r02p_bench_driver.err
timeout 200 python bench.py --steps 480 --warmup 10 --no-cpu --no-b1 --no-faithful-leg > gpurun_out/r02p_bench_k480.json 2> gpurun_out/r02p_bench_k480.err
timeout 300 python bench.py --steps 100 --warmup 5 --workload sphere_irregular_ensemble --no-cpu --no-b1 > gpurun_out/r02p_bench_sphere.json 2> gpurun_out/r02p_bench_sphere.err
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02p_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.3e ms %.4f e2e %.3e e2e_ms %.4f enq %s parity %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['run']['enqueue_ms_per_step'], (d.get('parity') or {}).get('worst_rel')))
    except Exception as ex:
        print(f, 'ERR', ex)
P
tail -n 3 gpurun_out/r02p_bench_sphere.err
