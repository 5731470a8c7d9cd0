#!/bin/bash
# round 2: k_step_lat (small-ensemble step kernel) -- GPU parity suite, then A/B at 1024 / 2048 / 4096 instances per GPU
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r02s_pytest.log
for B in 2048 1024 4096; do
  for L in 0 1; do
    HC_KSTEP_LAT=$L timeout 300 python bench.py --batch $B --steps 480 --warmup 10 --no-cpu --no-b1 --no-faithful-leg > gpurun_out/r02s_b${B}_lat${L}.json 2> gpurun_out/r02s_b${B}_lat${L}.err
  done
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02s_b*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.3e ms %.4f e2e %.3e e2e_ms %.4f parity %s kernel_ms %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], (d.get('parity') or {}).get('worst_rel'), d.get('kernel_ms')))
    except Exception as ex:
        print(f, 'ERR', ex)
P
