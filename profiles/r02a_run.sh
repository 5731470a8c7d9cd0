#!/bin/bash
# round 2, first GPU pass: GPU tests, the driver's bench line, pass-pacing modes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02a_smi.txt
nproc >> gpurun_out/r02a_smi.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench_driver.json 2> gpurun_out/r02a_bench_driver.err
for m in 1 2 3; do
  timeout 200 python bench.py --steps 480 --warmup 10 --rad-pass-mode $m --no-cpu --no-b1 --no-parity --no-faithful-leg > gpurun_out/r02a_bench_mode$m.json 2> gpurun_out/r02a_bench_mode$m.err
done
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02a_bench_ref.json 2> gpurun_out/r02a_bench_ref.err
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02a_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.3e ms %.4f e2e %.3e enq %s parity %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('run') or {}).get('enqueue_ms_per_step'), (d.get('parity') or {}).get('worst_rel')))
    except Exception as ex:
        print(f, 'ERR', ex)
P
