#!/bin/bash
# round 2: full GPU suite, the driver's bench line, the multi-device handle with two shards on one GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02f_pytest.log
tail -4 gpurun_out/r02f_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench_driver.json 2> gpurun_out/r02f_bench_driver.err
timeout 200 python bench.py --steps 480 --warmup 10 --no-cpu --no-b1 --no-parity --no-faithful-leg > gpurun_out/r02f_bench_k480.json 2> gpurun_out/r02f_bench_k480.err
timeout 300 python bench.py --multi-devices 0,0 --batch 4096 --steps 96 --warmup 5 > gpurun_out/r02f_bench_multi00.json 2> gpurun_out/r02f_bench_multi00.err
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02f_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.3e ms %.4f e2e %.3e enq %s parity %s pass_ms %s strong %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('run') or {}).get('enqueue_ms_per_step'), (d.get('parity') or {}).get('worst_rel'), (d.get('roofline') or {}).get('launch_ms'), (d.get('strong') or {}).get('value')))
    except Exception as ex:
        print(f, 'ERR', ex)
P
tail -3 gpurun_out/r02f_bench_multi00.err
