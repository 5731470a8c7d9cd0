#!/bin/bash
# round 2, second GPU pass: GPU tests, the driver's bench line, one vs two alternating pass streams
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02b_pytest.log
tail -5 gpurun_out/r02b_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench_driver.json 2> gpurun_out/r02b_bench_driver.err
for ns in 1 2; do
  HC_RB_STREAMS=$ns timeout 200 python bench.py --steps 480 --warmup 10 --no-cpu --no-b1 --no-parity --no-faithful-leg > gpurun_out/r02b_bench_ns$ns.json 2> gpurun_out/r02b_bench_ns$ns.err
done
HC_RB_STREAMS=2 timeout 200 python bench.py --steps 480 --warmup 10 --rad-pass-mode 2 --no-cpu --no-b1 --no-parity --no-faithful-leg > gpurun_out/r02b_bench_ns2_mode2.json 2> gpurun_out/r02b_bench_ns2_mode2.err
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02b_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.3e ms %.4f e2e %.3e enq %s parity %s pass_ms %.3f frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('run') or {}).get('enqueue_ms_per_step'), (d.get('parity') or {}).get('worst_rel'), d['roofline'].get('launch_ms',0), d['roofline']['frac']))
    except Exception as ex:
        print(f, 'ERR', ex)
P
