#!/bin/bash
# round 2: pass pacing modes with the excitation stream ranked above an ungated pass; then the ncu recipe
mkdir -p gpurun_out
for m in 1 3 2; do
  timeout 200 python bench.py --steps 480 --warmup 10 --rad-pass-mode $m --no-cpu --no-b1 --no-parity --no-faithful-leg > gpurun_out/r02j_bench_mode$m.json 2> gpurun_out/r02j_bench_mode$m.err
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02j_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.3e ms %.4f e2e %.3e enq %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['run']['enqueue_ms_per_step']))
    except Exception as ex:
        print(f, 'ERR', ex)
P
bash profiles/run_ncu.sh r02a
