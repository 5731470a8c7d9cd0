#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "zero_copy or benchmark_state or multi_device or large_ensemble" > gpurun_out/r02o_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02o_pytest.log
tail -4 gpurun_out/r02o_pytest.log | cut -c1-300
for zc in 1 0; do
  HC_ZERO_COPY=$zc timeout 200 python bench.py --batch 2048 --steps 480 --warmup 10 --no-cpu --no-b1 --no-parity --no-faithful-leg > gpurun_out/r02o_bench_b2048_zc$zc.json 2> gpurun_out/r02o_bench_b2048_zc$zc.err
done
timeout 200 python bench.py --steps 480 --warmup 10 --no-cpu --no-b1 --no-faithful-leg > gpurun_out/r02o_bench_k480.json 2> gpurun_out/r02o_bench_k480.err
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02o_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.3e ms %.4f e2e %.3e e2e_ms %.4f enq %s parity %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['run']['enqueue_ms_per_step'], (d.get('parity') or {}).get('worst_rel')))
    except Exception as ex:
        print(f, 'ERR', ex)
P
