#!/bin/bash
# round 2: full GPU suite with the zero-copy host path; 2048 instances on one GPU (the per-GPU share of the north-star split) with and without it; slice capture
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02m_pytest.log
tail -6 gpurun_out/r02m_pytest.log | cut -c1-300
for zc in 1 0; do
  HC_ZERO_COPY=$zc timeout 200 python bench.py --batch 2048 --steps 480 --warmup 10 --no-cpu --no-b1 --no-parity --no-faithful-leg > gpurun_out/r02m_bench_b2048_zc$zc.json 2> gpurun_out/r02m_bench_b2048_zc$zc.err
done
timeout 200 python bench.py --steps 480 --warmup 10 --no-cpu --no-b1 --no-faithful-leg > gpurun_out/r02m_bench_k480.json 2> gpurun_out/r02m_bench_k480.err
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02m_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.3e ms %.4f e2e %.3e e2e_ms %.4f enq %s parity %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['run']['enqueue_ms_per_step'], (d.get('parity') or {}).get('worst_rel')))
    except Exception as ex:
        print(f, 'ERR', ex)
P
ncu --set full --clock-control none --import-source on -k regex:k_rad_block -s 3000 -c 2 \
    -o gpurun_out/prof_radslice_r02a -f python bench.py --steps 60 --warmup 3 --no-cpu --no-b1 --no-parity --no-faithful-leg \
    > gpurun_out/ncu_radslice_r02a.log 2>&1
ls -la gpurun_out/prof_radslice_r02a.ncu-rep
