python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
for args in "" "--rad-lookahead 1" "--workload sphere_irregular_ensemble"; do
python bench.py --steps 960 --warmup 10 --no-cpu $args 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_last.json
python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_last.json').read())
r=d['roofline']
print('$args', 'value %.2fM e2e %.2fM ms/step %.4f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step']), {k: round(v,4) for k,v in d['kernel_ms'].items()}, r['kernel'][:24], r['bound'], 'ach %.1f peak %.1f frac %.3f' % (r['achieved'], r['peak'], r['frac']), 'launch_ms', r.get('launch_ms'), 'exc frac %.3f' % r['excitation']['frac'], d['clocks'], 'faithful:', (d.get('faithful_bracketing') or {}).get('value'), 'launches', d['gpu_launches'], d['step_roofline'])"
tail -3 gpurun_out/bench_err.log
done
