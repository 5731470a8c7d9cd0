python -m pytest tests -m gpu -x -q -k "rm3_irregular or large_ensemble" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
run() {
python bench.py --steps 960 --warmup 10 --no-cpu $ARGS 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_last.json
python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_last.json').read())
r=d['roofline']
print('$TAG $ARGS', 'value %.2fM e2e %.2fM ms/step %.4f e2e ms %.4f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], d['e2e']['ms_per_step']), {k: round(v,4) for k,v in d['kernel_ms'].items()}, r['kernel'][:20], 'launch_ms', r.get('launch_ms'), 'frac %.3f' % r['frac'], 'exc frac %.3f' % r['excitation']['frac'])"
tail -2 gpurun_out/bench_err.log
}
ARGS="" TAG="default" run
true
