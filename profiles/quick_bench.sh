python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for args in "--rad-kernel 1" "--rad-kernel 3" "--rad-kernel 3 --rad-chunk 40" "--rad-kernel 3 --rad-chunk 64"; do
python bench.py --steps 800 --warmup 10 --no-cpu $args 2>gpurun_out/bench_err.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$args', 'value %.2fM e2e %.2fM ms/step %.4f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step']), {k: round(v,4) for k,v in d['kernel_ms'].items()}, 'rad frac %.3f' % d['roofline']['frac'], 'exc frac %.3f' % d['roofline']['excitation']['frac'], d['clocks'], 'faithful:', (d.get('faithful_bracketing') or {}).get('value'))"
done
