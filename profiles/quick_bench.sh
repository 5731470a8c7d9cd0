python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python bench.py --steps 1000 --warmup 10 --no-cpu 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_last.json
python -c "
import json
d=json.load(open('gpurun_out/bench_last.json'))
print('value %.2fM e2e %.2fM' % (d['value']/1e6, d['e2e']['value']/1e6), d['config']['timing'][:160])"
tail -3 gpurun_out/bench_err.log
