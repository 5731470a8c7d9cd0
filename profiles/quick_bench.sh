python bench.py --no-cpu --steps 300 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_final2.json
python -c "
import json
d=json.load(open('gpurun_out/bench_final2.json'))
print('value %.2fM e2e %.2fM' % (d['value']/1e6, d['e2e']['value']/1e6), d['config']['l2'])"
tail -2 gpurun_out/bench_err.log
