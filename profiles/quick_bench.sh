python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --no-cpu 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_final.json
python -c "
import json
d=json.load(open('gpurun_out/bench_final.json'))
print('value %.2fM e2e %.2fM ms/step %.4f launches %d' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], d['gpu_launches']), d['clocks']['sm_mhz'], d['roofline']['frac'], d['roofline']['traffic'])"
