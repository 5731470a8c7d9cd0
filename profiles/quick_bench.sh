nproc
for i in 1 2 3 4 5 6 7; do timeout 25 python -c "
while True: pass" & done
sleep 1
python bench.py --steps 1000 --warmup 10 --no-cpu 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_last.json
python -c "
import json
d=json.load(open('gpurun_out/bench_last.json'))
print('with 7 burners: value %.2fM e2e %.2fM' % (d['value']/1e6, d['e2e']['value']/1e6), d['config']['timing'][:160])"
wait
