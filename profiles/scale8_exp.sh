run() {
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 \
    bench.py --gpus 8 --steps 600 --warmup 5 --no-cpu 2>/dev/null | tail -1 > gpurun_out/scale8_$2.json
python -c "
import json
d=json.load(open('gpurun_out/scale8_$2.json'))
print('$2 value %.2fM e2e %.2fM ms/step %.4f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step']), d['clocks'].get('samples'), d['config']['timing'][60:150])"
}
HC_BENCH_NO_CLOCKS=1 run 29701 noclocks
HC_BENCH_CLOCK_PERIOD=0.02 run 29702 nvml20ms
