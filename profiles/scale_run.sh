#!/bin/bash
# weak-scaling sweep on one box: N = 1, 2, 4, 8 ranks, 16384 instances per GPU
for N in ${SCALE_NS:-1 2 4 8}; do
  if [ $N -eq 1 ]; then
    python bench.py --gpus 1 --steps 600 --warmup 10 --no-cpu 2>/dev/null | tail -1 > gpurun_out/scale_$N.json
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) \
        bench.py --gpus $N --steps 600 --warmup 10 --no-cpu 2>/dev/null | tail -1 > gpurun_out/scale_$N.json
  fi
  python -c "
import json
d=json.load(open('gpurun_out/scale_$N.json'))
print('N=$N value %.2fM e2e %.2fM ms/step %.4f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step']), d['clocks'])"
done
