#!/bin/bash
# round 2, last build: 8 GPUs -- torchrun ranks (the driver's launch; weak + north-star strong split in one line), the reference arm
# under torchrun, the single-process multi-device handle (hc_multi_*), the multi-device parity test
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02z_gpus.txt; nproc >> gpurun_out/r02z_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02z_bench_n8_k20.json 2> gpurun_out/r02z_bench_n8_k20.err
timeout 300 $TR --master-port 29523 bench.py --gpus 8 --impl reference --steps 20 --warmup 5 > gpurun_out/r02z_ref_n8.json 2> gpurun_out/r02z_ref_n8.err
timeout 400 python bench.py --gpus 8 --steps 96 --warmup 5 > gpurun_out/r02z_bench_multi8.json 2> gpurun_out/r02z_bench_multi8.err
timeout 200 python -m pytest tests/test_gpu_parity.py -q -k "multi_device" > gpurun_out/r02z_pytest_multi.log 2>&1; tail -2 gpurun_out/r02z_pytest_multi.log
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02z_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        s=d.get('strong') or {}
        print(f, 'value %.3e ms %.4f e2e %.3e enq %s parity %s | strong value %s ms %s e2e %s enq %s parity %s cores %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('run') or {}).get('enqueue_ms_per_step'), (d.get('parity') or {}).get('worst_rel'), s.get('value'), s.get('ms_per_step'), (s.get('e2e') or {}).get('value'), s.get('enqueue_ms_per_step'), s.get('parity_worst_rel'), (d.get('cpu_baseline') or {}).get('cores')))
    except Exception as ex:
        print(f, 'ERR', ex)
P
for f in gpurun_out/r02z_bench_n8_k20.err gpurun_out/r02z_bench_multi8.err; do echo "== $f"; tail -n 3 $f; done
