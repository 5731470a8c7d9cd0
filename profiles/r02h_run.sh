#!/bin/bash
# round 2: two ranks under torchrun (weak + strong legs, parity on every rank), then the single-process multi-device handle on the same two GPUs
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02h_gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02h_bench_n2.json 2> gpurun_out/r02h_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 20 --warmup 5 > gpurun_out/r02h_ref_n2.json 2> gpurun_out/r02h_ref_n2.err
timeout 600 python bench.py --gpus 2 --steps 96 --warmup 5 > gpurun_out/r02h_bench_multi2.json 2> gpurun_out/r02h_bench_multi2.err
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "multi_device" > gpurun_out/r02h_pytest_multi.log 2>&1; tail -2 gpurun_out/r02h_pytest_multi.log
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02h_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        s=d.get('strong') or {}
        print(f, 'value %.3e ms %.4f e2e %.3e parity %s | strong value %s e2e %s enq %s parity %s cores %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('parity') or {}).get('worst_rel'), s.get('value'), (s.get('e2e') or {}).get('value'), s.get('enqueue_ms_per_step'), s.get('parity_worst_rel'), (d.get('cpu_baseline') or {}).get('cores')))
    except Exception as ex:
        print(f, 'ERR', ex)
P
tail -3 gpurun_out/r02h_bench_n2.err gpurun_out/r02h_bench_multi2.err
