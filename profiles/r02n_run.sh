#!/bin/bash
# round 2: launch list at 2048 instances per GPU (the per-GPU share of the north-star split)
mkdir -p gpurun_out
ARGS="--batch 2048 --steps 120 --warmup 3 --no-cpu --no-b1 --no-parity --no-faithful-leg"
SKIP=$(python bench.py $ARGS 2>/dev/null | tail -1 | python -c "import json,sys; print(json.loads(sys.stdin.read())['run']['launches_before_timed_region'] + 8)")
echo "launch skip = $SKIP"
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c 300 --csv \
    --log-file gpurun_out/launches_r02a_b2048.csv python bench.py $ARGS > gpurun_out/ncu_bench_r02a_b2048.log 2>&1
HC_TRACE=1 timeout 100 python bench.py --batch 2048 --steps 480 --warmup 10 --no-cpu --no-b1 --no-parity --no-faithful-leg 2>&1 | grep "hc trace"
