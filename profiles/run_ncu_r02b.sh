#!/bin/bash
# round 2, last build: launch list of the timed region + one full capture of the phase-2 kernel that the last build
# launches at 16384 instances (k_step2<12>); k_rad_block<12> / k_exc_block_mma<12> are unchanged since the r02a captures.
TAG=r02b
mkdir -p gpurun_out
ARGS="--steps 120 --warmup 3 --no-cpu --no-b1 --no-parity --no-faithful-leg"
SKIP=$(python bench.py $ARGS 2>/dev/null | tail -1 | python -c "import json,sys; print(json.loads(sys.stdin.read())['run']['launches_before_timed_region'] + 8)")
echo "launch skip = $SKIP"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c 300 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py $ARGS > gpurun_out/ncu_bench_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 6020 -c 1 \
    -o gpurun_out/prof_step_$TAG -f python bench.py --steps 8 --warmup 3 --no-cpu --no-b1 --no-parity --no-faithful-leg \
    > gpurun_out/ncu_step_$TAG.log 2>&1
ls -la gpurun_out/ | grep $TAG
