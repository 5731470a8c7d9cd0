#!/bin/bash
# Profiling recipe (B200_PROFILING.md) for the kernels of bench.py.  Run under gpurun:
#   gpurun --timeout 1700 -- 'bash profiles/run_ncu.sh r02a'
# Outputs land in gpurun_out/ (scratch); summaries are produced here with profiles/summarize.py and committed.
TAG=${1:-r02}
mkdir -p gpurun_out
ARGS="--steps 120 --warmup 3 --no-cpu --no-b1 --no-parity --no-faithful-leg"
# (1) every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes).
#     Default mode per step: k_step<12> and one slice of k_rad_block<12> (the next block's pass, side stream);
#     every 8th step additionally k_la_brackets, k_la_taps, k_exc_block_mma<12> (side stream).
#     The number of launches before the timed region comes from an un-profiled run of the same command.
SKIP=$(python bench.py $ARGS 2>/dev/null | tail -1 | python -c "import json,sys; print(json.loads(sys.stdin.read())['run']['launches_before_timed_region'] + 8)")
echo "launch skip = $SKIP"
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c 300 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py $ARGS > gpurun_out/ncu_bench_$TAG.log 2>&1
# (2) full-set captures (-s counts launches of the filtered kernel)
#     (a) ONE slice of the pass as the timed region launches it (side stream, one wave of 444 CTAs): 6010 prefill
#         steps = 125 blocks x ~62 slice launches; skip well into steady state
ncu --set full --clock-control none --import-source on -k regex:k_rad_block -s 7800 -c 2 \
    -o gpurun_out/prof_radslice_$TAG -f python bench.py --steps 60 --warmup 3 --no-cpu --no-b1 --no-parity --no-faithful-leg \
    > gpurun_out/ncu_radslice_$TAG.log 2>&1
#     (b) the whole pass in one launch (--rad-lookahead 3): 6010 prefill steps = 126 passes
ncu --set full --clock-control none --import-source on -k regex:k_rad_block -s 127 -c 1 \
    -o gpurun_out/prof_radblock_$TAG -f python bench.py --steps 60 --warmup 3 --no-cpu --no-b1 --no-parity --no-faithful-leg --rad-lookahead 3 \
    > gpurun_out/ncu_radblock_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_exc_block_mma -s 754 -c 1 \
    -o gpurun_out/prof_excblock_$TAG -f python bench.py --steps 24 --warmup 3 --no-cpu --no-b1 --no-parity --no-faithful-leg \
    > gpurun_out/ncu_excblock_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step -s 6020 -c 1 \
    -o gpurun_out/prof_step_$TAG -f python bench.py --steps 8 --warmup 3 --no-cpu --no-b1 --no-parity --no-faithful-leg \
    > gpurun_out/ncu_step_$TAG.log 2>&1
ls -la gpurun_out/ | grep $TAG
