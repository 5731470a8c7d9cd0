#!/bin/bash
# Profiling recipe (B200_PROFILING.md) for the kernels of bench.py.  Run under gpurun:
#   gpurun --timeout 1700 -- 'bash profiles/run_ncu.sh r01d'
# Outputs land in gpurun_out/ (scratch); summaries are produced here with profiles/summarize.py and committed.
TAG=${1:-r01}
mkdir -p gpurun_out
# (1) every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes).
#     Default mode per step: k_rad_step, k_finalize, one slice of k_rad_block12 (the next block's pass);
#     every 8th step additionally k_la_brackets, k_la_taps, k_exc_block_mma.  Skip the eta kernel + the prefill.
SKIP=$(( 1 + 5 + 6016 * 3 + 753 * 3 + 1 ))
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c 330 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 120 --warmup 3 --no-cpu --no-graph \
    > gpurun_out/ncu_bench_$TAG.log 2>&1
# (2) full-set captures (-s counts launches of the filtered kernel)
#     whole radiation pass in one launch (--rad-lookahead 3): 6010 prefill steps = 126 passes
ncu --set full --clock-control none --import-source on -k regex:k_rad_block -s 127 -c 1 \
    -o gpurun_out/prof_radblock_$TAG -f python bench.py --steps 60 --warmup 3 --no-cpu --no-graph --rad-lookahead 3 \
    > gpurun_out/ncu_radblock_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_exc_block_mma -s 754 -c 1 \
    -o gpurun_out/prof_excblock_$TAG -f python bench.py --steps 24 --warmup 3 --no-cpu --no-graph \
    > gpurun_out/ncu_excblock_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step -s 6020 -c 1 \
    -o gpurun_out/prof_radstep_$TAG -f python bench.py --steps 8 --warmup 3 --no-cpu --no-graph \
    > gpurun_out/ncu_radstep_$TAG.log 2>&1
ls -la gpurun_out/
