#!/bin/bash
# Profiling recipe (B200_PROFILING.md) for the per-step kernels of bench.py.  Run under gpurun:
#   gpurun --timeout 1700 -- 'bash profiles/run_ncu.sh r01c'
# Outputs land in gpurun_out/ (scratch); summaries are produced here with profiles/summarize.py and committed.
TAG=${1:-r01}
mkdir -p gpurun_out
# (1) every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes).
#     per step: k_prestep(plan), k_radiation, k_prestep(append), k_finalize; every 8th step additionally
#     k_la_brackets, k_la_taps, k_exc_block.  Skip the eta kernel + the 6010 prefill steps.
SKIP=$(( 1 + 6016 * 4 + 752 * 3 ))
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c 280 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 80 --warmup 3 --no-cpu --no-graph \
    > gpurun_out/ncu_bench_$TAG.log 2>&1
# (2) full-set captures (-s counts launches of the filtered kernel)
ncu --set full --clock-control none --import-source on -k regex:k_radiation -s 6020 -c 2 \
    -o gpurun_out/prof_rad_$TAG -f python bench.py --steps 8 --warmup 3 --no-cpu --no-graph \
    > gpurun_out/ncu_rad_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_exc_block -s 753 -c 2 \
    -o gpurun_out/prof_excblock_$TAG -f python bench.py --steps 24 --warmup 3 --no-cpu --no-graph \
    > gpurun_out/ncu_excblock_$TAG.log 2>&1
ls -la gpurun_out/
