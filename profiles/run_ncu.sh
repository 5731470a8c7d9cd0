#!/bin/bash
# Profiling recipe (B200_PROFILING.md) for the per-step kernels of bench.py.  Run under gpurun:
#   gpurun --timeout 1500 -- 'bash profiles/run_ncu.sh r01a'
# Outputs land in gpurun_out/ (scratch); summaries are copied into profiles/ by hand.
TAG=${1:-r01}
SKIP=$(( 1 + (6010 + 3 + 3) * 5 ))   # eta kernel + (prefill + warmups) x 5 kernels per step (plan, excitation, append, radiation, finalize)
mkdir -p gpurun_out
# (1) every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c 200 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 40 --warmup 3 --no-cpu --no-graph \
    > gpurun_out/ncu_bench_$TAG.log 2>&1
# (2) full-set capture of the two convolution kernels (2 launches each)
ncu --set full --clock-control none --import-source on -k regex:k_radiation -s 6020 -c 2 \
    -o gpurun_out/prof_rad_$TAG -f python bench.py --steps 4 --warmup 3 --no-cpu --no-graph \
    > gpurun_out/ncu_rad_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_excitation -s 6020 -c 2 \
    -o gpurun_out/prof_exc_$TAG -f python bench.py --steps 4 --warmup 3 --no-cpu --no-graph \
    > gpurun_out/ncu_exc_$TAG.log 2>&1
ls -la gpurun_out/
