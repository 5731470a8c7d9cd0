#!/bin/bash
# round 2: per-step (bit-faithful) hc_step latency of mid-size ensembles with the excitation and radiation convolutions
# as parallel branches of the phase-1 graph (HC_COMPACT_FORK=1) or in sequence (0)
mkdir -p gpurun_out
for B in 512 1024 2048 4096 8192; do for F in 0 1; do echo -n "fork=$F "; HC_FORK_MAX_BATCH=100000 HC_COMPACT_FORK=$F python profiles/b_small_probe.py $B 300 6010; done; done 2>&1 | tee gpurun_out/r02w3_fork_mid_b.txt
