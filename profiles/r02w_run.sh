#!/bin/bash
# round 2: hc_step latency of small ensembles (RM3 shape, irregular waves, full window, default options) vs ensemble size
mkdir -p gpurun_out
for B in 1 16 80 256; do python profiles/b_small_probe.py $B 1000 6010; done 2>&1 | tee gpurun_out/r02w_latency_vs_b.txt
