#!/bin/bash
mkdir -p gpurun_out
for B in 16 40 80 256; do for Dm in 1 0; do echo -n "direct=$Dm "; HC_COMPACT_DIRECT=$Dm python profiles/b_small_probe.py $B 1000 1500; done; done 2>&1 | tee gpurun_out/r02w_direct_vs_b.txt
