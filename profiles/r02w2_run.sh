#!/bin/bash
# (superseded by r02w5_run.sh: HC_FINALIZE_WARP_ITEMS no longer exists) per-step hc_step latency of mid-size ensembles, k_finalize (thread per item) vs k_finalize_warp
mkdir -p gpurun_out
for B in 512 1024 2048 4096; do for W in 4096 1000000; do echo -n "warp_items=$W "; HC_FINALIZE_WARP_ITEMS=$W python profiles/b_small_probe.py $B 300 6010; done; done 2>&1 | tee gpurun_out/r02w2_finalize_mid_b.txt
