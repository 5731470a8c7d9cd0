#!/bin/bash
# round 2: finalize variants across ensemble sizes (per-step path, RM3 shape): HC_FINALIZE_MODE 1 = thread per item,
# 2 = warp per item, 3 = split (32 instances x 16 partial groups per CTA)
mkdir -p gpurun_out
run() { echo -n "mode=$2 "; HC_FINALIZE_MODE=$2 python profiles/b_small_probe.py $1 300 6010; }
{ for B in 16 80 256 512; do run $B 2; run $B 3; done
  run 1024 1; run 1024 2; run 1024 3
  for B in 2048 4096; do run $B 1; run $B 3; done; } 2>&1 | tee gpurun_out/r02w5_finalize_modes.txt
