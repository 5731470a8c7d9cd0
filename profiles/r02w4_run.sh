#!/bin/bash
# round 2: the one-graph (compact) host path extended beyond 64 KB of inputs vs the two-phase path, 512 - 2048 instances
mkdir -p gpurun_out
for B in 512 1024 2048; do
  echo -n "two-phase          "; python profiles/b_small_probe.py $B 300 6010
  echo -n "compact direct     "; HC_COMPACT_MAX_BYTES=1000000 python profiles/b_small_probe.py $B 300 6010
  echo -n "compact copy node  "; HC_COMPACT_MAX_BYTES=1000000 HC_COMPACT_DIRECT=0 python profiles/b_small_probe.py $B 300 6010
done 2>&1 | tee gpurun_out/r02w4_compact_mid_b.txt
