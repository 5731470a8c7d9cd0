#!/bin/bash
# round 2: B = 1 drop-in latency measured from C++ (no Python in the loop): hc_step through the C ABI, and the first
# ComponentFunc evaluation at a new ChTime through TestHydro (host layer) -- sphere (real tables) and RM3 shapes
mkdir -p gpurun_out
python - <<'P'
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import common
from hydrochrono_b200 import h5io, synth
h5io.write_bemio('gpurun_out/sphere.h5', common.sphere_raw())
h5io.write_bemio('gpurun_out/rm3.h5', synth.rm3_like())
P
make -C hydrochrono_b200/host -j8 > /dev/null
B=hydrochrono_b200/host/build/bench_b1_latency
{ echo "== sphere (D = 6, L = 1001, Le = 8334, dt = 0.015)"; $B gpurun_out/sphere.h5 1 0.015 1010 3000;
  echo "== RM3 shape (D = 12, L = 1001, Le = 6000, dt = 0.01)"; $B gpurun_out/rm3.h5 2 0.01 6010 3000; } 2>&1 | tee gpurun_out/r02q_b1_cpp.txt
rm -f gpurun_out/sphere.h5 gpurun_out/rm3.h5
