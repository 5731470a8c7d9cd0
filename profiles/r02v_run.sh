#!/bin/bash
# round 2: GPU suite after extending the warp finalize to every compact-path size; latency at B = 1 / 16 / 80 / 256
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu > gpurun_out/r02v_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r02v_pytest.log
python - <<'P' > gpurun_out/r02v_latency_vs_b.txt 2>&1
import sys, time, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import hydrochrono_b200 as hc
from hydrochrono_b200 import synth
raw = synth.rm3_like(); T = hc.Tables.from_raw(raw); D = 12; dt = 0.01
amp, om = synth.prescribed_motion(D)
for B in (1, 16, 80, 256, 340):
    ens = hc.Ensemble(T, batch=B, dt_hint=dt)
    n0, n1 = 6010, 400
    ens.set_waves_irregular(dt=dt, duration=(n0 + n1 + 64) * dt, seeds=np.arange(1, B + 1, dtype=np.int32), Hs=2.5, Tp=8.0, gamma=3.3, nfreq=200, ramp=20.0)
    out = np.empty((B, D)); t = 0.0; lat = []
    for n in range(n0 + n1):
        pose = np.repeat((amp * np.sin(om * t))[None, :], B, 0); vel = np.repeat((amp * om * np.cos(om * t))[None, :], B, 0)
        t0 = time.perf_counter(); ens.step(t, pose, vel, out=out)
        if n >= n0: lat.append(time.perf_counter() - t0)
        t += dt
    print("B = %4d: median %.1f us  p90 %.1f" % (B, 1e6 * np.median(lat), 1e6 * np.percentile(lat, 90)), flush=True)
    ens.close()
P
cat gpurun_out/r02v_latency_vs_b.txt
