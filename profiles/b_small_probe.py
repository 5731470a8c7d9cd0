"""Small-ensemble hc_step latency: usage b_small_probe.py B [nsteps] [prefill]  (RM3 shape, irregular waves, default options)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hydrochrono_b200 as hc
from hydrochrono_b200 import synth
B = int(sys.argv[1]); n1 = int(sys.argv[2]) if len(sys.argv) > 2 else 400; n0 = int(sys.argv[3]) if len(sys.argv) > 3 else 6010
raw = synth.rm3_like(); T = hc.Tables.from_raw(raw); D = 12; dt = 0.01
amp, om = synth.prescribed_motion(D)
ens = hc.Ensemble(T, batch=B, dt_hint=dt)
ens.set_waves_irregular(dt=dt, duration=(n0 + n1 + 64) * dt, seeds=np.arange(1, B + 1, dtype=np.int32), Hs=2.5, Tp=8.0, gamma=3.3, nfreq=200, ramp=20.0)
out = np.empty((B, D)); t = 0.0; lat = []
for n in range(n0 + n1):
    pose = np.repeat((amp * np.sin(om * t))[None, :], B, 0); vel = np.repeat((amp * om * np.cos(om * t))[None, :], B, 0)
    t0 = time.perf_counter(); ens.step(t, pose, vel, out=out)
    if n >= n0: lat.append(time.perf_counter() - t0)
    t += dt
print("B = %4d: median %.1f us  p90 %.1f" % (B, 1e6 * np.median(lat), 1e6 * np.percentile(lat, 90)), flush=True)
