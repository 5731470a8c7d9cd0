"""B = 1 drop-in step: per-call wall time and (under ncu) the per-kernel device times.  usage: b1_probe.py rm3|sphere [nsteps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import hydrochrono_b200 as hc
from hydrochrono_b200 import synth
import common
which = sys.argv[1] if len(sys.argv) > 1 else "rm3"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
if which == "sphere":
    raw, dt, D, prefill = common.sphere_raw(), 0.015, 6, 1010
    sea = dict(Hs=2.0, Tp=12.0, gamma=1.0, nfreq=1000, ramp=60.0)
else:
    raw, dt, D, prefill = synth.rm3_like(), 0.01, 12, 6010
    sea = dict(Hs=2.5, Tp=8.0, gamma=3.3, nfreq=1000, ramp=20.0)
T = hc.Tables.from_raw(raw)
ens = hc.Ensemble(T, batch=1, dt_hint=dt)
ens.set_waves_irregular(dt=dt, duration=(prefill + nsteps + 64) * dt, seeds=np.array([1], dtype=np.int32), **sea)
amp, om = synth.prescribed_motion(D)
t, out = 0.0, np.empty((1, D))
lat = []
for n in range(prefill + nsteps):
    pose = (amp * np.sin(om * t))[None, :].copy(); vel = (amp * om * np.cos(om * t))[None, :].copy()
    t0 = time.perf_counter()
    ens.step(t, pose, vel, out=out)
    if n >= prefill:
        lat.append(time.perf_counter() - t0)
    t += dt
print("%s: median %.1f us  p10 %.1f  p90 %.1f" % (which, 1e6 * np.median(lat), 1e6 * np.percentile(lat, 10), 1e6 * np.percentile(lat, 90)))
