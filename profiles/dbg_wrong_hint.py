import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import hydrochrono_b200 as hc
from hydrochrono_b200 import synth
from oracle import hc_oracle as orc
import common
raw = synth.rm3_like()
T, O = hc.Tables.from_raw(raw), orc.Tables(raw)
B, D, dt = 4, 12, 0.01
def motion(D, B, t, seed=3):
    amp, om = synth.prescribed_motion(D, seed)
    ph = 0.37 * np.arange(B)[:, None] + 0.11 * np.arange(D)[None, :]
    return amp * np.sin(om * t + ph), amp * om * np.cos(om * t + ph)
ens = hc.Ensemble(T, batch=B, dt_hint=2 * dt, bracket_snap=1e-8, rad_lookahead=int(sys.argv[1]) if len(sys.argv) > 1 else 2)
insts = [orc.Instance(O) for _ in range(B)]
t = 0.0
got, want, ns = [], [], []
for n in range(3300):
    pose, vel = motion(D, B, t)
    F = ens.step(t, pose, vel)
    ref = np.array([i.force(t, pose[b], vel[b]) for b, i in enumerate(insts)])
    got.append(F.copy()); want.append(ref); ns.append(n)
    t += dt
got, want = np.array(got), np.array(want)
tol = common.force_tol(want)
ratio = np.abs(got - want) / tol
bad = np.where(ratio.reshape(len(ns), -1).max(axis=1) > 1)[0]
print("lookahead", ens.lookahead_state(), "stats", ens.rad_block_stats(reset=False), "bad steps", len(bad), bad[:40], bad[-5:] if len(bad) else None)
print("max ratio", ratio.max())
