"""Host integrator stand-in for Chrono's default (linearised Euler) time stepper (TEST INFRASTRUCTURE).

SURVEY.md A.10: the reference evaluates the hydro force once per distinct time value with the state
at that moment, then Chrono advances  v_{n+1} = v_n + dt (M + A_inf)^-1 F_n ,  x_{n+1} = x_n + dt v_{n+1}.
Rotations stay small in every fixture used here, so Cardan angles are integrated like positions.
"""
import numpy as np


def run(force_fn, added_mass, masses, inertias, pose0, dt, nsteps, gvec=(0.0, 0.0, -9.81), free=None,
        damping=None, vel0=None, record=None):
    """force_fn(t, pose, vel) -> hydro force (D,).  masses [N], inertias [N][3] (diag).  free: bool mask (D,)
    of unconstrained DoFs (None = all).  damping: (D,) explicit linear damper coefficients (TSDA).
    Returns (times, poses) recorded after each step, like the reference mains do."""
    pose = np.array(pose0, dtype=np.float64)
    D = pose.size
    N = D // 6
    vel = np.zeros(D) if vel0 is None else np.array(vel0, dtype=np.float64)
    g = np.asarray(gvec, dtype=np.float64)
    Mdiag = np.zeros(D)
    fg = np.zeros(D)
    for b in range(N):
        Mdiag[6 * b:6 * b + 3] = masses[b]
        Mdiag[6 * b + 3:6 * b + 6] = inertias[b]
        fg[6 * b:6 * b + 3] = masses[b] * g
    M = np.diag(Mdiag) + added_mass
    free = np.ones(D, bool) if free is None else np.asarray(free, bool)
    idx = np.where(free)[0]
    Minv = np.linalg.inv(M[np.ix_(idx, idx)])
    damping = np.zeros(D) if damping is None else np.asarray(damping, dtype=np.float64)
    times = np.empty(nsteps)
    poses = np.empty((nsteps, D))
    t = 0.0
    for n in range(nsteps):
        F = force_fn(t, pose, vel) + fg - damping * vel
        if record is not None:
            record(n, t, pose, vel, F)
        acc = np.zeros(D)
        acc[idx] = Minv @ F[idx]
        vel = vel + dt * acc
        vel[~free] = 0.0
        pose = pose + dt * vel
        t = t + dt
        times[n] = t
        poses[n] = pose
    return times, poses
