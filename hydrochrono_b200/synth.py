"""Synthetic, seeded BEMIO-shaped hydro tables (RM3 / OSWEC / DeepCWind-like) -- SURVEY.md section 8(d).

The reference snapshot ships only sphere.h5 (rm3.h5, oswec.h5, deepcwind.h5, f3of.h5 are stripped large
blobs), so the multi-body configurations of BASELINE.json run on tables generated here.  The arrays are
RAW file-convention values (K and A_inf normalised by rho, excitation by rho*g, exactly what a BEMIO .h5
holds), so the product and the oracle apply the reference's scalings themselves.
"""
import numpy as np

RM3_MASSES = (725834.0, 886691.0)                 # demos/rm3/demo_rm3_reg_waves.cpp:97-125
RM3_INERTIAS = ((20907301.0, 21306090.66, 37085481.11), (94419614.57, 94407091.24, 28542224.82))
RM3_CG = ((0.0, 0.0, -0.72), (0.0, 0.0, -21.29))


def make_tables(num_bodies=2, rirf_steps=1001, rirf_duration=60.0, exc_half_window=30.0, exc_irf_steps=1001,
                num_freqs=260, omega_max=5.2, rho=1000.0, g=9.81, water_depth=200.0, seed=20261017,
                masses=None):
    """Returns the raw dict layout of tests/h5lite.load_bemio."""
    rng = np.random.Generator(np.random.MT19937(seed))
    N, D, L = num_bodies, 6 * num_bodies, rirf_steps
    if masses is None:
        masses = [RM3_MASSES[b % 2] for b in range(N)]
    t = np.linspace(0.0, rirf_duration, L)
    # radiation IRF: K[r,c,s] = a_rc exp(-t/tau_rc) cos(om_rc t), symmetric in (r,c)
    a = rng.uniform(1e4, 1e6, size=(D, D)) * np.where(np.eye(D, dtype=bool), 1.0, 0.1)
    tau = rng.uniform(2.0, 10.0, size=(D, D))
    om = rng.uniform(0.5, 2.0, size=(D, D))
    a, tau, om = [(x + x.T) / 2 for x in (a, tau, om)]
    K = a[:, :, None] * np.exp(-t[None, None, :] / tau[:, :, None]) * np.cos(om[:, :, None] * t[None, None, :])
    K_raw = K / rho
    # added mass: SPD, ~0.5 x body mass on the diagonal
    Q = rng.standard_normal((D, D))
    scale = np.concatenate([[0.5 * masses[b]] * 3 + [5.0 * masses[b]] * 3 for b in range(N)])
    A = 0.02 * (Q @ Q.T) / D
    A = (A + np.eye(D)) * np.sqrt(np.outer(scale, scale))
    A_raw = A / rho
    w = omega_max / num_freqs * np.arange(1, num_freqs + 1)        # uniform grid starting at d_omega
    te = np.linspace(-exc_half_window, exc_half_window, exc_irf_steps)
    bodies = []
    for b in range(N):
        Kh = np.zeros((6, 6))
        heave = rng.uniform(80.0, 500.0)
        Kh[2, 2] = heave
        Kh[3, 3] = rng.uniform(500.0, 5000.0)
        Kh[4, 4] = rng.uniform(500.0, 5000.0)
        Kh[2, 4] = Kh[4, 2] = 0.05 * heave
        Kh[3, 4] = Kh[4, 3] = 1.0
        cg = np.array(RM3_CG[b % 2])
        mag = np.empty((6, 1, num_freqs))
        ph = np.empty((6, 1, num_freqs))
        f_irf = np.empty((6, 1, exc_irf_steps))
        for r in range(6):
            m0, wc, bw = rng.uniform(50.0, 500.0), rng.uniform(0.6, 1.5), rng.uniform(0.3, 0.8)
            mag[r, 0] = m0 * np.exp(-0.5 * ((w - wc) / bw) ** 2)
            ph[r, 0] = rng.uniform(-np.pi, np.pi) + rng.uniform(-0.5, 0.5) * w
            f0, sg, wf, p0 = rng.uniform(20.0, 200.0), rng.uniform(3.0, 8.0), rng.uniform(0.5, 1.5), rng.uniform(0, 6.28)
            f_irf[r, 0] = f0 * np.exp(-0.5 * (te / sg) ** 2) * np.cos(wf * te + p0)
        bodies.append({
            "disp_vol": masses[b] / rho,
            "cg": cg, "cb": cg + np.array([0.0, 0.0, 0.1]),
            "lin_matrix": Kh,
            "inf_added_mass": A_raw[6 * b:6 * b + 6, :].copy(),
            "rirf_K": K_raw[6 * b:6 * b + 6, :, :].copy(),
            "rirf_t": t.copy(),
            "exc_mag": mag, "exc_phase": ph,
            "exc_irf_f": f_irf, "exc_irf_t": te.copy(),
        })
    return {"rho": rho, "g": g, "water_depth": water_depth, "w": w, "bodies": bodies}


def rm3_like(**kw):
    """The headline workload's design: N = 2, D = 12, L = 1001 lags over 60 s (6 x dt at dt = 0.01)."""
    return make_tables(num_bodies=2, **kw)


def prescribed_motion(D, seed=7):
    """Amplitudes / angular frequencies of the synthetic body motion used by bench.py and the CPU baseline:
    pose_d(t) = amp_d sin(om_d t + 0.01 i), vel_d(t) = amp_d om_d cos(om_d t + 0.01 i) for instance i."""
    rng = np.random.Generator(np.random.MT19937(seed))
    amp = rng.uniform(0.05, 0.5, size=D)
    om = rng.uniform(0.4, 1.2, size=D)
    return amp, om
