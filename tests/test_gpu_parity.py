"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar (north star): every force component at every step within 1e-9 relative of the reference computation,
with an absolute floor of 1e-9 * 1e-3 * max_t|F component| for components that cross zero
(common.force_tol).  Trajectories: within the reference's regression gates against its golden files.
"""
import numpy as np
import pytest

import common
import stepper
import hydrochrono_b200 as hc
from hydrochrono_b200 import synth
from oracle import hc_oracle as orc

pytestmark = pytest.mark.gpu

G981 = (0.0, 0.0, -9.81)


def _motion(D, B, t, seed=3):
    amp, om = synth.prescribed_motion(D, seed)
    ph = 0.37 * np.arange(B)[:, None] + 0.11 * np.arange(D)[None, :]
    pose = amp * np.sin(om * t + ph)
    vel = amp * om * np.cos(om * t + ph)
    return pose, vel


def _assert_parity(got, ref, what, rel=1e-9):
    got, ref = np.asarray(got), np.asarray(ref)
    tol = common.force_tol(ref, rel=rel)
    err = np.abs(got - ref)
    ratio = np.where(tol > 0, err / np.where(tol > 0, tol, 1.0), np.where(err > 0, np.inf, 0.0))
    worst = np.unravel_index(np.argmax(ratio), err.shape)
    bad = np.unique(np.argwhere(err > tol)[:, 0]) if err.ndim > 1 else np.argwhere(err > tol).ravel()
    assert np.all(err <= tol), "%s: worst at %s: got %r ref %r (err/tol %.3g); %d failing entries along axis 0: %s ... %s" % (
        what, worst, got[worst], ref[worst], ratio[worst], bad.size, bad[:12].tolist(), bad[-4:].tolist())
    return float(ratio.max()) * rel       # worst error in units of max(|F_ref|, floor)


def _run_pair(ens, insts, times, D, gvec=G981, seed=3):
    """Steps the GPU ensemble and the oracle instances through the same prescribed motion."""
    B = len(insts)
    tot, hs, rad, wv = [np.empty((len(times), B, D)) for _ in range(4)]
    rtot, rhs, rrad, rwv = [np.empty((len(times), B, D)) for _ in range(4)]
    for n, t in enumerate(times):
        pose, vel = _motion(D, B, t, seed)
        tot[n] = ens.step(t, pose, vel, gvec)
        hs[n], rad[n], wv[n] = ens.components()
        for b, inst in enumerate(insts):
            rtot[n, b], rhs[n, b], rrad[n, b], rwv[n, b] = inst.force(t, pose[b], vel[b], gvec, components=True)
    return (tot, hs, rad, wv), (rtot, rhs, rrad, rwv)


def _acc_times(n, dt):
    """t_{k+1} = t_k + dt, as Chrono advances ChTime."""
    t = np.empty(n)
    x = 0.0
    for i in range(n):
        t[i] = x
        x += dt
    return t


# ---------------------------------------------------------------------------------------------
# sphere (real BEMIO tables)
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def sphere():
    raw = common.sphere_raw()
    return hc.Tables.from_raw(raw), orc.Tables(raw)


def test_sphere_decay_trajectory_and_forces(sphere):
    """BASELINE config 0: sphere heave decay through the GPU path, single instance (B = 1)."""
    T, O = sphere
    ens = hc.Ensemble(T, batch=1, dt_hint=common.SPHERE_DT)
    inst = orc.Instance(O)
    gold = common.sphere_goldens()["decay_um"] * 1e-6
    pose0 = np.zeros(6)
    pose0[2] = -1.0
    rec_g, rec_o = [], []

    def f_gpu(t, x, v):
        F = ens.step(t, x[None, :], v[None, :], G981)[0].copy()
        rec_g.append(F)
        rec_o.append(inst.force(t, x, v, G981))
        return F

    t, x = stepper.run(f_gpu, T.added_mass(), [common.SPHERE_MASS], [[1.0, 1.0, 1.0]], pose0, common.SPHERE_DT, gold.size)
    n1, n2 = common.traj_norms(x[:, 2], gold)
    assert n1 <= 1e-4 and n2 <= 0.02 and n2 <= 1e-6, (n1, n2)
    _assert_parity(np.array(rec_g), np.array(rec_o), "sphere decay total force")
    assert ens.history_len() == inst.history_len()


def test_sphere_regular_and_hydrostatics_bitwise(sphere):
    T, O = sphere
    B = 5
    ens = hc.Ensemble(T, batch=B, dt_hint=common.SPHERE_DT)
    amps = np.array(common.TASK10_AMPS[:B])
    oms = np.array(common.TASK10_OMEGAS[:B])
    ens.set_waves_regular(amps, oms)
    insts = []
    for b in range(B):
        i = orc.Instance(O)
        i.set_regular(amps[b], oms[b])
        insts.append(i)
        mag, ph, k = ens.regular_coeffs(b)
        rmag, rph, rk = i.regular()
        np.testing.assert_array_equal(mag, rmag)
        np.testing.assert_array_equal(ph, rph)
        assert k == rk
    times = _acc_times(1300, common.SPHERE_DT)        # beyond the 1001-lag window: pruning active
    (tot, hs, rad, wv), (rtot, rhs, rrad, rwv) = _run_pair(ens, insts, times, 6)
    np.testing.assert_array_equal(hs, rhs)            # same arithmetic order, no FMA contraction
    _assert_parity(rad, rrad, "radiation")
    _assert_parity(wv, rwv, "regular-wave excitation")
    _assert_parity(tot, rtot, "total")
    assert ens.history_len() == insts[0].history_len()


def test_sphere_irregular_eta_and_forces(sphere):
    T, O = sphere
    B = 3
    seeds = [1, 2, 77]
    ens = hc.Ensemble(T, batch=B, dt_hint=common.SPHERE_DT)
    kw = dict(dt=common.SPHERE_DT, duration=40.0, ramp=10.0, Hs=2.0, Tp=12.0, fmin=0.001, fmax=1.0, nfreq=500,
              gamma=1.0)
    ens.set_waves_irregular(seeds=seeds, **kw)
    insts = []
    for b in range(B):
        i = orc.Instance(O)
        i.set_irregular(seed=seeds[b], share_irf_from=insts[0] if insts else None, **kw)
        insts.append(i)
        g, r = ens.irregular(b), i.irregular()
        for k in ("freqs", "S", "widths", "phases", "wavenumbers", "eta_t"):
            np.testing.assert_array_equal(g[k], r[k], err_msg=k)
        # eta: identical operation order; device cos vs libm cos differ by <= 2 ulp per term
        assert np.abs(g["eta"] - r["eta"]).max() <= 1e-12 * np.abs(r["eta"]).max()
        assert g["eta"][g["eta_t"] <= 0].max() == 0.0 and np.abs(g["eta"]).max() > 0.1
    times = _acc_times(600, common.SPHERE_DT)
    (tot, hs, rad, wv), (rtot, rhs, rrad, rwv) = _run_pair(ens, insts, times, 6)
    _assert_parity(wv, rwv, "irregular excitation")
    _assert_parity(rad, rrad, "radiation")
    _assert_parity(tot, rtot, "total")


def test_sphere_irregular_golden_trajectory(sphere):
    """The reference's irregular-wave regression case end to end through the GPU path (first 100 s)."""
    T, O = sphere
    ens = hc.Ensemble(T, batch=1, dt_hint=common.SPHERE_DT)
    ens.set_waves_irregular(dt=common.SPHERE_DT, duration=600.0, ramp=60.0, Hs=2.0, Tp=12.0, nfreq=1000, gamma=1.0,
                            seed=1)
    assert ens.irregular_sizes() == (1000, 56668, [8334])
    nsteps = 6700
    gold = common.sphere_goldens()["irreg_um"][:nsteps] * 1e-6
    pose0 = np.zeros(6)
    pose0[2] = -2.0
    free = np.zeros(6, bool)
    free[2] = True
    t, x = stepper.run(lambda t, p, v: ens.step(t, p[None, :], v[None, :], G981)[0], T.added_mass(),
                       [common.SPHERE_MASS], [[1.0, 1.0, 1.0]], pose0, common.SPHERE_DT, nsteps, free=free)
    n1, n2 = common.traj_norms(x[:, 2], gold)
    assert n1 <= 1e-4 and n2 <= 0.02 and n2 <= 2e-4, (n1, n2)


# ---------------------------------------------------------------------------------------------
# RM3-shaped two-body system (synthetic tables), the headline workload's design
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def rm3():
    raw = synth.rm3_like()
    return hc.Tables.from_raw(raw), orc.Tables(raw)


IRR = dict(dt=0.01, duration=20.0, ramp=5.0, Hs=2.5, Tp=8.0, fmin=0.001, fmax=1.0, nfreq=200, gamma=3.3)


@pytest.mark.parametrize("snap,lookahead,rad_kernel", [(0.0, 1, 2), (1e-8, 1, 1), (0.0, 2, 1), (1e-8, 2, 2), (0.0, 3, 2),
                                                        (1e-8, 3, 1), (0.0, 4, 3), (1e-8, 5, 3)])
def test_rm3_irregular_ensemble(rm3, snap, lookahead, rad_kernel):
    """12-DoF coupled radiation + two-body excitation; B = 7 is ragged against the 64-instance lane tile.
    snap = 0 is the bit-faithful bracket test; snap = 1e-8 + excitation look-ahead is what bench.py measures."""
    T, O = rm3
    B = 7
    ens = hc.Ensemble(T, batch=B, dt_hint=0.01, bracket_snap=snap, exc_lookahead=lookahead, rad_kernel=rad_kernel)
    seeds = list(range(1, B + 1))
    ens.set_waves_irregular(seeds=seeds, **IRR)
    insts = []
    for b in range(B):
        i = orc.Instance(O)
        i.set_irregular(seed=seeds[b], share_irf_from=insts[0] if insts else None, **IRR)
        insts.append(i)
    times = _acc_times(700, 0.01)
    (tot, hs, rad, wv), (rtot, rhs, rrad, rwv) = _run_pair(ens, insts, times, 12)
    np.testing.assert_array_equal(hs, rhs)
    worst = _assert_parity(rad, rrad, "radiation")
    _assert_parity(wv, rwv, "irregular excitation")
    _assert_parity(tot, rtot, "total")
    # bit-faithful bracketing differs from the oracle only by summation order / FMA contraction
    assert worst < (1e-11 if snap == 0.0 else 1e-9), worst
    launches = ens.profile()["kernel_launches"]
    # small ensemble through hc_step = the compact graph.  rad_kernel 1 (k_radiation<12>): the convolution kernels plan
    # their own lags and the radiation kernel appends -> radiation, finalize (+ excitation) per step; the DMMA / hybrid
    # radiation kernels keep a plan + append kernel in front
    k = 2 if rad_kernel == 1 else 3
    if lookahead in (2, 4):  # 700 steps = 88 blocks of 8: k kernels per step + 3 per block (+ eta synthesis)
        assert launches == 1 + k * 700 + 3 * 88, launches
    elif lookahead == 3:  # background mode: one more block is prefetched on the side stream
        assert launches == 1 + k * 700 + 3 * 89, launches
    elif lookahead == 5:  # background DMMA mode: prefetched blocks are built in two halves (plan 2 + 2 launches); the
        assert launches == 1 + k * 700 + 3 + 4 * 88 - 1, launches      # last one's second half is still pending
    else:
        assert launches == 1 + (k + 1) * 700, launches


@pytest.mark.parametrize("snap,exc_la,m,rad_la,nb", [(1e-8, 5, 6, 2, 2), (1e-8, 1, 6, 3, 2), (1e-8, 1, 6, 2, 2),
                                                      (0.0, 1, 6, 2, 2), (1e-8, 1, 1, 2, 2), (1e-8, 4, 2, 3, 2),
                                                      (1e-8, 1, 2, 2, 2), (1e-8, 1, 1, 2, 1), (1e-8, 5, 2, 2, 1),
                                                      (1e-8, 1, 2, 2, 3), (1e-8, 5, 1, 3, 3)])
def test_rm3_radiation_lookahead(rm3, snap, exc_la, m, rad_la, nb):
    """Radiation look-ahead (k_rad_block<12> + k_rad_step): the resident rows' share of 8 m steps per pass over the
    history, m = RIRF lag spacing / dt.  With snap = 0 and dt = 0.01 the lags are true interpolations, so every step
    must fall back to the per-step kernel.  rad_la = 2: blocks evaluated one block ahead on a side stream; 3: in-stream."""
    if m == 6:
        T, O = rm3
    else:   # lag spacing 0.01 (m = 1) / 0.02 (m = 2): history window shorter than the run, so rows get pruned;
        #     nb = 1 / 2 / 3 bodies = the D = 6 / 12 / 18 instantiations of k_rad_block and k_step
        raw = synth.make_tables(num_bodies=nb, rirf_steps=401 if m == 1 else 201, rirf_duration=4.0)
        T, O = hc.Tables.from_raw(raw), orc.Tables(raw)
    D = 6 * nb
    B = 5
    ens = hc.Ensemble(T, batch=B, dt_hint=0.01, bracket_snap=snap, exc_lookahead=exc_la, rad_lookahead=rad_la)
    seeds = list(range(3, B + 3))
    ens.set_waves_irregular(seeds=seeds, **IRR)
    insts = []
    for b in range(B):
        i = orc.Instance(O)
        i.set_irregular(seed=seeds[b], share_irf_from=insts[0] if insts else None, **IRR)
        insts.append(i)
    times = _acc_times(700, 0.01)
    (tot, hs, rad, wv), (rtot, rhs, rrad, rwv) = _run_pair(ens, insts, times, D)
    np.testing.assert_array_equal(hs, rhs)
    worst = _assert_parity(rad, rrad, "radiation")
    _assert_parity(wv, rwv, "irregular excitation")
    _assert_parity(tot, rtot, "total")
    assert worst < (1e-11 if snap == 0.0 else 1e-9), worst
    launches = ens.profile()["kernel_launches"]
    st = ens.rad_block_stats()
    ngroups = 3 if nb == 3 else 1       # excitation IRF groups: bodies merged for D = 6 / 12, one per body otherwise
    nblocks = -(-699 // (8 * m))
    per_step = (2 if nb < 3 else 3) + ngroups   # compact graph: radiation (+ plan kernel for D = 18), excitation, finalize
    if snap == 0.0:
        assert launches == 1 + per_step * 700 and st["steps_served"] == 0, (launches, st)
        return
    assert st["steps_served"] == 699, st                # every step but the first (empty history)
    # rad_la = 2 plans one block ahead (one more pass started than blocks served)
    assert st["launches"] == nblocks + (1 if rad_la == 2 else 0), st
    if exc_la == 1 and rad_la == 3:     # step 0 per-step, then 699 steps of 3 kernels + one whole pass per block
        assert launches == 1 + per_step + (2 + ngroups) * 699 + nblocks, launches
    elif exc_la == 1:                   # + one slice of the next block's pass after every step
        assert launches <= 1 + per_step + (2 + ngroups) * 699 + 1 + 699, launches


@pytest.mark.parametrize("rad_la,snap", [(1, 0.0), (1, 1e-8), (2, 1e-8)])
def test_rm3_long_run_window_full(rm3, rad_la, snap):
    """History longer than the RIRF window (6000 steps at dt = 0.01): pruning, ring wrap-around, steady state; with
    the per-step kernel (faithful / snapped brackets) and with the radiation look-ahead blocks."""
    T, O = rm3
    B = 2
    ens = hc.Ensemble(T, batch=B, dt_hint=0.01, rad_lookahead=rad_la, bracket_snap=snap)
    insts = [orc.Instance(O) for _ in range(B)]
    times = _acc_times(6300, 0.01)
    check = set(range(0, 6300, 450)) | set(range(5990, 6300, 7))
    D = 12
    got, want = [], []
    for n, t in enumerate(times):
        pose, vel = _motion(D, B, t)
        F = ens.step(t, pose, vel, G981)
        if n in check:
            ref = np.array([i.force(t, pose[b], vel[b], G981) for b, i in enumerate(insts)])
            if snap == 0.0:     # per step, no floor across the series
                _assert_parity(F[None], ref[None], "step %d" % n, rel=1e-9)
            got.append(F.copy()); want.append(ref)
        else:
            for b, i in enumerate(insts):
                i.force(t, pose[b], vel[b], G981)
    # snapped brackets drop ~1e-11 weights of terms far larger than a component near its zero crossing: judged
    # against the series (floor = 1e-3 of the component's maximum), like every other trajectory test
    _assert_parity(np.array(got), np.array(want), "long run", rel=1e-9)
    assert ens.history_len() == insts[0].history_len()
    assert 6000 <= ens.history_len() <= 6003
    if rad_la == 2:
        st = ens.rad_block_stats()
        assert st["steps_served"] == 6299 and st["launches"] == -(-6299 // 48) + 1, st


def test_regular_waves_two_bodies_phase_quirk(rm3):
    """RegularWave uses body 0's interpolated phases for every body (wave_types.cpp:323) -- preserved."""
    T, O = rm3
    ens = hc.Ensemble(T, batch=2, dt_hint=0.01)
    ens.set_waves_regular([1.0], [2.10])           # demos/rm3/demo_rm3_reg_waves.cpp: A = 1.0, omega = 2.10
    insts = [orc.Instance(O), orc.Instance(O)]
    for i in insts:
        i.set_regular(1.0, 2.10)
    times = _acc_times(150, 0.01)
    (tot, hs, rad, wv), (rtot, rhs, rrad, rwv) = _run_pair(ens, insts, times, 12)
    _assert_parity(wv, rwv, "regular excitation")
    _assert_parity(tot, rtot, "total")
    with pytest.raises(IndexError):
        ens.set_waves_regular([1.0], [50.0])       # omega beyond the frequency table


@pytest.mark.parametrize("rad_la", [1, 2])
def test_tapered_direct_mode(rm3, rad_la):
    """TaperedDirect preprocessing of the kernel (one-off, host); the look-ahead blocks convolve with the same
    processed table."""
    raw = synth.rm3_like()
    T, O = hc.Tables.from_raw(raw), orc.Tables(raw)
    T.set_convolution_mode("TaperedDirect", taper_start_percent=0.5, taper_end_percent=0.9)
    O.set_tapered(start=0.5, end=0.9)
    ens = hc.Ensemble(T, batch=2, dt_hint=0.01, rad_lookahead=rad_la, bracket_snap=1e-8 if rad_la == 2 else 0.0)
    insts = [orc.Instance(O), orc.Instance(O)]
    times = _acc_times(400, 0.01)
    (tot, hs, rad, wv), (rtot, rhs, rrad, rwv) = _run_pair(ens, insts, times, 12)
    _assert_parity(rad, rrad, "tapered radiation")
    assert ens.rad_block_stats()["steps_served"] == (399 if rad_la == 2 else 0)


# ---------------------------------------------------------------------------------------------
# semantics of the reference's TestHydro preserved at the boundary
# ---------------------------------------------------------------------------------------------
def test_time_cache_duplicate_and_window_errors(sphere):
    T, O = sphere
    ens = hc.Ensemble(T, batch=2, dt_hint=common.SPHERE_DT)
    pose, vel = _motion(6, 2, 0.0)
    F0 = ens.step(0.0, pose, vel).copy()
    assert ens.last_recomputed
    # same time value again: cached totals, whatever the state passed in (hydro_forces.cpp:742-744)
    F1 = ens.step(0.0, pose * 3.0, vel * -2.0)
    assert not ens.last_recomputed
    np.testing.assert_array_equal(F0, F1)
    assert ens.history_len() == 1
    ens.step(0.015, pose, vel)
    with pytest.raises(hc.HydroError):             # time went backwards: history no longer brackets the query
        ens.step(0.010, pose, vel)
    # first evaluation with one history entry: radiation is zero (hydro_forces.cpp:580-584)
    ens.reset()
    ens.step(5.0, pose, vel)
    hs, rad, wv = ens.components()
    assert np.all(rad == 0.0) and np.all(wv == 0.0)
    # excitation outside the precomputed eta window throws in the reference (wave_types.cpp:833-840)
    ens.set_waves_irregular(dt=common.SPHERE_DT, duration=5.0, Hs=1.0, Tp=8.0, nfreq=50)
    ens.reset()
    ens.step(0.0, pose, vel)
    with pytest.raises(hc.EtaWindowError):
        ens.step(500.0, pose, vel)


def test_lookahead_misprediction_falls_back(rm3):
    """Excitation look-ahead predicts t + k dt; when the caller's times do not follow the prediction the results must
    not change: every step is recomputed from the actual time and look-ahead switches itself off."""
    T, O = rm3
    B = 3
    ens = hc.Ensemble(T, batch=B, dt_hint=0.01, exc_lookahead=3, rad_lookahead=2, bracket_snap=1e-8)
    kw = dict(IRR)
    ens.set_waves_irregular(seed=4, **kw)
    insts = []
    for b in range(B):
        i = orc.Instance(O)
        i.set_irregular(seed=4, share_irf_from=insts[0] if insts else None, **kw)
        insts.append(i)
    rng = np.random.default_rng(11)
    # 40 regular steps (predictions hold), then jittered steps (predictions fail), then regular again
    dts = np.concatenate([np.full(40, 0.01), rng.uniform(0.004, 0.012, size=60), np.full(40, 0.01)])
    times = np.concatenate([[0.0], np.cumsum(dts)])
    (tot, hs, rad, wv), (rtot, rhs, rrad, rwv) = _run_pair(ens, insts, times, 12)
    _assert_parity(wv, rwv, "excitation with mispredicted look-ahead")
    _assert_parity(tot, rtot, "total")
    # the wave-only entry point (WaveBase::GetForceAtTime) is independent of the look-ahead cache
    t_probe = float(times[70])
    w_at = ens.wave_force_at_time(t_probe)
    _assert_parity(w_at[None], rwv[70][None], "wave-only force at a past time")


def test_irregular_dt_and_ring_growth(sphere):
    """Step size smaller than the hint: the history ring has to grow; irregular step sizes exercise the lerp."""
    T, O = sphere
    ens = hc.Ensemble(T, batch=3, dt_hint=0.05)    # ring sized for ~300 entries
    insts = [orc.Instance(O) for _ in range(3)]
    rng = np.random.default_rng(5)
    times = np.cumsum(rng.uniform(0.004, 0.011, size=2500))
    (tot, hs, rad, wv), (rtot, rhs, rrad, rwv) = _run_pair(ens, insts, times, 6)
    _assert_parity(rad, rrad, "radiation, irregular dt")
    assert ens.history_len() == insts[0].history_len() > 300


def test_bracket_search_with_gaps_and_bursts(sphere):
    """The bracket search guesses a lag's history index from the mean step and walks a few entries before it bisects
    (hydro_forces.cpp:374-381 is a linear walk; the index is unique either way).  A long pause, a burst of tiny steps
    and a change of step size put the guess far off, on both sides, for most lags."""
    T, O = sphere
    ens = hc.Ensemble(T, batch=2, dt_hint=0.01)
    insts = [orc.Instance(O) for _ in range(2)]
    dts = np.concatenate([np.full(300, 0.01), [0.8], np.full(400, 0.001), np.full(250, 0.02), [2.5], np.full(120, 0.013)])
    times = np.concatenate([[0.0], np.cumsum(dts)])
    (tot, hs, rad, wv), (rtot, rhs, rrad, rwv) = _run_pair(ens, insts, times, 6)
    _assert_parity(rad, rrad, "radiation, gaps and bursts")
    _assert_parity(tot, rtot, "total, gaps and bursts")
    assert ens.history_len() == insts[0].history_len()


def test_imported_eta_series(sphere):
    """SURVEY a16 -- the free-surface elevation imported as a (time, eta) series (IrregularWaveParams::eta_file_path_,
    wave_types.cpp:451-453,480-500).  (i) a synthesised series fed back as an import reproduces the irregular-wave forces
    bit for bit; (ii) oracle parity on that grid, per instance and shared; (iii) a coarse, non-uniform grid (true
    interpolation on every tap); (iv) input validation."""
    T, O = sphere
    B, dt = 3, 0.015
    irr = dict(dt=dt, duration=8.0, ramp=2.0, Hs=2.0, Tp=12.0, nfreq=100)
    a = hc.Ensemble(T, batch=B, dt_hint=dt)
    a.set_waves_irregular(seeds=[4, 5, 6], **irr)
    series = [a.irregular(b) for b in range(B)]
    eta_t = series[0]["eta_t"]
    eta = np.stack([s["eta"] for s in series])
    b_ = hc.Ensemble(T, batch=B, dt_hint=dt)
    b_.set_waves_series(dt, eta_t, eta)
    got = b_.irregular(1)
    np.testing.assert_array_equal(got["eta_t"], eta_t)
    np.testing.assert_array_equal(got["eta"], eta[1])
    assert b_.irregular_sizes()[0] == 0                                  # no spectrum behind an imported series
    with pytest.raises(hc.HydroError, match="Spectrum has not been created"):
        hc._check(hc.lib.hc_waves_irregular_spectrum(b_._h, 0, None, None, None, None, None))
    insts = []
    for b in range(B):
        i = orc.Instance(O)
        i.set_irregular_series(dt, eta_t, eta[b], share_irf_from=insts[0] if insts else None)
        insts.append(i)
    times = _acc_times(200, dt)
    fa, fb, fo, wv, rwv = [], [], [], [], []
    for t in times:
        pose, vel = _motion(6, B, t)
        fa.append(a.step(t, pose, vel, G981).copy())
        fb.append(b_.step(t, pose, vel, G981).copy())
        wv.append(b_.components()[2].copy())
        r = [i.force(t, pose[k], vel[k], G981, components=True) for k, i in enumerate(insts)]
        fo.append(np.array([x[0] for x in r])); rwv.append(np.array([x[3] for x in r]))
    np.testing.assert_array_equal(np.array(fa), np.array(fb))            # (i)
    _assert_parity(np.array(wv), np.array(rwv), "excitation, imported series")        # (ii)
    _assert_parity(np.array(fb), np.array(fo), "total, imported series")
    assert np.abs(np.array(wv)).max() > 1e3
    # shared series: every instance sees instance 0's elevation
    c = hc.Ensemble(T, batch=B, dt_hint=dt)
    c.set_waves_series(dt, eta_t, eta[0])
    for n, t in enumerate(times[:40]):
        pose, vel = _motion(6, B, t)
        pose[:] = pose[0]; vel[:] = vel[0]
        F = c.step(t, pose, vel, G981)
        assert np.array_equal(F[0], F[1]) and np.array_equal(F[0], F[2])
    # (iii) coarse non-uniform grid wide enough for the excitation IRF window at every step
    tau = a.irregular(0)["irf"][0]["t"]
    rng = np.random.default_rng(11)
    g = np.cumsum(rng.uniform(0.03, 0.09, size=4000)) + (times[0] - tau[-1] - 1.0)
    assert g[-1] > times[-1] - tau[0] + 1.0
    e = 0.8 * np.sin(0.7 * g) + 0.3 * np.cos(1.9 * g + 0.4)
    d = hc.Ensemble(T, batch=2, dt_hint=dt)
    d.set_waves_series(dt, g, e)
    oi = orc.Instance(O)
    oi.set_irregular_series(dt, g, e)
    wv, rwv = [], []
    for t in times[:120]:
        pose, vel = _motion(6, 2, t)
        d.step(t, pose, vel, G981)
        wv.append(d.components()[2][0].copy())
        rwv.append(oi.force(t, pose[0], vel[0], G981, components=True)[3])
    _assert_parity(np.array(wv)[:, None, :], np.array(rwv)[:, None, :], "excitation, non-uniform imported grid")
    # (iv)
    with pytest.raises(hc.HydroError, match="strictly increasing"):
        d.set_waves_series(dt, np.array([0.0, 1.0, 1.0]), np.zeros(3))
    with pytest.raises(hc.HydroError, match="at least two"):
        d.set_waves_series(dt, np.array([0.0]), np.zeros(1))


def test_gravity_vector_and_body_count_generic_path():
    """3-body system (D = 18) runs the run-time-D radiation kernel; tilted gravity exercises the buoyancy cross term."""
    raw = synth.make_tables(num_bodies=3, rirf_steps=301, rirf_duration=15.0, exc_irf_steps=201, exc_half_window=10.0)
    T, O = hc.Tables.from_raw(raw), orc.Tables(raw)
    B = 4
    ens = hc.Ensemble(T, batch=B, dt_hint=0.05)
    kw = dict(dt=0.05, duration=10.0, ramp=0.0, Hs=1.5, Tp=7.0, nfreq=64, gamma=2.0)
    ens.set_waves_irregular(seed=9, **kw)
    insts = []
    for b in range(B):
        i = orc.Instance(O)
        i.set_irregular(seed=9, share_irf_from=insts[0] if insts else None, **kw)
        insts.append(i)
    times = _acc_times(420, 0.05)
    (tot, hs, rad, wv), (rtot, rhs, rrad, rwv) = _run_pair(ens, insts, times, 18, gvec=(0.3, -0.2, -9.7))
    np.testing.assert_array_equal(hs, rhs)
    _assert_parity(rad, rrad, "radiation D=18")
    _assert_parity(wv, rwv, "excitation D=18")
    _assert_parity(tot, rtot, "total D=18")


@pytest.mark.parametrize("name,cfg", [
    # BASELINE.json configs 2..4 as parity cases (their .h5 files are stripped from the reference snapshot, so the
    # tables are synthetic with the demos' shapes and step sizes)
    ("oswec_jonswap", dict(tables=dict(num_bodies=2, rirf_steps=601, rirf_duration=30.0, exc_irf_steps=401,
                                       exc_half_window=24.0), dt=0.03, steps=1400,
                           sea=dict(Hs=1.5, Tp=10.0, gamma=3.3, nfreq=150, ramp=3.0))),     # demos/oswec: N = 2, dt = 0.03
    ("deepcwind_long_rirf", dict(tables=dict(num_bodies=1, rirf_steps=4001, rirf_duration=320.0, exc_irf_steps=601,
                                             exc_half_window=48.0), dt=0.08, steps=4100,
                                 sea=dict(Hs=6.0, Tp=12.0, gamma=2.2, nfreq=120, ramp=8.0))),  # demos/DeepCWind: N = 1, dt = 0.08
    ("f3of_three_bodies", dict(tables=dict(num_bodies=3, rirf_steps=401, rirf_duration=20.0, exc_irf_steps=301,
                                           exc_half_window=15.0), dt=0.02, steps=1300,
                               sea=dict(Hs=1.0, Tp=6.0, gamma=1.0, nfreq=100, ramp=0.0))),   # demos/f3of: N = 3, D = 18
])
@pytest.mark.parametrize("rad_la", [1, 2])
def test_baseline_config_shapes(name, cfg, rad_la):
    """rad_la = 2: radiation look-ahead blocks.  DeepCWind's lag spacing equals dt (lag-grid blocks from the second
    step on); OSWEC's and F3OF's ratios are 5/3 and 5/2, so their blocks convolve the rows with the row-grid kernel
    (interpolation weights folded in) once the history window is full (1000 steps)."""
    raw = synth.make_tables(**cfg["tables"])
    T, O = hc.Tables.from_raw(raw), orc.Tables(raw)
    N = cfg["tables"]["num_bodies"]
    D, dt, steps = 6 * N, cfg["dt"], cfg["steps"]
    B = 3
    ens = hc.Ensemble(T, batch=B, dt_hint=dt, exc_lookahead=5 if N <= 2 else 0, rad_lookahead=rad_la,
                      bracket_snap=1e-8 if rad_la == 2 else 0.0)
    kw = dict(dt=dt, duration=steps * dt + 1.0, **cfg["sea"])
    seeds = [3, 4, 5]
    ens.set_waves_irregular(seeds=seeds, **kw)
    insts = []
    for b in range(B):
        i = orc.Instance(O)
        i.set_irregular(seed=seeds[b], share_irf_from=insts[0] if insts else None, **kw)
        insts.append(i)
    times = _acc_times(steps, dt)
    check = set(range(0, steps, max(1, steps // 40))) | set(range(steps - 60, steps))
    got, ref = [], []
    for n, t in enumerate(times):
        pose, vel = _motion(D, B, t)
        F = ens.step(t, pose, vel, G981)
        r = np.array([i.force(t, pose[b], vel[b], G981) for b, i in enumerate(insts)])
        if n in check:
            got.append(F.copy())
            ref.append(r)
    _assert_parity(np.array(got), np.array(ref), name)
    assert ens.history_len() == insts[0].history_len()
    if rad_la == 2:
        served = ens.rad_block_stats()["steps_served"]
        full_window = int(round(cfg["tables"]["rirf_duration"] / dt))
        assert served >= (steps - 8 if name.startswith("deepcwind") else steps - full_window - 24), served


def test_added_mass_mv(rm3):
    T, O = rm3
    B, n_sys = 9, 18                               # one extra non-hydro body in the system
    ens = hc.Ensemble(T, batch=B, dt_hint=0.01)
    rng = np.random.default_rng(2)
    w = rng.standard_normal((B, n_sys))
    R = rng.standard_normal((B, n_sys))
    out = ens.added_mass_mv(0.37, w, R)
    ref = np.array([O.added_mass_mv(0.37, w[b], R[b]) for b in range(B)])
    np.testing.assert_array_equal(out, ref)
    np.testing.assert_array_equal(out[:, 12:], R[:, 12:])


# ---------------------------------------------------------------------------------------------
# full-size ensemble: size-independent properties (the oracle cannot follow 16384 instances)
# ---------------------------------------------------------------------------------------------
def test_large_ensemble_properties(rm3):
    import torch
    T, O = rm3
    B, D = 16384, 12
    ens = hc.Ensemble(T, batch=B, dt_hint=0.05, bracket_snap=1e-9)
    kw = dict(dt=0.05, duration=12.0, ramp=2.0, Hs=2.5, Tp=8.0, nfreq=32, gamma=3.3)
    # instances b and b + B/2 share a seed: identical realisations must give identical forces
    seeds = np.concatenate([np.arange(1, B // 2 + 1), np.arange(1, B // 2 + 1)]).astype(np.int32)
    ens.set_waves_irregular(seeds=seeds, **kw)
    amp, om = synth.prescribed_motion(D)
    half = B // 2
    ph = (0.01 * np.arange(half))[:, None]
    sample = [0, 1, 63, 64, 511, 512, half - 1]
    insts = []
    for b in sample:
        i = orc.Instance(O)
        i.set_irregular(seed=int(seeds[b]), share_irf_from=insts[0] if insts else None, **kw)
        insts.append(i)
    times = _acc_times(230, 0.05)
    dev = torch.device("cuda", 0)
    d_pose = torch.empty((B, D), dtype=torch.float64, device=dev)
    d_vel = torch.empty_like(d_pose)
    d_force = torch.empty_like(d_pose)
    for n, t in enumerate(times):
        p = amp * np.sin(om * t + ph)
        v = amp * om * np.cos(om * t + ph)
        pose = np.concatenate([p, p])
        vel = np.concatenate([v, v])
        d_pose.copy_(torch.from_numpy(pose))
        d_vel.copy_(torch.from_numpy(vel))
        torch.cuda.synchronize()
        ens.step_device(t, d_pose, d_vel, d_force)          # device-resident inputs/outputs
        ens.sync()
        F = d_force.cpu().numpy()
        assert np.array_equal(F[:half], F[half:])           # replicas agree bit for bit
        if n % 23 == 0 or n > 220:
            ref = np.array([i.force(t, pose[b], vel[b], G981) for b, i in zip(sample, insts)])
            _assert_parity(F[sample][None], ref[None], "sampled instances, step %d" % n)
        else:
            for b, i in zip(sample, insts):
                i.force(t, pose[b], vel[b], G981)
    # linearity of the radiation term in the velocity history: F_rad(2v) == 2 F_rad(v) (exact in binary)
    e1 = hc.Ensemble(T, batch=64, dt_hint=0.05)
    e2 = hc.Ensemble(T, batch=64, dt_hint=0.05)
    for t in _acc_times(40, 0.05):
        pose, vel = _motion(D, 64, t)
        e1.step(t, pose, vel)
        e2.step(t, pose, 2.0 * vel)
    r1, r2 = e1.components()[1], e2.components()[1]
    np.testing.assert_array_equal(2.0 * r1, r2)


@pytest.mark.gpu
def test_large_ensemble_lookahead_paths(rm3):
    """The bench configuration's kernels at the bench batch size: 16384 instances, dt = 0.01 (6 history rows per RIRF
    lag), snapped brackets -> radiation look-ahead blocks (k_rad_block<12> / k_step<12>) and excitation look-ahead blocks
    (k_exc_block_mma) are selected automatically.  Replicated seeds must agree bit for bit, sampled instances must
    match the oracle, and the radiation term must stay exactly linear in the velocity history."""
    import torch
    T, O = rm3
    B, D, dt = 16384, 12, 0.01
    ens = hc.Ensemble(T, batch=B, dt_hint=dt, bracket_snap=1e-8)
    assert ens.rad_lookahead_steps() == 48
    kw = dict(dt=dt, duration=4.0, ramp=1.0, Hs=2.5, Tp=8.0, nfreq=32, gamma=3.3)
    seeds = np.concatenate([np.arange(1, B // 2 + 1), np.arange(1, B // 2 + 1)]).astype(np.int32)
    ens.set_waves_irregular(seeds=seeds, **kw)
    amp, om = synth.prescribed_motion(D)
    half = B // 2
    ph = (0.01 * np.arange(half))[:, None]
    sample = [0, 1, 15, 16, 63, 64, 4095, half - 1]
    insts = []
    for b in sample:
        i = orc.Instance(O)
        i.set_irregular(seed=int(seeds[b]), share_irf_from=insts[0] if insts else None, **kw)
        insts.append(i)
    nsteps = 260
    times = _acc_times(nsteps, dt)
    dev = torch.device("cuda", 0)
    d_pose = torch.empty((B, D), dtype=torch.float64, device=dev)
    d_vel = torch.empty_like(d_pose)
    d_force = torch.empty_like(d_pose)
    got, want = [], []
    for n, t in enumerate(times):
        p = amp * np.sin(om * t + ph)
        v = amp * om * np.cos(om * t + ph)
        pose = np.concatenate([p, p])
        vel = np.concatenate([v, v])
        d_pose.copy_(torch.from_numpy(pose))
        d_vel.copy_(torch.from_numpy(vel))
        torch.cuda.synchronize()
        ens.step_device(t, d_pose, d_vel, d_force)
        ens.sync()
        F = d_force.cpu().numpy()
        assert np.array_equal(F[:half], F[half:])           # replicas agree bit for bit
        ref = np.array([i.force(t, pose[b], vel[b], G981) for b, i in zip(sample, insts)])
        got.append(F[sample].copy()); want.append(ref)
    _assert_parity(np.array(got), np.array(want), "sampled instances")
    st = ens.rad_block_stats()
    assert st["steps_served"] == nsteps - 1, st
    launches = ens.profile()["kernel_launches"]
    assert launches < 1 + 5 + 2 * nsteps + 3 * (nsteps // 8 + 2) + 8, launches     # k_step<12> + slice per step, blocks
    ens.close()
    # exact linearity of the block path in the velocity history
    e1 = hc.Ensemble(T, batch=64, dt_hint=dt, bracket_snap=1e-8, rad_lookahead=2)
    e2 = hc.Ensemble(T, batch=64, dt_hint=dt, bracket_snap=1e-8, rad_lookahead=2)
    for t in _acc_times(120, dt):
        pose, vel = _motion(D, 64, t)
        e1.step(t, pose, vel)
        e2.step(t, pose, 2.0 * vel)
    assert e1.rad_block_stats()["steps_served"] == 119
    r1, r2 = e1.components()[1], e2.components()[1]
    np.testing.assert_array_equal(2.0 * r1, r2)


@pytest.mark.gpu
def test_radiation_lookahead_reset_and_wrong_hint(rm3):
    """(a) hc_ensemble_reset in the middle of a block: the second run reproduces the first bit for bit.
    (b) dt_hint twice the step actually used: every prediction misses, the per-step kernels serve all steps, the
    history ring (sized from the hint) has to grow, and the results still match the oracle."""
    T, O = rm3
    B, D, dt = 4, 12, 0.01
    ens = hc.Ensemble(T, batch=B, dt_hint=dt, bracket_snap=1e-8, rad_lookahead=2)
    times = _acc_times(131, dt)          # stops inside the third block of 48

    def run():
        out = []
        for t in times:
            pose, vel = _motion(D, B, t)
            out.append(ens.step(t, pose, vel, G981).copy())
        return np.array(out)

    first = run()
    assert ens.rad_block_stats()["steps_served"] == 130
    ens.reset()
    second = run()
    assert ens.rad_block_stats()["steps_served"] == 130
    np.testing.assert_array_equal(first, second)
    ens.close()

    ens = hc.Ensemble(T, batch=B, dt_hint=2 * dt, bracket_snap=1e-8, rad_lookahead=2)
    insts = [orc.Instance(O) for _ in range(B)]
    times = _acc_times(3300, dt)         # ring sized for 60 s / 0.02 = 3000 rows (+ slack)
    got, want = [], []
    for n, t in enumerate(times):
        pose, vel = _motion(D, B, t)
        F = ens.step(t, pose, vel, G981)
        ref = np.array([i.force(t, pose[b], vel[b], G981) for b, i in enumerate(insts)])
        if n % 97 == 0 or n > 3250:
            got.append(F.copy()); want.append(ref)
    _assert_parity(np.array(got), np.array(want), "wrong dt_hint")
    assert ens.rad_block_stats()["steps_served"] <= 3        # at most the first step of a block that then misses
    assert ens.history_len() == insts[0].history_len() > 3100


@pytest.mark.gpu
def test_mixing_host_and_device_stepping_inside_a_block():
    """hc_step (equal pass slices) and hc_step_device (wave-aligned pass slices) alternate in the middle of look-ahead
    blocks: the slices of every pass must still cover all of its work items.  Batch large enough (2304 instances = 36
    instance tiles x 2 chunks x 6 classes = 432+ items) for the two slicings to differ."""
    import torch
    raw = synth.make_tables(num_bodies=2, rirf_steps=101, rirf_duration=6.0)       # lag spacing 0.06 = 6 dt
    T, O = hc.Tables.from_raw(raw), orc.Tables(raw)
    B, D, dt = 2304, 12, 0.01
    ens = hc.Ensemble(T, batch=B, dt_hint=dt, bracket_snap=1e-8, rad_lookahead=2)
    assert ens.rad_lookahead_steps() == 48
    sample = [0, 63, 64, 1000, B - 1]
    insts = [orc.Instance(O) for _ in sample]
    dev = torch.device("cuda", 0)
    d_pose = torch.empty((B, D), dtype=torch.float64, device=dev)
    d_vel = torch.empty_like(d_pose)
    d_force = torch.empty_like(d_pose)
    got, want = [], []
    for n, t in enumerate(_acc_times(700, dt)):
        pose, vel = _motion(D, B, t)
        if (n // 19) % 2 == 0:                      # switch the API every 19 steps: never aligned with the 48-step blocks
            F = ens.step(t, pose, vel, G981)
        else:
            d_pose.copy_(torch.from_numpy(pose)); d_vel.copy_(torch.from_numpy(vel))
            torch.cuda.synchronize()
            ens.step_device(t, d_pose, d_vel, d_force, G981)
            ens.sync()
            F = d_force.cpu().numpy()
        ref = np.array([i.force(t, pose[b], vel[b], G981) for b, i in zip(sample, insts)])
        got.append(F[sample].copy()); want.append(ref)
    _assert_parity(np.array(got), np.array(want), "mixed stepping")
    assert ens.rad_block_stats()["steps_served"] == 699


@pytest.mark.gpu
@pytest.mark.parametrize("snap,pass_mode", [(1e-8, 1), (1e-8, 2), (0.0, 1)])
def test_benchmark_state_parity(rm3, snap, pass_mode):
    """Exactly what bench.py times: B = 16384, D = 12, L = 1001 lags over 60 s, dt = 0.01,
    Le = 6000, nf = 1000 -- stepped from an empty history through the full window (6001 rows), past the point where
    the history ring wraps, over > 120 radiation blocks of 48 steps; hc_step_device and hc_step alternate in runs of
    7 / 5 steps.  Eight sampled instances are compared with the oracle at EVERY step.
    Reference: src/hydro_forces.cpp:537-691 (radiation + history), src/wave_types.cpp:776-844 (excitation)."""
    import torch
    T, O = rm3
    B, D, dt, NB = 16384, 12, 0.01, 8
    nsteps = 6200
    ens = hc.Ensemble(T, batch=B, dt_hint=dt, bracket_snap=snap, rad_pass_mode=pass_mode)
    kw = dict(dt=dt, duration=(nsteps + 64) * dt, ramp=20.0, Hs=2.5, Tp=8.0, fmin=0.001, fmax=1.0, nfreq=1000, gamma=3.3)
    seeds = np.arange(1, B + 1, dtype=np.int32)
    ens.set_waves_irregular(seeds=seeds, **kw)
    assert ens.irregular_sizes()[2] == [6000, 6000]
    amp, om = synth.prescribed_motion(D)
    ph = 0.01 * np.arange(B)[:, None]
    pose = np.stack([amp * np.sin(om * (i * dt) + ph) for i in range(NB)])
    vel = np.stack([amp * om * np.cos(om * (i * dt) + ph) for i in range(NB)])
    dev = torch.device("cuda", 0)
    h_pose = [torch.from_numpy(pose[i]).pin_memory() for i in range(NB)]
    h_vel = [torch.from_numpy(vel[i]).pin_memory() for i in range(NB)]
    d_pose = [x.to(dev) for x in h_pose]
    d_vel = [x.to(dev) for x in h_vel]
    d_force = torch.empty((B, D), dtype=torch.float64, device=dev)
    h_force = torch.empty((B, D), dtype=torch.float64).pin_memory()
    sample = [0, 1, 63, 64, 4097, 8191, 12345, B - 1]
    d_idx = torch.tensor(sample, device=dev)
    times = _acc_times(nsteps, dt)
    got = np.empty((nsteps, len(sample), D))
    for n in range(nsteps):
        if (n % 12) < 7:
            ens.step_device(times[n], d_pose[n % NB], d_vel[n % NB], d_force)
            ens.sync()
            got[n] = d_force[d_idx].cpu().numpy()
        else:
            ens.step(times[n], h_pose[n % NB].numpy(), h_vel[n % NB].numpy(), G981, out=h_force.numpy())
            got[n] = h_force.numpy()[sample]
    hist_len = ens.history_len()
    st = ens.rad_block_stats()
    if snap > 0:
        assert st["steps_served"] >= nsteps - 2, st              # every step but the first ones came from a block
        assert st["launches"] >= nsteps // 48, st
    else:
        assert ens.lookahead_state()["radiation"] in (0, -1)     # bit-faithful brackets: the per-step kernels serve
    ens.close()
    insts = []
    for b in sample:
        i = orc.Instance(O, omp_mode=0)
        i.set_irregular(seed=int(seeds[b]), share_irf_from=insts[0] if insts else None, **kw)
        insts.append(i)
    _, _, ref = orc.bench_lockstep(insts, times, pose[:, sample, :].copy(), vel[:, sample, :].copy(), buf0=0, mode=1,
                                   gvec=G981, want_forces=True)
    assert insts[0].history_len() == hist_len >= 6001          # full window (+ the extra entry kept for bracketing)
    worst = _assert_parity(got, ref, "benchmark state, snap %g, pass mode %d" % (snap, pass_mode))
    print("benchmark-state parity: worst relative error %.2e over %d steps x %d instances" % (worst, nsteps, len(sample)))


@pytest.mark.gpu
def test_lookahead_restage_after_auto_disable_and_reset_rearm():
    """the radiation look-ahead switches itself off after mispredicted blocks.  A TaperedDirect switch +
    refresh_rirf while it is off must still re-stage the block path's kernel tables, and hc_ensemble_reset must arm
    the path again: the next run is served by the blocks and convolves with the NEW kernel."""
    raw = synth.rm3_like()
    T, O = hc.Tables.from_raw(raw), orc.Tables(raw)
    B, D, dt = 64, 12, 0.01
    ens = hc.Ensemble(T, batch=B, dt_hint=dt, bracket_snap=1e-8, rad_lookahead=2)
    check = [0, 31, 63]
    assert ens.lookahead_state()["radiation"] == 1
    # irregular step sizes: every prediction misses, three poor blocks switch the path off
    t = 0.0
    for n in range(40):
        pose, vel = _motion(D, B, t)
        ens.step(t, pose, vel)
        t += dt * (1.0 + 0.37 * ((n * 7) % 5))
    assert ens.lookahead_state()["radiation"] == 0
    # switch the kernel while the path is off
    T.set_convolution_mode("TaperedDirect", smoothing="sg", window_length=5, rirf_end_time=40.0, taper_start_percent=0.7,
                           taper_end_percent=0.95, taper_final_amplitude=0.0)
    O.set_tapered(smoothing="sg", window_length=5, rirf_end_time=40.0, start=0.7, end=0.95, final=0.0)
    ens.refresh_rirf()
    ens.reset()                                                   # a new run: empty history, look-ahead armed again
    assert ens.lookahead_state()["radiation"] == 1
    ens.rad_block_stats(reset=True)
    insts = [orc.Instance(O) for _ in check]
    got, want = [], []
    for tt in _acc_times(200, dt):
        pose, vel = _motion(D, B, tt)
        F = ens.step(tt, pose, vel)
        got.append(F[check].copy())
        want.append(np.array([i.force(tt, pose[b], vel[b], G981) for b, i in zip(check, insts)]))
    assert ens.rad_block_stats(reset=False)["steps_served"] >= 190
    _assert_parity(np.array(got), np.array(want), "re-staged kernel after auto-disable")
    ens.close()


@pytest.mark.gpu
def test_eta_window_error_leaves_reference_state():
    """The reference throws from ComputeForceWaves (wave_types.cpp:833-840) after ComputeForceRadiationDampingConv has
    pushed the sample and prev_time was set (hydro_forces.cpp:747-756): the history keeps the sample and a second call
    at the same time returns the cached (zero) totals instead of throwing again."""
    raw = synth.rm3_like(exc_irf_steps=201, exc_half_window=5.0)     # eta window ends at t ~ 6 s, long before the
    T, O = hc.Tables.from_raw(raw), orc.Tables(raw)                  # 60 s radiation window starts pruning
    B, D, dt = 3, 12, 0.05
    ens = hc.Ensemble(T, batch=B, dt_hint=dt)
    kw = dict(dt=dt, duration=1.0, ramp=0.5, Hs=2.5, Tp=8.0, nfreq=16, gamma=3.3)
    ens.set_waves_irregular(seeds=np.arange(1, B + 1, dtype=np.int32), **kw)
    inst = orc.Instance(O)
    inst.set_irregular(seed=1, **kw)
    t, n_ok = 0.0, 0
    while True:
        pose, vel = _motion(D, B, t)
        try:
            ref = inst.force(t, pose[0], vel[0], G981)
        except orc.OracleError:
            break
        F = ens.step(t, pose, vel)
        _assert_parity(F[0], ref, "inside the window")
        n_ok += 1
        t += dt
    assert n_ok > 5
    hl = ens.history_len()
    with pytest.raises(hc.HydroError) as ei:
        ens.step(t, pose, vel)
    assert "out of bounds" in str(ei.value)
    assert ens.history_len() == hl + 1 == inst.history_len()      # sample pushed before the throw, as the reference
    F = ens.step(t, pose, vel)                                    # cached totals of that time: the zeros of the reset
    assert ens.last_recomputed is False and np.all(F == 0.0)
    ens.close()


@pytest.mark.gpu
def test_multi_device_handle_matches_single_ensemble(rm3):
    """hc_multi_* (SURVEY 8e): the instances partitioned into contiguous shards, one hc_ensemble + one host thread per
    shard.  Run here as 3 uneven shards on whatever devices exist (several shards may share a GPU).  Forces and the
    gathered components must equal, bit for bit, what three stand-alone ensembles of the shard sizes give (seeds follow
    the GLOBAL instance index), and agree with one ensemble of all 200 instances to rounding (the lag-chunk and
    eta-segment partitions, hence the summation grouping, depend on the batch size)."""
    T, O = rm3
    B, D, dt = 200, 12, 0.01
    ndev = hc.device_count()
    devices = [i % ndev for i in range(3)]
    kw = dict(dt=dt, duration=3.0, ramp=1.0, Hs=2.5, Tp=8.0, nfreq=24, gamma=3.3)
    opts = dict(dt_hint=dt, bracket_snap=1e-8, rad_lookahead=2, exc_lookahead=4)
    seeds = np.arange(11, 11 + B, dtype=np.int32)
    one = hc.Ensemble(T, batch=B, **opts)
    one.set_waves_irregular(seeds=seeds, **kw)
    multi = hc.MultiEnsemble(T, batch=B, devices=devices, **opts)
    multi.set_waves_irregular(seeds=seeds, **kw)
    sh = multi.shards()
    assert [s["count"] for s in sh] == [67, 67, 66] and [s["first"] for s in sh] == [0, 67, 134]
    assert [s["device"] for s in sh] == devices
    parts = []
    for s in sh:
        e = hc.Ensemble(T, batch=s["count"], device=s["device"], **opts)
        e.set_waves_irregular(seeds=seeds[s["first"]:s["first"] + s["count"]], **kw)
        parts.append(e)
    for t in _acc_times(130, dt):
        pose, vel = _motion(D, B, t)
        F1 = one.step(t, pose, vel)
        F2 = multi.step(t, pose, vel)
        F3 = np.concatenate([e.step(t, pose[s["first"]:s["first"] + s["count"]], vel[s["first"]:s["first"] + s["count"]])
                             for e, s in zip(parts, sh)])
        np.testing.assert_array_equal(F2, F3)
        _assert_parity(F2, F1, "sharded vs one ensemble", rel=1e-11)
    for a, b in zip(multi.components(), [np.concatenate(c) for c in zip(*[e.components() for e in parts])]):
        np.testing.assert_array_equal(a, b)
    F4 = multi.step(t, pose, vel)                           # same time again: every shard serves its cache
    assert multi.last_recomputed is False
    np.testing.assert_array_equal(F4, F2)
    inst = orc.Instance(O)
    inst.set_irregular(seed=int(seeds[150]), **kw)
    ref = [inst.force(tt, *[x[150] for x in _motion(D, B, tt)], G981) for tt in _acc_times(130, dt)]
    _assert_parity(F2[150], ref[-1], "instance 150 (third shard)")
    with pytest.raises(hc.HydroError):                      # time going backwards is reported with the shard it came from
        multi.step(0.0, pose, vel)
    one.close(); multi.close()
    # imported per-instance free-surface series: every shard receives its own rows of eta[B][n]
    eta_t = parts[0].irregular(0)["eta_t"]
    eta = np.stack([parts[k].irregular(i)["eta"] for k, s in enumerate(sh) for i in (0, s["count"] - 1)])   # 6 series
    m2 = hc.MultiEnsemble(T, batch=6, devices=devices, dt_hint=dt)
    m2.set_waves_series(dt, eta_t, eta)
    singles = []
    for i in range(6):
        e = hc.Ensemble(T, batch=2, dt_hint=dt)            # (shards of 6 over 3 devices hold 2 instances each)
        e.set_waves_series(dt, eta_t, eta[2 * (i // 2):2 * (i // 2) + 2])
        singles.append(e)
    for t in _acc_times(20, dt):
        pose, vel = _motion(D, 6, t)
        F = m2.step(t, pose, vel)
        for k in range(3):
            np.testing.assert_array_equal(F[2 * k:2 * k + 2], singles[2 * k].step(t, pose[2 * k:2 * k + 2], vel[2 * k:2 * k + 2]))
    assert np.abs(m2.components()[2]).max() > 1.0
    m2.close()
    for e in parts + singles:
        e.close()


@pytest.mark.gpu
def test_pinned_and_pageable_host_buffers_agree(rm3):
    """hc_step on a modest ensemble served by the look-aheads (the 2048-per-GPU split of the north star) with pinned
    host buffers and with pageable arrays: bit-identical results, both matching the oracle."""
    import torch
    T, O = rm3
    B, D, dt = 1024, 12, 0.01
    opts = dict(dt_hint=dt, bracket_snap=1e-8, rad_lookahead=2, exc_lookahead=4)
    kw = dict(dt=dt, duration=4.0, ramp=1.0, Hs=2.5, Tp=8.0, nfreq=24, gamma=3.3)
    seeds = np.arange(1, B + 1, dtype=np.int32)
    a = hc.Ensemble(T, batch=B, **opts)
    b = hc.Ensemble(T, batch=B, **opts)
    a.set_waves_irregular(seeds=seeds, **kw)
    b.set_waves_irregular(seeds=seeds, **kw)
    st = torch.empty((2, B, D), dtype=torch.float64).pin_memory()         # [vel, pose] adjacent: one upload
    hv, hp = st[0], st[1]
    hf = torch.empty((B, D), dtype=torch.float64).pin_memory()
    check = [0, 31, 32, 500, B - 1]
    insts = []
    for k in check:
        i = orc.Instance(O)
        i.set_irregular(seed=int(seeds[k]), share_irf_from=insts[0] if insts else None, **kw)
        insts.append(i)
    got, want = [], []
    for t in _acc_times(150, dt):
        pose, vel = _motion(D, B, t)
        hp.copy_(torch.from_numpy(pose)); hv.copy_(torch.from_numpy(vel))
        Fa = a.step(t, hp.numpy(), hv.numpy(), G981, out=hf.numpy())       # pinned in, pinned out
        Fb = b.step(t, pose.copy(), vel.copy(), G981)                      # pageable
        np.testing.assert_array_equal(Fa, Fb)
        got.append(Fa[check].copy())
        want.append(np.array([i.force(t, pose[k], vel[k], G981) for k, i in zip(check, insts)]))
    assert a.rad_block_stats(reset=False)["steps_served"] >= 148
    for x, y in zip(a.components(), b.components()):
        np.testing.assert_array_equal(x, y)
    _assert_parity(np.array(got), np.array(want), "pinned host buffers")
    a.close(); b.close()
