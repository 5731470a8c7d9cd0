"""Pins the CPU oracle against the reference's own golden trajectories.

Fixtures: tests/golden/sphere_goldens.npz = reference tests/regression/reference_data/sphere/**.
Pass thresholds of the reference: n1 <= 1e-4, n2 <= 0.02 (tests/regression/sphere/compare.py:49);
the asserts below are far tighter (print precision of the golden files is 1e-6).
"""
import numpy as np
import pytest

import common
import stepper
from oracle import hc_oracle as orc


@pytest.fixture(scope="module")
def tables():
    return orc.Tables(common.sphere_raw())


def _run(inst, tables, z0, nsteps, free=None, damping=None):
    pose0 = np.zeros(6)
    pose0[2] = z0
    return stepper.run(lambda t, x, v: inst.force(t, x, v), tables.added_mass(), [common.SPHERE_MASS],
                       [[1.0, 1.0, 1.0]], pose0, common.SPHERE_DT, nsteps, free=free, damping=damping)


def test_sphere_decay(tables):
    # tests/regression/sphere/demo_sphere_decay.cpp: z0 = -1, unconstrained body, NoWave, 40 s
    g = common.sphere_goldens()["decay_um"] * 1e-6
    inst = orc.Instance(tables)
    t, x = _run(inst, tables, -1.0, g.size)
    n1, n2 = common.traj_norms(x[:, 2], g)
    assert n1 <= 1e-4 and n2 <= 0.02          # the reference's own gate
    assert n2 <= 1.0e-6, (n1, n2)             # print precision of the golden
    # symmetric sphere: nothing but heave is excited
    assert np.abs(x[:, [0, 1, 3, 4, 5]]).max() < 1e-9


@pytest.mark.parametrize("k", range(10))
def test_sphere_regular_waves(tables, k):
    # tests/regression/sphere/reg_waves/sphere_reg_waves_test.cpp: prismatic (heave only), TSDA damper
    g = common.sphere_goldens()["reg%d_um" % (k + 1)] * 1e-6
    inst = orc.Instance(tables)
    inst.set_regular(common.TASK10_AMPS[k], common.TASK10_OMEGAS[k])
    free = np.zeros(6, bool)
    free[2] = True
    damp = np.zeros(6)
    damp[2] = common.TASK10_DAMPING[k]
    t, x = _run(inst, tables, -2.0, g.size, free=free, damping=damp)
    n1, n2 = common.traj_norms(x[:, 2], g)
    assert n1 <= 1e-4 and n2 <= 0.02
    assert n2 <= 2.0e-6, (n1, n2)


def test_sphere_irregular_waves(tables):
    # tests/regression/sphere/irreg_waves/sphere_irreg_waves_test.cpp: Hs 2, Tp 12, nf 1000, ramp 60, seed 1
    g = common.sphere_goldens()["irreg_um"] * 1e-6
    inst = orc.Instance(tables)
    inst.set_irregular(dt=common.SPHERE_DT, duration=600.0, ramp=60.0, Hs=2.0, Tp=12.0, fmin=0.001, fmax=1.0,
                       nfreq=1000, gamma=1.0, seed=1)
    ir = inst.irregular()
    assert ir["eta"].size == 56668 and ir["irf"][0]["t"].size == 8334
    free = np.zeros(6, bool)
    free[2] = True
    t, x = _run(inst, tables, -2.0, g.size, free=free)
    n1, n2 = common.traj_norms(x[:, 2], g)
    assert n1 <= 1e-4 and n2 <= 0.02
    assert n1 <= 1e-7 and n2 <= 2e-4, (n1, n2)


def test_imported_eta_series_reproduces_the_irregular_run(tables):
    """SURVEY a16 (wave_types.cpp:451-453,480-500): the oracle fed the free-surface series of an irregular-wave set-up as
    an imported (time, eta) table gives the same forces bit for bit -- the excitation convolution only sees the table;
    a text round trip with 17 significant digits ("time : eta" lines, the reference's file format) changes nothing."""
    a = orc.Instance(tables)
    a.set_irregular(dt=common.SPHERE_DT, duration=6.0, ramp=1.0, Hs=2.0, Tp=12.0, nfreq=64, seed=9)
    d = a.irregular()
    lines = ["%.17g : %.17g" % (t, e) for t, e in zip(d["eta_t"], d["eta"])]
    tt = np.array([float(x.split(":")[0]) for x in lines]); ee = np.array([float(x.split(":")[1]) for x in lines])
    np.testing.assert_array_equal(tt, d["eta_t"]); np.testing.assert_array_equal(ee, d["eta"])
    b = orc.Instance(tables)
    b.set_irregular_series(common.SPHERE_DT, tt, ee, share_irf_from=a)
    gv = np.array([0.0, 0.0, -9.81])
    t = 0.0
    for n in range(60):
        pose = 0.05 * np.sin(0.9 * t + np.arange(6)); vel = 0.045 * np.cos(0.9 * t + np.arange(6))
        fa, fb = a.force(t, pose, vel, gv), b.force(t, pose, vel, gv)
        np.testing.assert_array_equal(fa, fb)
        t += common.SPHERE_DT
    assert abs(fa[2]) > 1.0
    assert b.irregular()["freqs"].size == 0          # no spectrum behind an imported series


def test_iea_sphere_cli_golden(tables):
    """tests/regression/run_hydrochrono/iea_sphere/decay: the CLI harness' gate is RMS-relative error <= 0.02
    (run_tests.py:235, compare_results.py:103-107).  The golden was produced with Chrono's HHT integrator; the
    linearised-Euler stand-in lands at ~1e-3."""
    g = common.sphere_goldens()
    t_ref, z_ref = g["iea_decay_t"], g["iea_decay_z"]
    inst = orc.Instance(tables)
    free = np.zeros(6, bool)
    free[2] = True
    pose0 = np.zeros(6)
    pose0[2] = -1.0
    gv = (0.0, 0.0, -9.8)
    t, x = stepper.run(lambda t_, x_, v_: inst.force(t_, x_, v_, gv), tables.added_mass(), [261800.0],
                       [[999.0, 999.0, 999.0]], pose0, 0.01, 4000, gvec=gv, free=free)
    assert common.rms_relative_error(z_ref, np.interp(t_ref, t, x[:, 2])) <= 0.02
    assert common.rms_relative_error(z_ref, np.interp(t_ref, t, x[:, 2])) <= 2e-3
