"""Water kinematics: the product's host arithmetic behind WaveBase::GetElevation / GetVelocity /
GetAcceleration (hc_wave_kinematics, C ABI) against the oracle's line-by-line restatement of
/root/reference/src/wave_types.cpp:14-160 (GetEta, GetWaterVelocity, GetWaterAcceleration and the irregular sums),
:301-313 (RegularWave) and :515-550 (IrregularWaves incl. Wheeler stretching).  No GPU needed: the reference computes
these on the host too.  Bar: 1e-12 relative (the two are the same operations in the same order; in practice bitwise)."""
import numpy as np
import pytest

import common
import hydrochrono_b200 as hc
from oracle import hc_oracle as orc

POINTS = [(0.0, 0.0, 0.0), (3.5, 1.0, -2.0), (-12.0, 0.0, -7.5), (40.0, -3.0, -0.25), (1.0, 0.0, 0.4)]
TIMES = [0.0, 0.37, 11.2, 95.0]


def _tables(depth):
    raw = common.sphere_raw()
    raw = dict(raw)
    raw["water_depth"] = depth
    return raw


def _close(got, ref, what):
    got, ref = np.asarray(got, dtype=float), np.asarray(ref, dtype=float)
    scale = max(1e-300, np.abs(ref).max())
    assert np.all(np.abs(got - ref) <= 1e-12 * scale), (what, got, ref)


@pytest.mark.parametrize("depth", [200.0, 20.0, 8.0, float("inf")])
@pytest.mark.parametrize("mwl", [0.0, 0.3])
def test_regular_wave_kinematics(depth, mwl):
    O = orc.Tables(_tables(depth))
    inst = orc.Instance(O)
    amp, omega, phase = 0.8, 0.9, 0.4
    inst.set_regular(amp, omega, phase)
    k = inst.regular()[2]
    assert k == hc.compute_wave_number(omega, depth, float(common.sphere_raw()["g"]))
    for p in POINTS:
        for t in TIMES:
            eta_r, v_r, a_r = inst.kinematics(p, t, wave_stretching=False, mwl=mwl)
            eta, v, a = hc.wave_kinematics(omega, amp, phase, k, p, t, depth, mwl=mwl)
            _close(eta, eta_r, "eta")
            _close(v, v_r, "velocity")
            _close(a, a_r, "acceleration")
            assert v[1] == 0.0 and a[1] == 0.0
    # analytic sanity (Airy): at the surface under a crest the horizontal velocity is omega A (deep water)
    if depth == float("inf"):
        eta, v, a = hc.wave_kinematics(omega, amp, 0.0, k, (0.0, 0.0, 0.0), 0.0, depth)
        assert abs(eta - amp) < 1e-15 and abs(v[0] - omega * amp) < 1e-15 and abs(a[2] + omega * omega * amp) < 1e-15


@pytest.mark.parametrize("depth", [200.0, 15.0])
@pytest.mark.parametrize("stretch", [False, True])
def test_irregular_wave_kinematics(depth, stretch):
    raw = _tables(depth)
    O = orc.Tables(raw)
    inst = orc.Instance(O)
    kw = dict(dt=0.05, duration=20.0, ramp=0.0, Hs=2.0, Tp=9.0, fmin=0.02, fmax=0.6, nfreq=60, gamma=3.3, seed=7)
    inst.set_irregular(**kw)
    sp = inst.irregular()
    # components the way the host layer forms them from the ensemble's spectrum (wave_types.cpp:38-40)
    f, widths = sp["freqs"], sp["widths"]
    S = hc.jonswap_spectrum_hz(f, kw["Hs"], kw["Tp"], kw["gamma"])
    np.testing.assert_array_equal(S, sp["S"])
    phases = hc.random_phases(kw["seed"], f.size)
    np.testing.assert_array_equal(phases, sp["phases"])
    ks = np.array([hc.compute_wave_number(2 * np.pi * fi, depth, float(raw["g"])) for fi in f])
    np.testing.assert_array_equal(ks, sp["wavenumbers"])
    amp = np.sqrt(2 * S * widths)
    omega = 2 * np.pi * f
    for mwl in (0.0, -0.2):
        for p in POINTS:
            for t in TIMES:
                eta_r, v_r, a_r = inst.kinematics(p, t, wave_stretching=stretch, mwl=mwl)
                eta, v, a = hc.wave_kinematics(omega, amp, phases, ks, p, t, depth, mwl=mwl, wheeler_stretching=stretch)
                _close(eta, eta_r, "eta")
                _close(v, v_r, "velocity")
                _close(a, a_r, "acceleration")
    # the elevation at the origin is what the excitation convolution consumes (un-ramped): eta table cross-check
    j = 700
    eta, _, _ = hc.wave_kinematics(omega, amp, phases, ks, (0.0, 0.0, 0.0), sp["eta_t"][j], depth)
    _close(eta, sp["eta"][j], "eta table")


def test_nowave_kinematics_are_zero():
    O = orc.Tables(_tables(200.0))
    inst = orc.Instance(O)
    inst.set_nowave()
    eta, v, a = inst.kinematics((1.0, 2.0, -3.0), 4.0)
    assert eta == 0.0 and not v.any() and not a.any()
    eta, v, a = hc.wave_kinematics([], [], [], [], (1.0, 2.0, -3.0), 4.0, 200.0)
    assert eta == 0.0 and not v.any() and not a.any()
