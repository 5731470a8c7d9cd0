"""C++ host layer (hydrochrono_b200/host): HydroChrono's own class surface -- TestHydro, ForceFunc6d, ComponentFunc,
ChLoadAddedMass, H5FileInfo/HydroData, NoWave/RegularWave/IrregularWaves, ReadHydroYAML, SetupHydroFromYAML -- over
the C ABI.  The demo mains mirror the reference's regression mains (tests/regression/sphere/**) call for call and
their trajectories are held against the reference's golden files."""
import os
import subprocess

import numpy as np
import pytest

import common
from hydrochrono_b200 import h5io, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "hydrochrono_b200", "host")
BUILD = os.path.join(HOST, "build")


@pytest.fixture(scope="module")
def host_build():
    subprocess.check_call(["make", "-C", HOST, "-j8"], stdout=subprocess.DEVNULL)
    return BUILD


@pytest.fixture(scope="module")
def sphere_h5(tmp_path_factory):
    f = tmp_path_factory.mktemp("h5") / "sphere.h5"
    h5io.write_bemio(f, common.sphere_raw())
    return str(f)


def _read_traj(path):
    a = np.loadtxt(path, skiprows=1)
    return a[:, 0], a[:, 1]


def test_yaml_parser(host_build):
    out = subprocess.run([os.path.join(host_build, "test_hydro_yaml_parser"), os.path.join(HOST, "tests", "data")],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "All hydro YAML parser tests passed" in out.stdout


def test_compat_steppers_converge(host_build):
    """chrono_compat's linearised-Euler (order 1) and HHT-alpha (order 2) steps against a closed-form oscillator whose
    force comes from state-reading, time-keyed ChFunction callbacks (the way ComponentFunc feeds Chrono).  CPU only."""
    out = subprocess.run([os.path.join(host_build, "test_stepper")], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "stepper test passed" in out.stdout


def test_host_layer_fails_loudly_without_gpu(host_build, sphere_h5, tmp_path):
    import hydrochrono_b200 as hc
    if hc.device_count() > 0:
        pytest.skip("CUDA device present")
    out = subprocess.run([os.path.join(host_build, "demo_sphere_decay"), sphere_h5, str(tmp_path / "o.txt")],
                         capture_output=True, text=True)
    assert out.returncode == 1
    assert "no CPU fallback" in out.stderr


@pytest.mark.gpu
def test_demo_sphere_decay_golden(host_build, sphere_h5, tmp_path):
    o = tmp_path / "decay.txt"
    out = subprocess.run([os.path.join(host_build, "demo_sphere_decay"), sphere_h5, str(o)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    t, z = _read_traj(o)
    gold = common.sphere_goldens()["decay_um"] * 1e-6
    assert z.size == gold.size
    n1, n2 = common.traj_norms(z, gold)
    assert n1 <= 1e-4 and n2 <= 0.02 and n2 <= 1.0e-6, (n1, n2)   # reference gate, then print precision
    assert "radiation_calls %d" % gold.size in out.stdout           # exactly one device evaluation per time value


@pytest.mark.gpu
@pytest.mark.parametrize("wave_num", [1, 7])
def test_demo_sphere_regular_waves_golden(host_build, sphere_h5, tmp_path, wave_num):
    o = tmp_path / "reg.txt"
    duration = 90.0
    out = subprocess.run([os.path.join(host_build, "demo_sphere_waves"), sphere_h5, str(o), "regular", str(wave_num),
                          str(duration)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    t, z = _read_traj(o)
    gold = common.sphere_goldens()["reg%d_um" % wave_num][:z.size] * 1e-6
    n1, n2 = common.traj_norms(z, gold)
    assert n1 <= 1e-4 and n2 <= 0.02 and n2 <= 2.0e-6, (n1, n2)


@pytest.mark.gpu
def test_demo_sphere_irregular_waves_golden(host_build, sphere_h5, tmp_path):
    o = tmp_path / "irr.txt"
    out = subprocess.run([os.path.join(host_build, "demo_sphere_waves"), sphere_h5, str(o), "irregular", "75.0"],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    t, z = _read_traj(o)
    gold = common.sphere_goldens()["irreg_um"][:z.size] * 1e-6
    n1, n2 = common.traj_norms(z, gold)
    assert n1 <= 1e-4 and n2 <= 0.02 and n2 <= 2e-4, (n1, n2)


@pytest.mark.gpu
def test_eta_file_import_reproduces_the_irregular_run(host_build, sphere_h5, tmp_path):
    """IrregularWaveParams::eta_file_path_ (src/wave_types.cpp:451-453,480-500).  The irregular-wave demo
    dumps its free-surface series in the reference's "time : eta" format; a second run that imports that file through
    IrregularWaves must retrace the first run's heave trajectory exactly.  Parse errors carry the reference's messages."""
    demo = os.path.join(host_build, "demo_sphere_waves")
    o1, o2, eta = tmp_path / "irr.txt", tmp_path / "imp.txt", tmp_path / "eta.txt"
    out = subprocess.run([demo, sphere_h5, str(o1), "irregular", "12.0", str(eta)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    first = open(eta).readline()
    assert " : " in first
    out2 = subprocess.run([demo, sphere_h5, str(o2), "eta", str(eta), "12.0"], capture_output=True, text=True)
    assert out2.returncode == 0, out2.stdout + out2.stderr
    t1, z1 = _read_traj(o1)
    t2, z2 = _read_traj(o2)
    np.testing.assert_array_equal(t1, t2)
    np.testing.assert_array_equal(z1, z2)
    assert np.ptp(z1) > 1e-3
    # the wave force line printed at the end agrees too; the elevation helper has no spectrum to sum over
    assert out.stdout.split("elevation")[0] == out2.stdout.split("elevation")[0]
    bad = tmp_path / "bad.txt"
    bad.write_text("0.0 : 0.1\n0.015 ; 0.2\n")
    out3 = subprocess.run([demo, sphere_h5, str(o2), "eta", str(bad), "1.0"], capture_output=True, text=True)
    assert out3.returncode == 1 and "Could not parse line: 0.015 ; 0.2." in out3.stderr
    out4 = subprocess.run([demo, sphere_h5, str(o2), "eta", str(tmp_path / "missing.txt"), "1.0"], capture_output=True, text=True)
    assert out4.returncode == 1 and "Unable to open file at:" in out4.stderr


@pytest.mark.gpu
def test_b1_latency_probe_runs(host_build, sphere_h5):
    """The C++ latency probe of the drop-in case (hc_step at B = 1 through the C ABI, and the first ComponentFunc
    evaluation at a new ChTime through TestHydro): both paths evaluate the same forces."""
    out = subprocess.run([os.path.join(host_build, "bench_b1_latency"), sphere_h5, "1", "0.015", "40", "20"],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "C ABI hc_step, B = 1, D = 6: median" in out.stdout and "first call at a new time: median" in out.stdout


@pytest.mark.gpu
def test_api_surface_two_bodies(host_build, tmp_path):
    h5 = tmp_path / "rm3_like.h5"
    h5io.write_bemio(h5, synth.rm3_like(rirf_steps=201, rirf_duration=10.0, exc_irf_steps=201, exc_half_window=5.0))
    y = tmp_path / "rm3.hydro.yaml"
    y.write_text("hydrodynamics:\n  bodies:\n    - name: body1\n      h5_file: rm3_like.h5\n    - name: body2\n"
                 "      h5_file: rm3_like.h5\n  waves:\n    type: irregular\n    height: 2.5\n    period: 8.0\n    seed: 5\n")
    out = subprocess.run([os.path.join(host_build, "test_api_surface"), str(h5), str(y)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "API surface test passed" in out.stdout


@pytest.mark.gpu
def test_iea_sphere_yaml_run_and_results_file(host_build, sphere_h5, tmp_path):
    """The reference's CLI regression case iea_sphere/decay through ReadHydroYAML + SetupHydroFromYAML + TestHydro +
    SimulationExporter: results .h5 in schema v0.3, heave held against expected/results.still.h5 with the harness'
    own gate (RMS-relative <= 0.02)."""
    y = tmp_path / "iea_sphere_decay.hydro.yaml"
    y.write_text("hydrodynamics:\n  bodies:\n    - name: body1\n      h5_file: %s\n\n  waves:\n    type: still\n" % sphere_h5)
    out_h5 = tmp_path / "results.still.h5"
    out = subprocess.run([os.path.join(host_build, "demo_iea_sphere_yaml"), str(y), str(out_h5)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    t = h5io.read_f64(out_h5, "results/time/time")
    pos = h5io.read_f64(out_h5, "results/model/bodies/body1/position")
    assert t.shape == (4000,) and pos.shape == (4000, 3)
    g = common.sphere_goldens()
    err = common.rms_relative_error(g["iea_decay_z"], np.interp(g["iea_decay_t"], t, pos[:, 2]))
    assert err <= 0.02 and err <= 2e-3, err
    assert h5io.list_group(out_h5, "/") == ["inputs", "meta", "results"]
    assert h5io.list_group(out_h5, "results/model/bodies") == ["body1", "ground"]
    assert set(h5io.list_group(out_h5, "results/model/bodies/body1")) == {
        "position", "velocity", "acceleration", "orientation", "orientation_xyz", "angular_velocity"}
    np.testing.assert_array_equal(h5io.read_f64(out_h5, "inputs/simulation/environment/gravity"), [0.0, 0.0, -9.8])
    np.testing.assert_array_equal(h5io.read_f64(out_h5, "inputs/model/bodies/body1/location"), [0.0, 0.0, -1.0])
    # independent reader agrees
    from h5lite import H5Lite
    np.testing.assert_array_equal(H5Lite(str(out_h5)).read("results/model/bodies/body1/position"), pos)
    # irregular waves: the exporter also records the spectrum and the free-surface elevation the device synthesised
    y.write_text("hydrodynamics:\n  bodies:\n    - name: body1\n      h5_file: %s\n\n  waves:\n    type: irregular\n"
                 "    height: 2.0\n    period: 12.0\n    seed: 3\n" % sphere_h5)
    out_h5 = tmp_path / "results.irregular.h5"
    out = subprocess.run([os.path.join(host_build, "demo_iea_sphere_yaml"), str(y), str(out_h5)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    f = h5io.read_f64(out_h5, "inputs/simulation/waves/irregular/frequencies_hz")
    S = h5io.read_f64(out_h5, "inputs/simulation/waves/irregular/spectral_densities")
    eta = h5io.read_f64(out_h5, "inputs/simulation/waves/irregular/free_surface_eta")
    et = h5io.read_f64(out_h5, "inputs/simulation/waves/irregular/free_surface_time")
    assert f.size == S.size == 40 and eta.size == et.size > 4000      # nf = ceil((1.0 - 0.001) * 40 s)
    assert np.abs(eta).max() > 0.1


@pytest.mark.gpu
def test_wave_kinematics_through_the_class_surface(host_build, sphere_h5, tmp_path):
    """RegularWave / IrregularWaves::GetElevation / GetVelocity / GetAcceleration of the C++ host layer
    (spectrum fetched from the device ensemble, Airy arithmetic behind hc_wave_kinematics) against the oracle's
    restatement of /root/reference/src/wave_types.cpp:14-160,301-313,515-550, Wheeler stretching off and on,
    mean water level 0 and != 0.  Bar: 1e-12 relative."""
    from oracle import hc_oracle as orc
    o = tmp_path / "kin.txt"
    out = subprocess.run([os.path.join(host_build, "test_kinematics"), sphere_h5, str(o)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    O = orc.Tables(common.sphere_raw())
    reg = orc.Instance(O)
    reg.set_regular(0.8, 0.9, 0.4)
    irr = orc.Instance(O)
    irr.set_irregular(dt=0.05, duration=20.0, ramp=0.0, Hs=2.0, Tp=9.0, fmin=0.02, fmax=0.6, nfreq=60, gamma=3.3, seed=7)
    cases = {"regular": (reg, False, 0.0), "regular_mwl": (reg, False, 0.3), "irregular": (irr, False, 0.0),
             "irregular_mwl": (irr, False, -0.2), "irregular_wheeler": (irr, True, 0.0),
             "irregular_wheeler_mwl": (irr, True, -0.2)}
    seen = {k: 0 for k in cases}
    for line in open(o):
        tag, *vals = line.split()
        v = np.array([float(x) for x in vals])
        inst, stretch, mwl = cases[tag]
        eta, vel, acc = inst.kinematics(v[0:3], v[3], wave_stretching=stretch, mwl=mwl)
        ref = np.concatenate([[eta], vel, acc])
        scale = np.abs(ref).max()
        assert np.all(np.abs(v[4:] - ref) <= 1e-12 * scale), (tag, v, ref)
        seen[tag] += 1
    assert all(n == 20 for n in seen.values()), seen


@pytest.mark.gpu
def test_demo_rm3_reg_waves_constrained_two_body(host_build, tmp_path):
    """BASELINE config 2 / SURVEY f3: RM3 float + spar plate on a prismatic joint between two MOVING bodies with a
    linear PTO damper (demos/yaml/rm3/rm3_linearPTO.model.yaml: 1.2e6 N s/m), HHT at dt = 0.01, regular waves
    A = 1.0 m / omega = 2.10 rad/s, through the C++ TestHydro on the GPU
    (/root/reference/demos/rm3/demo_rm3_reg_waves.cpp:63,97-150).  Every evaluation TestHydro made (time, pose,
    velocity, gravity -> total 12-DoF force) is replayed through the oracle: per-step force parity at 1e-9 on the
    same states.  rm3.h5 is stripped from the reference snapshot, so the tables are the synthetic RM3-shaped ones and
    the trajectory itself is only checked for physical sanity and for the joint holding."""
    from oracle import hc_oracle as orc
    raw = synth.rm3_like()
    h5 = tmp_path / "rm3_like.h5"
    h5io.write_bemio(h5, raw)
    o, trace = tmp_path / "rm3_reg_waves.txt", tmp_path / "trace.txt"
    env = dict(os.environ, HYDROC_STATE_TRACE=str(trace))
    res = tmp_path / "results.regular.h5"
    out = subprocess.run([os.path.join(host_build, "demo_rm3_reg_waves"), str(h5), str(o), "12.0", "1200000", str(res)],
                         capture_output=True, text=True, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    a = np.loadtxt(o, skiprows=1)
    t, zf, zp, xf = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
    assert t.size == 1201 and abs(t[-1] - 12.01) < 1e-9
    words = out.stdout.split()
    assert int(words[words.index("radiation_calls") + 1]) == t.size + 1     # one evaluation per time value (+ t = 0)
    assert float(words[words.index("joint_transverse") + 1]) < 1e-10 and float(words[words.index("joint_rel_rot") + 1]) < 1e-12
    # sanity: both bodies heave at the wave frequency around their equilibrium with a bounded amplitude; the PTO
    # couples them (the relative heave is smaller than it would be for free bodies is not asserted -- synthetic tables)
    late = t > 6.0
    assert 0.01 < np.ptp(zf[late]) < 6.0 and 0.001 < np.ptp(zp[late]) < 6.0
    assert abs(zf[late].mean() + 0.72) < 1.0 and abs(zp[late].mean() + 21.29) < 1.0 and np.all(np.isfinite(a))
    # per-step force parity on the same states
    tr = np.loadtxt(trace)
    assert tr.shape == (t.size + 1, 1 + 12 + 12 + 3 + 12)
    O = orc.Tables(raw)
    inst = orc.Instance(O)
    inst.set_regular(1.0, 2.10)
    ref = np.array([inst.force(r[0], r[1:13], r[13:25], r[25:28]) for r in tr])
    got = tr[:, 28:40]
    tol = common.force_tol(ref)
    assert np.all(np.abs(got - ref) <= tol), float((np.abs(got - ref) / tol).max())
    # results file: joint / TSDA channels of the reference's schema (src/simulation_exporter.cpp:67-69,120-152,303-353)
    from h5lite import H5Lite
    H = H5Lite(str(res))
    assert H.read("inputs/model/joints/names") == ["joint_1"] and H.read("inputs/model/tsdas/names") == ["TSDA_1"]
    assert H.read("inputs/model/rsdas/names") == []
    assert h5io.list_group(res, "results/model/tsdas") == ["TSDA_1"] and h5io.list_group(res, "results/model/joints") == ["joint_1"]
    speed = h5io.read_f64(res, "results/model/tsdas/TSDA_1/speed")
    damp = h5io.read_f64(res, "results/model/tsdas/TSDA_1/damping_force")
    fmag = h5io.read_f64(res, "results/model/tsdas/TSDA_1/force_mag")
    ext = h5io.read_f64(res, "results/model/tsdas/TSDA_1/extension")
    assert speed.shape == (t.size,) and np.abs(speed).max() > 1e-3
    np.testing.assert_array_equal(damp, 1200000.0 * speed)
    np.testing.assert_allclose(fmag, -damp, rtol=1e-9, atol=1e-3)     # no spring: ChLinkTSDA::GetForce = -(k ext + c d(len)/dt)
    p1 = h5io.read_f64(res, "results/model/bodies/body1/position")
    p2 = h5io.read_f64(res, "results/model/bodies/body2/position")
    np.testing.assert_allclose(ext, np.linalg.norm(p1 - p2, axis=1) - (21.29 - 0.72), atol=1e-9)   # length - free length
    f1 = h5io.read_f64(res, "results/model/joints/joint_1/reaction1_force")
    f2 = h5io.read_f64(res, "results/model/joints/joint_1/reaction2_force")
    assert f1.shape == (t.size, 3) and np.abs(f1[:, 0]).max() > 1e3           # the joint carries the float's surge load
    np.testing.assert_array_equal(f2, -f1)
    q = h5io.read_f64(res, "results/model/bodies/body2/orientation")         # wxyz; joint axis = the spar's z axis
    axis = np.stack([2 * (q[:, 1] * q[:, 3] + q[:, 0] * q[:, 2]), 2 * (q[:, 2] * q[:, 3] - q[:, 0] * q[:, 1]),
                     1 - 2 * (q[:, 1] ** 2 + q[:, 2] ** 2)], axis=1)
    along = np.abs((f1 * axis).sum(axis=1))
    assert along.max() <= 1e-6 * np.abs(f1).max()                             # a prismatic joint transmits nothing along its axis
    # the joint makes the 12 DoF truly coupled: both bodies carry the same angular velocity in every evaluation
    np.testing.assert_allclose(tr[:, 13 + 3:13 + 6], tr[:, 13 + 9:13 + 12], rtol=0, atol=1e-14)
    assert np.abs(tr[:, 13 + 4]).max() > 1e-6                               # ... and it is not trivially zero (pitch)


@pytest.mark.gpu
@pytest.mark.parametrize("stepper_name,wave", [("euler", "regular"), ("hht", "irregular")])
def test_yaml_period_sweep_maps_onto_ensemble_instances(host_build, sphere_h5, tmp_path, stepper_name, wave):
    """waves.period.values / linspace (parsed by the reference, /root/reference/src/hydro_yaml_parser.cpp:441-524,
    but never consumed) become the instances of ONE batched device ensemble (SetupHydroSweepFromYAML + TestHydroEnsemble):
    every instance must reproduce the single-system TestHydro run of its sweep point, the batch takes one device
    evaluation per time value, and (regular waves, Euler) the heave of every instance matches the oracle stepped with
    the reference's integrator."""
    from oracle import hc_oracle as orc
    import stepper
    y = tmp_path / "sweep.hydro.yaml"
    if wave == "regular":
        y.write_text("hydrodynamics:\n  bodies:\n    - name: body1\n      h5_file: %s\n  waves:\n    type: regular\n"
                     "    height: 0.5\n    period:\n      values: [3.0, 4.4, 6.0]\n" % sphere_h5)
        seeds = 1
    else:
        y.write_text("hydrodynamics:\n  bodies:\n    - name: body1\n      h5_file: %s\n  waves:\n    type: irregular\n"
                     "    height: 1.5\n    seed: 4\n    period:\n      linspace: { start: 8.0, stop: 12.0, num: 3 }\n" % sphere_h5)
        seeds = 2
    o = tmp_path / "sweep.txt"
    duration = 12.0
    out = subprocess.run([os.path.join(host_build, "demo_sweep_yaml"), str(y), str(o), stepper_name, str(duration), str(seeds)],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    words = out.stdout.split()
    B, nsteps = 3 * seeds, int(round(duration / 0.015))
    assert int(words[words.index("instances") + 1]) == B and int(words[words.index("steps") + 1]) == nsteps
    # one batched evaluation per time value: Euler evaluates at t_0 .. t_{n-1}; HHT at t_0 (initial accelerations) + t_1 .. t_n
    assert int(words[words.index("device_evaluations") + 1]) == nsteps + (1 if stepper_name == "hht" else 0)
    assert float(words[words.index("max_abs_diff_vs_single_runs") + 1]) <= 1e-12
    assert float(words[words.index("added_mass_mv_batched_rel_diff") + 1]) <= 1e-13      # k_added_mass_mv vs the host load
    a = np.loadtxt(o)
    assert a.shape == (nsteps, 1 + B)
    assert np.abs(a[:, 1:] + 2.0).max() < 1.5 and np.ptp(a[:, 1]) > 1e-3
    # the instances differ (different periods / seeds)
    assert np.abs(a[:, 1] - a[:, 2]).max() > 1e-4 and np.abs(a[:, 1] - a[:, -1]).max() > 1e-4
    if wave == "regular":
        O = orc.Tables(common.sphere_raw())
        for k, T in enumerate([3.0, 4.4, 6.0]):
            inst = orc.Instance(O)
            inst.set_regular(0.25, 2.0 * np.pi / T)
            free = np.zeros(6, bool); free[2] = True
            _, x = stepper.run(lambda t_, x_, v_: inst.force(t_, x_, v_), O.added_mass(), [common.SPHERE_MASS], [[1.0, 1.0, 1.0]],
                               [0, 0, -2.0, 0, 0, 0], 0.015, nsteps, free=free)
            assert np.abs(x[:, 2] - a[:, 1 + k]).max() < 1e-8, (k, np.abs(x[:, 2] - a[:, 1 + k]).max())
