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
