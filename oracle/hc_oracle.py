"""ctypes binding of the CPU oracle (oracle/hc_oracle.cpp).  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package hydrochrono_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force=False):
    so = os.path.join(_HERE, "libhc_oracle.so")
    src = os.path.join(_HERE, "hc_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libhc_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_last_error.restype = C.c_char_p
        L.orc_tables_create.restype = C.c_void_p
        L.orc_tables_create.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_double, C.c_double, C.c_double, _dp, _dp, _dp,
                                        _dp, _dp, C.c_int, _dp, _dp, _dp, C.c_int, _dp, _dp]
        L.orc_tables_destroy.argtypes = [C.c_void_p]
        L.orc_tables_set_tapered.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                             C.c_double]
        L.orc_tables_get_rirf.argtypes = [C.c_void_p, _dp]
        L.orc_tables_get_rirf_width.argtypes = [C.c_void_p, _dp]
        L.orc_added_mass.argtypes = [C.c_void_p, C.c_int, _dp]
        L.orc_added_mass_mv.argtypes = [C.c_void_p, C.c_int, C.c_double, _dp, _dp]
        L.orc_instance_create.restype = C.c_void_p
        L.orc_instance_create.argtypes = [C.c_void_p, C.c_int]
        L.orc_instance_destroy.argtypes = [C.c_void_p]
        L.orc_set_nowave.argtypes = [C.c_void_p]
        L.orc_set_regular.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.orc_set_irregular.argtypes = [C.c_void_p] + [C.c_double] * 9 + [C.c_int, C.c_int, C.c_void_p]
        L.orc_set_irregular_series.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_irregular_sizes.argtypes = [C.c_void_p, _ip, _ip, _ip]
        L.orc_irregular_get.argtypes = [C.c_void_p] + [_dp] * 7
        L.orc_irregular_get_irf.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
        L.orc_regular_get.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.orc_force.argtypes = [C.c_void_p, C.c_double] + [_dp] * 7
        L.orc_history_len.argtypes = [C.c_void_p]
        L.orc_profile.argtypes = [C.c_void_p, _dp]
        L.orc_wave_number.restype = C.c_double
        L.orc_wave_number.argtypes = [C.c_double] * 3
        L.orc_jonswap.argtypes = [C.c_int, _dp, C.c_double, C.c_double, C.c_double, C.c_int, _dp]
        L.orc_pierson_moskowitz.argtypes = [C.c_int, _dp, C.c_double, C.c_double, _dp]
        L.orc_linspaced.argtypes = [C.c_int, C.c_double, C.c_double, _dp]
        L.orc_phases.argtypes = [C.c_int, C.c_int, _dp]
        L.orc_phases_stdlib.argtypes = [C.c_int, C.c_int, _dp]
        L.orc_spline_resample.argtypes = [C.c_int, C.c_int, _dp, C.c_int, _dp]
        L.orc_get_lower_index.restype = C.c_long
        L.orc_get_lower_index.argtypes = [C.c_double, C.c_int, _dp]
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_bench_steps.restype = C.c_double
        L.orc_bench_steps.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, _dp,
                                      _dp, _dp, _dp]
        L.orc_wave_kinematics.argtypes = [C.c_void_p, _dp, C.c_double, C.c_int, C.c_double, _dp, _dp, _dp]
        L.orc_set_history.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        L.orc_bench_lockstep.restype = C.c_double
        L.orc_bench_lockstep.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, _dp, C.c_int, C.c_int, _dp, _dp, C.c_int,
                                         _dp, _dp, _dp]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class OracleError(RuntimeError):
    pass


def _err():
    return lib().orc_last_error().decode()


class Tables:
    """HydroData after H5FileInfo::ReadH5Data.  `raw` is the dict of tests/h5lite.load_bemio or
    hydrochrono_b200.synth.make_tables (raw, unscaled file values)."""

    def __init__(self, raw):
        L = lib()
        bodies = raw["bodies"]
        N = len(bodies)
        self.N, self.D = N, 6 * N
        self.L = bodies[0]["rirf_t"].size
        rirf_t = _c(np.stack([b["rirf_t"] for b in bodies]))
        K = _c(np.stack([b["rirf_K"] for b in bodies]))
        assert K.shape == (N, 6, 6 * N, self.L), K.shape
        lin = _c(np.stack([b["lin_matrix"] for b in bodies]))
        ainf = _c(np.stack([b["inf_added_mass"] for b in bodies]))
        assert ainf.shape == (N, 6, 6 * N)
        dv = _c([b["disp_vol"] for b in bodies])
        cg = _c(np.stack([b["cg"] for b in bodies]))
        cb = _c(np.stack([b["cb"] for b in bodies]))
        w = _c(raw["w"])
        self.nw = w.size
        mag = _c(np.stack([np.asarray(b["exc_mag"]).reshape(6, -1, self.nw)[:, 0, :] for b in bodies]))
        ph = _c(np.stack([np.asarray(b["exc_phase"]).reshape(6, -1, self.nw)[:, 0, :] for b in bodies]))
        self.Le0 = bodies[0]["exc_irf_t"].size
        et = _c(np.stack([b["exc_irf_t"] for b in bodies]))
        ef = _c(np.stack([np.asarray(b["exc_irf_f"]).reshape(6, -1, self.Le0)[:, 0, :] for b in bodies]))
        self.h = L.orc_tables_create(N, self.L, _p(rirf_t), _p(K), raw["rho"], raw["g"], raw["water_depth"], _p(lin),
                                     _p(ainf), _p(dv), _p(cg), _p(cb), self.nw, _p(w), _p(mag), _p(ph), self.Le0,
                                     _p(et), _p(ef))
        if not self.h:
            raise OracleError(_err())

    def set_tapered(self, smoothing="sg", window_length=5, rirf_end_time=-1.0, start=0.8, end=1.0, final=0.0):
        if lib().orc_tables_set_tapered(self.h, 1 if smoothing == "moving_average" else 0, window_length,
                                        rirf_end_time, start, end, final):
            raise OracleError(_err())

    def rirf(self):
        out = np.empty((self.D, self.D, self.L))
        lib().orc_tables_get_rirf(self.h, _p(out))
        return out

    def rirf_width(self):
        out = np.empty(self.L)
        lib().orc_tables_get_rirf_width(self.h, _p(out))
        return out

    def added_mass(self, n_sys=None):
        n = n_sys or self.D
        M = np.empty((n, n))
        if lib().orc_added_mass(self.h, n, _p(M)):
            raise OracleError(_err())
        return M

    def added_mass_mv(self, c, w, R):
        R = _c(R).copy()
        w = _c(w)
        if lib().orc_added_mass_mv(self.h, w.size, c, _p(w), _p(R)):
            raise OracleError(_err())
        return R

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_tables_destroy(self.h)
            self.h = None


class Instance:
    """One TestHydro (one system instance)."""

    def __init__(self, tables, omp_mode=1):
        self.T = tables
        self.h = lib().orc_instance_create(tables.h, omp_mode)
        self.D = tables.D

    def set_nowave(self):
        lib().orc_set_nowave(self.h)

    def set_regular(self, amplitude, omega, phase=0.0):
        rc = lib().orc_set_regular(self.h, amplitude, omega, phase)
        if rc:
            raise (IndexError if rc == -2 else OracleError)(_err())

    def set_irregular(self, dt, duration, ramp=0.0, Hs=0.0, Tp=0.0, fmin=0.001, fmax=1.0, nfreq=0, gamma=1.0,
                      is_normalized=False, seed=1, share_irf_from=None):
        rc = lib().orc_set_irregular(self.h, dt, duration, ramp, Hs, Tp, fmin, fmax, float(nfreq), gamma,
                                     int(is_normalized), seed, share_irf_from.h if share_irf_from else None)
        if rc:
            raise OracleError(_err())

    def set_irregular_series(self, dt, time, eta, share_irf_from=None):
        """Imported free-surface series (wave_types.cpp:480-500): the excitation convolution runs over (time, eta)."""
        time = np.ascontiguousarray(time, dtype=np.float64); eta = np.ascontiguousarray(eta, dtype=np.float64)
        assert time.shape == eta.shape and time.ndim == 1
        rc = lib().orc_set_irregular_series(self.h, dt, len(time), _p(time), _p(eta),
                                            share_irf_from.h if share_irf_from else None)
        if rc:
            raise OracleError(_err())

    def irregular(self):
        nf, ne = C.c_int(), C.c_int()
        Le = (C.c_int * self.T.N)()
        if lib().orc_irregular_sizes(self.h, C.byref(nf), C.byref(ne), Le):
            raise OracleError("not irregular")
        nf, ne = nf.value, ne.value
        out = {k: np.empty(nf) for k in ("freqs", "S", "widths", "phases", "wavenumbers")}
        out["eta_t"] = np.empty(ne)
        out["eta"] = np.empty(ne)
        lib().orc_irregular_get(self.h, *[_p(out[k]) for k in ("freqs", "S", "widths", "phases", "wavenumbers",
                                                                "eta_t", "eta")])
        out["irf"] = []
        for b in range(self.T.N):
            n = Le[b]
            t, w, f = np.empty(n), np.empty(n), np.empty((6, n))
            lib().orc_irregular_get_irf(self.h, b, _p(t), _p(w), _p(f))
            out["irf"].append({"t": t, "w": w, "f": f})
        return out

    def regular(self):
        mag, ph = np.empty(self.D), np.empty(self.D)
        k = C.c_double()
        lib().orc_regular_get(self.h, _p(mag), _p(ph), C.byref(k))
        return mag, ph, k.value

    def force(self, t, pose, vel, gvec=(0.0, 0.0, -9.81), components=False):
        pose, vel, g = _c(pose), _c(vel), _c(gvec)
        tot = np.empty(self.D)
        hs = rad = wv = None
        if components:
            hs, rad, wv = np.empty(self.D), np.empty(self.D), np.empty(self.D)
        rc = lib().orc_force(self.h, t, _p(pose), _p(vel), _p(g), _p(tot), _p(hs), _p(rad), _p(wv))
        if rc < 0:
            raise (IndexError if rc == -2 else OracleError)(_err())
        if components:
            return tot, hs, rad, wv
        return tot

    def history_len(self):
        return lib().orc_history_len(self.h)

    def kinematics(self, position, t, wave_stretching=True, mwl=0.0):
        """(eta, velocity[3], acceleration[3]) of the wave object at a point: WaveBase::GetElevation / GetVelocity /
        GetAcceleration (src/wave_types.cpp:301-313,515-550)."""
        p = _c(position)
        eta = C.c_double()
        v, a = np.empty(3), np.empty(3)
        if lib().orc_wave_kinematics(self.h, _p(p), t, int(wave_stretching), mwl, C.byref(eta), _p(v), _p(a)):
            raise OracleError(_err())
        return eta.value, v, a

    def set_history(self, times_newest_first, vel):
        """Loads a velocity history (times newest first, vel[n][D]) as if the instance had been stepped through it."""
        t, v = _c(times_newest_first), _c(vel)
        assert v.shape == (t.size, self.D)
        if lib().orc_set_history(self.h, t.size, _p(t), _p(v)):
            raise OracleError(_err())

    def profile(self):
        s = np.empty(3)
        lib().orc_profile(self.h, _p(s))
        return {"hydrostatics_seconds": s[0], "radiation_seconds": s[1], "waves_seconds": s[2]}

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_instance_destroy(self.h)
            self.h = None


def wave_number(omega, depth, g):
    v = lib().orc_wave_number(omega, depth, g)
    if np.isnan(v):
        raise OracleError(_err())
    return v


def jonswap(f, Hs, Tp, gamma=3.3, is_normalized=False):
    f = _c(f)
    S = np.empty_like(f)
    lib().orc_jonswap(f.size, _p(f), Hs, Tp, gamma, int(is_normalized), _p(S))
    return S


def pierson_moskowitz(f, Hs, Tp):
    f = _c(f)
    S = np.empty_like(f)
    lib().orc_pierson_moskowitz(f.size, _p(f), Hs, Tp, _p(S))
    return S


def linspaced(n, lo, hi):
    out = np.empty(n)
    lib().orc_linspaced(n, lo, hi, _p(out))
    return out


def phases(seed, n, stdlib=False):
    out = np.empty(n)
    (lib().orc_phases_stdlib if stdlib else lib().orc_phases)(seed, n, _p(out))
    return out


def spline_resample(pts, n_new):
    pts = _c(pts)
    dim, n_old = pts.shape
    out = np.empty((dim, n_new))
    if lib().orc_spline_resample(dim, n_old, _p(pts), n_new, _p(out)):
        raise OracleError(_err())
    return out


def get_lower_index(value, ticks):
    ticks = _c(ticks)
    r = lib().orc_get_lower_index(value, ticks.size, _p(ticks))
    if r < 0:
        raise OracleError(_err())
    return r


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(n)


def bench_steps(instances, nsteps, t0, dt, mode, amp, om, gvec=(0.0, 0.0, -9.81)):
    """Times `nsteps` force evaluations over `instances` with the prescribed synthetic motion
    pose_d = amp_d*sin(om_d*t + 0.01*i).  Returns (seconds, checksum)."""
    arr = (C.c_void_p * len(instances))(*[i.h for i in instances])
    amp, om, g = _c(amp), _c(om), _c(gvec)
    cs = C.c_double()
    sec = lib().orc_bench_steps(arr, len(instances), nsteps, t0, dt, mode, _p(amp), _p(om), _p(g), C.byref(cs))
    return sec, cs.value


def bench_lockstep(instances, times, pose, vel, buf0=0, mode=1, gvec=(0.0, 0.0, -9.81), want_forces=False):
    """Steps `instances` through `times` with the GPU arm's inputs: step n takes pose/vel [nbuf][count][D] buffer
    (buf0 + n) % nbuf.  Returns (seconds, checksum, forces[nsteps][count][D] or None)."""
    arr = (C.c_void_p * len(instances))(*[i.h for i in instances])
    times, pose, vel, g = _c(times), _c(pose), _c(vel), _c(gvec)
    nbuf, count, D = pose.shape
    assert count == len(instances) and vel.shape == pose.shape
    F = np.empty((times.size, count, D)) if want_forces else None
    cs = C.c_double()
    sec = lib().orc_bench_lockstep(arr, count, times.size, _p(times), nbuf, buf0, _p(pose), _p(vel), mode, _p(g),
                                   _p(F), C.byref(cs))
    return sec, cs.value, F
