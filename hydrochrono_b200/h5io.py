"""HDF5 files without libhdf5, through the library's classic-format writer/reader (C ABI hc_h5_*).

write_bemio() lays a raw table dict out exactly like a BEMIO hydro file (the datasets H5FileInfo reads,
reference src/h5fileinfo.cpp:27-91), so synthetic or fixture tables can be handed to code that takes a file name
(H5FileInfo, TestHydro(bodies, h5_file_name), hydro.yaml `h5_file:`).
"""
import ctypes as C

import numpy as np

from . import _check
from ._capi import lib, dp


class H5Writer:
    def __init__(self):
        h = C.c_void_p()
        _check(lib.hc_h5_writer_create(C.byref(h)))
        self._h = h

    def group(self, path):
        _check(lib.hc_h5_writer_put_group(self._h, path.encode()))

    def put(self, path, value):
        if isinstance(value, str):
            _check(lib.hc_h5_writer_put_string(self._h, path.encode(), value.encode()))
            return
        a = np.ascontiguousarray(value, dtype=np.float64)
        dims = (C.c_uint64 * max(a.ndim, 1))(*a.shape)
        _check(lib.hc_h5_writer_put_f64(self._h, path.encode(), a.ndim, dims, a.ctypes.data_as(dp)))

    def attr(self, path, name, value):
        if isinstance(value, str):
            _check(lib.hc_h5_writer_attr_string(self._h, path.encode(), name.encode(), value.encode()))
        else:
            _check(lib.hc_h5_writer_attr_f64(self._h, path.encode(), name.encode(), float(value)))

    def save(self, file):
        _check(lib.hc_h5_writer_save(self._h, str(file).encode()))

    def close(self):
        if getattr(self, "_h", None):
            lib.hc_h5_writer_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


def read_f64(file, dataset):
    rank = C.c_int()
    dims = (C.c_uint64 * 8)()
    _check(lib.hc_h5_read_f64(str(file).encode(), dataset.encode(), C.byref(rank), dims, None, 0))
    shape = tuple(int(dims[i]) for i in range(rank.value))
    out = np.empty(shape if shape else (1,))
    _check(lib.hc_h5_read_f64(str(file).encode(), dataset.encode(), C.byref(rank), dims, out.ctypes.data_as(dp), out.size))
    return out.reshape(shape)


def read_string(file, dataset):
    buf = C.create_string_buffer(1 << 16)
    _check(lib.hc_h5_read_string(str(file).encode(), dataset.encode(), buf, len(buf)))
    return buf.value.decode()


def list_group(file, group="/"):
    buf = C.create_string_buffer(1 << 16)
    _check(lib.hc_h5_list(str(file).encode(), group.encode(), buf, len(buf)))
    s = buf.value.decode()
    return s.split("\n") if s else []


def write_bemio(file, raw):
    """raw: dict layout of hydrochrono_b200.synth.make_tables / tests/h5lite.load_bemio."""
    w = H5Writer()
    sp = "simulation_parameters/"
    w.put(sp + "rho", np.array([[raw["rho"]]]))
    w.attr(sp + "rho", "units", "kg/m^3")
    w.put(sp + "g", np.array([[raw["g"]]]))
    if np.isinf(raw["water_depth"]):
        w.put(sp + "water_depth", "infinite")
    else:
        w.put(sp + "water_depth", np.array([[raw["water_depth"]]]))
    w.put(sp + "w", np.asarray(raw["w"]).reshape(-1, 1))
    w.put("bem_data/code", "hydrochrono_b200")
    for i, b in enumerate(raw["bodies"]):
        bn = "body%d/" % (i + 1)
        hcx = bn + "hydro_coeffs/"
        nw = np.asarray(raw["w"]).size
        le0 = np.asarray(b["exc_irf_t"]).size
        w.put(bn + "properties/name", "body%d" % (i + 1))
        w.put(bn + "properties/disp_vol", np.array([[b["disp_vol"]]]))
        w.put(bn + "properties/cg", np.asarray(b["cg"]).reshape(3, 1))
        w.put(bn + "properties/cb", np.asarray(b["cb"]).reshape(3, 1))
        w.put(hcx + "linear_restoring_stiffness", b["lin_matrix"])
        w.put(hcx + "added_mass/inf_freq", b["inf_added_mass"])
        w.put(hcx + "radiation_damping/impulse_response_fun/K", b["rirf_K"])
        w.put(hcx + "radiation_damping/impulse_response_fun/t", np.asarray(b["rirf_t"]).reshape(-1, 1))
        w.put(hcx + "excitation/mag", np.asarray(b["exc_mag"]).reshape(6, -1, nw))
        w.put(hcx + "excitation/phase", np.asarray(b["exc_phase"]).reshape(6, -1, nw))
        w.put(hcx + "excitation/impulse_response_fun/f", np.asarray(b["exc_irf_f"]).reshape(6, -1, le0))
        w.put(hcx + "excitation/impulse_response_fun/t", np.asarray(b["exc_irf_t"]).reshape(-1, 1))
    w.save(file)
    w.close()
