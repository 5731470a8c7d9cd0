"""HDF5 writer/reader (no libhdf5): the C++ writer's files are read back by the C++ reader AND by the independent
pure-Python reader of tests/h5lite.py; a BEMIO-layout file written from the sphere fixture reproduces its tables."""
import os

import numpy as np
import pytest

import common
import hydrochrono_b200 as hc
from hydrochrono_b200 import h5io, synth
from h5lite import H5Lite, load_bemio
from oracle import hc_oracle as orc


def test_writer_roundtrip_generic(tmp_path):
    f = tmp_path / "out.h5"
    rng = np.random.default_rng(0)
    pos = rng.standard_normal((4000, 3))
    t = np.arange(4000) * 0.01
    w = h5io.H5Writer()
    w.put("results/time/time", t)
    w.put("results/model/bodies/body1/position", pos)
    w.put("results/model/bodies/body1/orientation", rng.standard_normal((4000, 4)))
    w.put("meta/run/tag", "decay test")
    w.put("inputs/simulation/environment/gravity", np.array([0.0, 0.0, -9.8]))
    w.attr("results/time/time", "units", "s")
    w.attr("results/model/bodies/body1", "mass", 261.8e3)
    w.group("results/model/rsdas")                        # empty group
    for i in range(23):                                   # > 8 children: several symbol nodes in one B-tree node
        w.put("many/d%02d" % i, np.full((2, 2), float(i)))
    w.save(f)
    # C++ reader
    np.testing.assert_array_equal(h5io.read_f64(f, "results/model/bodies/body1/position"), pos)
    np.testing.assert_array_equal(h5io.read_f64(f, "results/time/time"), t)
    assert h5io.read_string(f, "meta/run/tag") == "decay test"
    assert h5io.list_group(f, "/") == ["inputs", "many", "meta", "results"]
    assert h5io.list_group(f, "results/model") == ["bodies", "rsdas"]
    assert h5io.list_group(f, "results/model/rsdas") == []
    assert len(h5io.list_group(f, "many")) == 23
    np.testing.assert_array_equal(h5io.read_f64(f, "many/d17"), np.full((2, 2), 17.0))
    # independent Python reader
    p = H5Lite(str(f))
    np.testing.assert_array_equal(p.read("results/model/bodies/body1/position"), pos)
    assert p.read("meta/run/tag") == "decay test"
    assert p.keys("many") == ["d%02d" % i for i in range(23)]
    assert p.keys("/") == ["inputs", "many", "meta", "results"]
    with pytest.raises(hc.HydroError):
        h5io.read_f64(f, "results/nope")


@pytest.mark.parametrize("which", ["sphere", "rm3", "deep"])
def test_bemio_file_roundtrip(tmp_path, which):
    raw = common.sphere_raw() if which == "sphere" else synth.rm3_like(rirf_steps=101, exc_irf_steps=81, num_freqs=40)
    if which == "deep":
        raw["water_depth"] = float("inf")                 # written as the string "infinite" (h5fileinfo.cpp:207-218)
    f = tmp_path / (which + ".h5")
    h5io.write_bemio(f, raw)
    N = len(raw["bodies"])
    T = hc.Tables.from_h5(f, N)
    O = orc.Tables(raw)
    np.testing.assert_array_equal(T.rirf(), O.rirf())
    np.testing.assert_array_equal(T.added_mass(), O.added_mass())
    assert T.water_depth == raw["water_depth"]
    back = load_bemio(str(f), N)                          # pure-Python reader on the C++ writer's file
    for b0, b1 in zip(raw["bodies"], back["bodies"]):
        np.testing.assert_array_equal(np.asarray(b0["rirf_K"]), b1["rirf_K"])
        np.testing.assert_array_equal(np.asarray(b0["exc_irf_f"]).reshape(b1["exc_irf_f"].shape), b1["exc_irf_f"])
    assert back["water_depth"] == raw["water_depth"]


def test_corrupt_files_are_reported_not_crashed(tmp_path):
    """a truncated or corrupted .h5 must come back as HC_ERR_IO (H5FileInfo's 'Unable to open/read HDF5
    hydro data file', src/h5fileinfo.cpp:172-181), never as an out-of-bounds read: every file-controlled offset is
    bounds-checked, message sizes are validated against their block, group B-trees have a depth / cycle guard."""
    import hydrochrono_b200 as hc
    from hydrochrono_b200 import synth
    raw = synth.make_tables(num_bodies=1, rirf_steps=11, rirf_duration=1.0, exc_irf_steps=11, exc_half_window=1.0, num_freqs=8)
    good = tmp_path / "good.h5"
    h5io.write_bemio(good, raw)
    data = good.read_bytes()
    hc.Tables.from_h5(str(good), 1).close()
    rng = np.random.default_rng(5)
    outcomes = {"ok": 0, "io": 0}
    trials = [data[:n] for n in (0, 7, 95, 200, len(data) // 3, len(data) // 2, len(data) - 9)]
    for _ in range(300):                              # random byte / word / qword smashes in the metadata
        b = bytearray(data)
        for _ in range(int(rng.integers(1, 6))):
            at = int(rng.integers(8, len(b) - 8))
            width = int(rng.choice([1, 2, 8]))
            b[at:at + width] = rng.integers(0, 256, size=width, dtype=np.uint8).tobytes()
        trials.append(bytes(b))
    tree = data.find(b"TREE")                          # a group B-tree node that points at itself
    if tree >= 0:
        b = bytearray(data)
        b[tree + 5] = 1                                # level 1: children are TREE nodes
        b[tree + 24 + 8:tree + 24 + 16] = int(tree).to_bytes(8, "little")
        trials.append(bytes(b))
    for i, blob in enumerate(trials):
        f = tmp_path / ("bad%d.h5" % i)
        f.write_bytes(blob)
        try:
            hc.Tables.from_h5(str(f), 1).close()
            outcomes["ok"] += 1                        # the smash hit payload bytes or something the reader ignores
        except hc.HydroError as e:
            assert e.status in (1, 2, 6), (i, e.status, str(e))       # invalid shape / out of range / I/O -- never a crash
            outcomes["io"] += 1
        os.remove(f)
    assert outcomes["io"] >= 20, outcomes
