"""Minimal pure-Python reader for classic HDF5 files (TEST INFRASTRUCTURE).

Independent of the product's C++ reader (hydrochrono_b200/csrc/hc_h5.cpp) so the two can be
cross-checked.  Handles exactly the structures BEMIO files of the sphere.h5 vintage use
(SURVEY.md Appendix B): superblock v0, symbol-table groups (v1 B-tree + SNOD + local heap),
v1 object headers with continuation blocks, contiguous little-endian float64 datasets and
fixed-length strings.  No filters, no chunking.
"""
import struct

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"


class H5Lite:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.buf = f.read()
        b = self.buf
        if b[:8] != _SIG:
            raise ValueError("not an HDF5 file")
        if b[8] != 0:
            raise ValueError("only superblock v0 supported")
        self.O = b[13]
        self.L = b[14]
        if self.O != 8 or self.L != 8:
            raise ValueError("only 8-byte offsets/lengths supported")
        base = self._u64(24)
        if base != 0:
            raise ValueError("non-zero base address")
        ste = 24 + 4 * 8
        # root symbol-table entry: name off, header addr, cache type, reserved, scratch(btree, heap)
        self.root_hdr = self._u64(ste + 8)
        self._group_cache = {}

    # -- primitive readers -------------------------------------------------
    def _u16(self, o):
        return struct.unpack_from("<H", self.buf, o)[0]

    def _u32(self, o):
        return struct.unpack_from("<I", self.buf, o)[0]

    def _u64(self, o):
        return struct.unpack_from("<Q", self.buf, o)[0]

    # -- object header -----------------------------------------------------
    def _messages(self, addr):
        b = self.buf
        if b[addr] != 1:
            raise ValueError("only v1 object headers supported (got %d)" % b[addr])
        nmsg = self._u16(addr + 2)
        hsize = self._u32(addr + 8)
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            off, size = blocks.pop(0)
            end = off + size
            while off + 8 <= end and len(out) < nmsg:
                mtype = self._u16(off)
                msize = self._u16(off + 2)
                data = off + 8
                if mtype == 0x10:
                    blocks.append((self._u64(data), self._u64(data + 8)))
                out.append((mtype, data, msize))
                off = data + msize
        return out

    # -- groups --------------------------------------------------------------
    def _group_entries(self, hdr_addr):
        if hdr_addr in self._group_cache:
            return self._group_cache[hdr_addr]
        btree = heap = None
        for mtype, data, _ in self._messages(hdr_addr):
            if mtype == 0x11:
                btree = self._u64(data)
                heap = self._u64(data + 8)
        if btree is None:
            raise KeyError("object is not a group")
        if self.buf[heap:heap + 4] != b"HEAP":
            raise ValueError("bad local heap")
        heap_data = self._u64(heap + 24)
        entries = {}
        self._walk_btree(btree, heap_data, entries)
        self._group_cache[hdr_addr] = entries
        return entries

    def _walk_btree(self, addr, heap_data, entries):
        b = self.buf
        if b[addr:addr + 4] != b"TREE":
            raise ValueError("bad B-tree node")
        level = b[addr + 5]
        used = self._u16(addr + 6)
        p = addr + 8 + 16  # skip siblings
        for i in range(used):
            child = self._u64(p + 8)  # key_i (8) then child_i (8)
            p += 16
            if level > 0:
                self._walk_btree(child, heap_data, entries)
            else:
                self._read_snod(child, heap_data, entries)

    def _read_snod(self, addr, heap_data, entries):
        b = self.buf
        if b[addr:addr + 4] != b"SNOD":
            raise ValueError("bad symbol node")
        n = self._u16(addr + 6)
        p = addr + 8
        for _ in range(n):
            name_off = self._u64(p)
            hdr = self._u64(p + 8)
            s = heap_data + name_off
            e = b.index(b"\0", s)
            entries[b[s:e].decode()] = hdr
            p += 40

    def _resolve(self, path):
        hdr = self.root_hdr
        for part in [p for p in path.split("/") if p]:
            ent = self._group_entries(hdr)
            if part not in ent:
                raise KeyError(path)
            hdr = ent[part]
        return hdr

    def keys(self, path="/"):
        return sorted(self._group_entries(self._resolve(path)))

    # -- datasets ------------------------------------------------------------
    def read(self, path):
        hdr = self._resolve(path)
        dims = None
        dt_class = dt_size = None
        layout = None
        for mtype, data, msize in self._messages(hdr):
            b = self.buf
            if mtype == 0x01:
                ver, rank, flags = b[data], b[data + 1], b[data + 2]
                off = data + (8 if ver == 1 else 4)
                dims = [self._u64(off + 8 * i) for i in range(rank)]
            elif mtype == 0x03:
                dt_class = b[data] & 0x0F
                dt_size = self._u32(data + 4)
            elif mtype == 0x08:
                ver, cls = b[data], b[data + 1]
                if ver != 3:
                    raise ValueError("layout version %d unsupported" % ver)
                if cls == 1:
                    layout = ("contig", self._u64(data + 2), self._u64(data + 10))
                elif cls == 0:
                    sz = self._u16(data + 2)
                    layout = ("compact", data + 4, sz)
                else:
                    raise ValueError("chunked layout unsupported")
        if dims is None or dt_class is None or layout is None:
            raise KeyError("%s is not a dataset" % path)
        _, addr, size = layout
        raw = self.buf[addr:addr + size]
        if dt_class == 1 and dt_size == 8:
            n = int(np.prod(dims)) if dims else 1
            return np.frombuffer(raw, dtype="<f8", count=n).reshape(dims).copy()
        if dt_class == 3:
            if dims:        # 1-D array of fixed-length strings
                return [raw[i * dt_size:(i + 1) * dt_size].split(b"\0")[0].decode() for i in range(int(dims[0]))]
            return raw.split(b"\0")[0].decode()
        raise ValueError("datatype class %d size %d unsupported" % (dt_class, dt_size))


def load_bemio(path, num_bodies):
    """Datasets HydroChrono reads (reference src/h5fileinfo.cpp:27-91), raw/unscaled."""
    h = H5Lite(path)
    sp = "simulation_parameters/"
    depth = h.read(sp + "water_depth")
    if isinstance(depth, str):
        depth = float("inf") if depth == "infinite" else float("nan")
    else:
        depth = float(depth.ravel()[0])
    out = {
        "rho": float(h.read(sp + "rho").ravel()[0]),
        "g": float(h.read(sp + "g").ravel()[0]),
        "water_depth": depth,
        "w": h.read(sp + "w").ravel(),
        "bodies": [],
    }
    for i in range(num_bodies):
        bn = "body%d/" % (i + 1)
        hc = bn + "hydro_coeffs/"
        out["bodies"].append({
            "disp_vol": float(h.read(bn + "properties/disp_vol").ravel()[0]),
            "cg": h.read(bn + "properties/cg").ravel(),
            "cb": h.read(bn + "properties/cb").ravel(),
            "lin_matrix": h.read(hc + "linear_restoring_stiffness"),
            "inf_added_mass": h.read(hc + "added_mass/inf_freq"),
            "rirf_K": h.read(hc + "radiation_damping/impulse_response_fun/K"),
            "rirf_t": h.read(hc + "radiation_damping/impulse_response_fun/t").ravel(),
            "exc_mag": h.read(hc + "excitation/mag"),
            "exc_phase": h.read(hc + "excitation/phase"),
            "exc_irf_f": h.read(hc + "excitation/impulse_response_fun/f"),
            "exc_irf_t": h.read(hc + "excitation/impulse_response_fun/t").ravel(),
        })
    return out
