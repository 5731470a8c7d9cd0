// Classic-HDF5 reader for BEMIO hydro files, no libhdf5.
//
// Replaces the HDF5 C++ calls in H5FileInfo::ReadH5Data / InitScalar / Init1D / Init2D / Init3D
// (src/h5fileinfo.cpp:27-91,183-298).  Supports what BEMIO files of the sphere.h5 vintage contain
// (SURVEY.md Appendix B): superblock v0, 8-byte offsets/lengths, symbol-table groups (v1 B-tree,
// SNOD, local heap), v1 object headers incl. continuation blocks, dataspace v1/v2, contiguous or
// compact layout v3, little-endian IEEE float64 / float32 datasets and fixed-length strings.
// Anything else (chunking, filters, new-style groups) is reported as HC_ERR_IO.
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>

#include "hc_internal.h"

namespace hc {

struct H5File {
    std::vector<uint8_t> buf;
    uint64_t root_header = 0;
    std::string path;

    [[noreturn]] void bad(const std::string& why) const {
        fail(HC_ERR_IO, "Unable to open/read HDF5 hydro data file: " + path + "\nHDF5 error: " + why);
    }
    void need(uint64_t off, uint64_t n) const {
        if (off + n > buf.size() || off + n < off) bad("truncated file / address out of range");
    }
    uint64_t u(uint64_t off, int n) const {
        need(off, n);
        uint64_t v = 0;
        for (int i = n - 1; i >= 0; --i) v = (v << 8) | buf[off + i];
        return v;
    }
    uint8_t byte(uint64_t off) const { need(off, 1); return buf[off]; }   // every file-controlled offset is checked

    explicit H5File(const char* p) : path(p) {
        FILE* f = std::fopen(p, "rb");
        if (!f) bad("cannot open file");
        std::fseek(f, 0, SEEK_END);
        long sz = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        buf.resize(sz > 0 ? size_t(sz) : 0);
        size_t got = buf.empty() ? 0 : std::fread(buf.data(), 1, buf.size(), f);
        std::fclose(f);
        if (got != buf.size()) bad("short read");
        static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
        if (buf.size() < 96 || std::memcmp(buf.data(), sig, 8) != 0) bad("not an HDF5 file");
        if (buf[8] != 0) bad("unsupported superblock version " + std::to_string(buf[8]));
        if (buf[13] != 8 || buf[14] != 8) bad("unsupported offset/length size");
        if (u(24, 8) != 0) bad("non-zero base address");
        root_header = u(24 + 32 + 8, 8);  // root symbol-table entry: link-name offset, then header address
    }

    struct Msg { uint16_t type; uint64_t off; uint16_t size; };
    std::vector<Msg> messages(uint64_t hdr) const {
        need(hdr, 16);
        if (buf[hdr] != 1) bad("unsupported object header version");
        const unsigned count = unsigned(u(hdr + 2, 2));
        std::vector<std::pair<uint64_t, uint64_t>> blocks{{hdr + 16, u(hdr + 8, 4)}};
        std::vector<Msg> out;
        for (size_t bi = 0; bi < blocks.size() && out.size() < count; ++bi) {
            if (bi > 64) bad("too many object-header continuation blocks");
            uint64_t p = blocks[bi].first;
            need(p, blocks[bi].second);                            // the whole block lies inside the file
            const uint64_t end = p + blocks[bi].second;
            while (p + 8 <= end && out.size() < count) {
                Msg m{uint16_t(u(p, 2)), p + 8, uint16_t(u(p + 2, 2))};
                if (m.off + m.size > end) bad("object-header message runs past its block");
                if (m.type == 0x10) {
                    if (m.size < 16) bad("short continuation message");
                    blocks.push_back({u(m.off, 8), u(m.off + 8, 8)});
                }
                out.push_back(m);
                p = m.off + m.size;
            }
        }
        return out;
    }

    void collect(uint64_t node, uint64_t heap_data, std::map<std::string, uint64_t>& out, int depth = 0) const {
        if (depth > 32) bad("group B-tree too deep (cyclic node?)");
        need(node, 24);
        if (std::memcmp(&buf[node], "TREE", 4) == 0) {
            const unsigned level = byte(node + 5);
            if (int(level) + depth > 32) bad("group B-tree too deep (cyclic node?)");
            const unsigned used = unsigned(u(node + 6, 2));
            uint64_t p = node + 24;
            for (unsigned i = 0; i < used; ++i, p += 16) {
                const uint64_t child = u(p + 8, 8);
                if (child == node) bad("self-referencing group B-tree node");
                collect(child, heap_data, out, depth + 1);   // level > 0: another TREE node; level 0: a SNOD
            }
        } else if (std::memcmp(&buf[node], "SNOD", 4) == 0) {
            const unsigned n = unsigned(u(node + 6, 2));
            uint64_t p = node + 8;
            for (unsigned i = 0; i < n; ++i, p += 40) {
                const uint64_t s = heap_data + u(p, 8);
                need(s, 1);
                const char* name = reinterpret_cast<const char*>(&buf[s]);
                out[std::string(name, strnlen(name, buf.size() - s))] = u(p + 8, 8);
            }
        } else {
            bad("bad group node signature");
        }
    }

    std::map<std::string, uint64_t> children(uint64_t hdr) const {
        for (const Msg& m : messages(hdr))
            if (m.type == 0x11) {
                const uint64_t btree = u(m.off, 8), heap = u(m.off + 8, 8);
                need(heap, 32);
                if (std::memcmp(&buf[heap], "HEAP", 4) != 0) bad("bad local heap");
                std::map<std::string, uint64_t> out;
                collect(btree, u(heap + 24, 8), out);
                return out;
            }
        bad("object is not an old-style group");
    }

    uint64_t resolve(const std::string& p) const {
        uint64_t hdr = root_header;
        size_t i = 0;
        while (i < p.size()) {
            size_t j = p.find('/', i);
            if (j == std::string::npos) j = p.size();
            if (j > i) {
                auto kids = children(hdr);
                auto it = kids.find(p.substr(i, j - i));
                if (it == kids.end()) bad("object '" + p + "' not found");
                hdr = it->second;
            }
            i = j + 1;
        }
        return hdr;
    }

    struct Dataset {
        std::vector<uint64_t> dims;
        int type_class = -1;
        unsigned type_size = 0;
        uint64_t data_off = 0, data_size = 0;
    };
    Dataset dataset(const std::string& p) const {
        Dataset d;
        bool have_space = false, have_layout = false;
        for (const Msg& m : messages(resolve(p))) {
            if (m.type == 0x01) {
                const int ver = byte(m.off), rank = byte(m.off + 1);
                if (rank > 32) bad("dataspace rank out of range");
                uint64_t q = m.off + (ver == 1 ? 8 : 4);
                for (int r = 0; r < rank; ++r) d.dims.push_back(u(q + 8 * r, 8));
                have_space = true;
            } else if (m.type == 0x03) {
                d.type_class = byte(m.off) & 0x0f;
                d.type_size = unsigned(u(m.off + 4, 4));
                if (d.type_class == 1 && (byte(m.off + 1) & 1)) bad("big-endian floats unsupported");
            } else if (m.type == 0x08) {
                if (byte(m.off) != 3) bad("unsupported data layout version");
                const int cls = byte(m.off + 1);
                if (cls == 1) { d.data_off = u(m.off + 2, 8); d.data_size = u(m.off + 10, 8); }
                else if (cls == 0) { d.data_size = u(m.off + 2, 2); d.data_off = m.off + 4; }
                else bad("chunked datasets unsupported");
                have_layout = true;
            } else if (m.type == 0x0b) {
                bad("filtered datasets unsupported");
            }
        }
        if (!have_space || !have_layout || d.type_class < 0) bad("'" + p + "' is not a dataset");
        need(d.data_off, d.data_size);
        return d;
    }

    // numeric dataset flattened in row-major order (Init1D/2D/3D read into NATIVE_DOUBLE)
    dvec numbers(const std::string& p, std::vector<uint64_t>* dims = nullptr) const {
        Dataset d = dataset(p);
        if (d.type_class != 1 || (d.type_size != 8 && d.type_size != 4)) bad("'" + p + "' is not a float dataset");
        uint64_t n = 1;
        for (uint64_t x : d.dims) {                                 // checked multiply: dims come from the file
            if (x != 0 && n > (uint64_t(1) << 40) / x) bad("dataset '" + p + "' dimensions overflow");
            n *= x;
        }
        if (n * d.type_size > d.data_size) bad("dataset '" + p + "' storage too small");
        dvec out(n);
        for (uint64_t i = 0; i < n; ++i) {
            if (d.type_size == 8) { double v; std::memcpy(&v, &buf[d.data_off + 8 * i], 8); out[i] = v; }
            else { float v; std::memcpy(&v, &buf[d.data_off + 4 * i], 4); out[i] = v; }
        }
        if (dims) *dims = d.dims;
        return out;
    }

    // InitScalar (h5fileinfo.cpp:183-214): float/double scalar, or the string "infinite" -> +inf
    double scalar(const std::string& p) const {
        Dataset d = dataset(p);
        if (d.type_class == 3) {
            std::string s(reinterpret_cast<const char*>(&buf[d.data_off]), size_t(d.data_size));
            s = s.substr(0, s.find('\0'));
            if (s == "infinite") return std::numeric_limits<double>::infinity();
            return 0.0;  // reference leaves var untouched for other strings
        }
        dvec v = numbers(p);
        if (v.empty()) bad("empty scalar '" + p + "'");
        return v[0];
    }
};

hc_tables* load_bemio_h5(const char* path, int num_bodies) {
    if (num_bodies < 1) fail(HC_ERR_INVALID, "num_bodies must be >= 1");
    H5File f(path);
    hc_tables_desc d{};
    d.num_bodies = num_bodies;
    d.rho = f.scalar("simulation_parameters/rho");
    d.g = f.scalar("simulation_parameters/g");
    d.water_depth = f.scalar("simulation_parameters/water_depth");
    dvec w = f.numbers("simulation_parameters/w");
    const int N = num_bodies, D = 6 * N;
    dvec rirf_t, K, lin, ainf, vol, cg, cb, mag, ph, et, ef;
    int L = -1, Le0 = -1;
    const int nw = int(w.size());
    auto squeeze_mid = [&](const dvec& v, const std::vector<uint64_t>& dims, const std::string& name, int last) {
        // (6, ndir, n) -> direction 0 (SqueezeMid, h5fileinfo.cpp:168-181; regular waves use column j = 0)
        if (dims.size() != 3 || dims[0] != 6 || int(dims[2]) != last) f.bad("unexpected shape for " + name);
        dvec out(size_t(6) * last);
        for (int i = 0; i < 6; ++i)
            for (int k = 0; k < last; ++k) out[size_t(i) * last + k] = v[(size_t(i) * dims[1] + 0) * dims[2] + k];
        return out;
    };
    for (int b = 0; b < N; ++b) {
        const std::string body = "body" + std::to_string(b + 1);
        const std::string hcx = body + "/hydro_coeffs/";
        vol.push_back(f.scalar(body + "/properties/disp_vol"));
        dvec t = f.numbers(hcx + "radiation_damping/impulse_response_fun/t");
        if (L < 0) L = int(t.size());
        if (int(t.size()) != L) f.bad("RIRF time vectors differ in length between bodies");
        rirf_t.insert(rirf_t.end(), t.begin(), t.end());
        dvec v = f.numbers(body + "/properties/cg");
        if (v.size() < 3) f.bad("cg too short");
        cg.insert(cg.end(), v.begin(), v.begin() + 3);
        v = f.numbers(body + "/properties/cb");
        if (v.size() < 3) f.bad("cb too short");
        cb.insert(cb.end(), v.begin(), v.begin() + 3);
        std::vector<uint64_t> dims;
        v = f.numbers(hcx + "linear_restoring_stiffness", &dims);
        if (v.size() != 36) f.bad("linear_restoring_stiffness must be 6x6");
        lin.insert(lin.end(), v.begin(), v.end());
        v = f.numbers(hcx + "added_mass/inf_freq", &dims);
        if (dims.size() != 2 || dims[0] != 6 || int(dims[1]) != D)
            f.bad("added_mass/inf_freq must be 6 x 6*num_bodies (file has a different body count?)");
        ainf.insert(ainf.end(), v.begin(), v.end());
        v = f.numbers(hcx + "radiation_damping/impulse_response_fun/K", &dims);
        if (dims.size() != 3 || dims[0] != 6 || int(dims[1]) != D || int(dims[2]) != L)
            f.bad("impulse_response_fun/K must be 6 x 6*num_bodies x len(t)");
        K.insert(K.end(), v.begin(), v.end());
        v = f.numbers(hcx + "excitation/mag", &dims);
        dvec sq = squeeze_mid(v, dims, "excitation/mag", nw);
        mag.insert(mag.end(), sq.begin(), sq.end());
        v = f.numbers(hcx + "excitation/phase", &dims);
        sq = squeeze_mid(v, dims, "excitation/phase", nw);
        ph.insert(ph.end(), sq.begin(), sq.end());
        t = f.numbers(hcx + "excitation/impulse_response_fun/t");
        if (Le0 < 0) Le0 = int(t.size());
        if (int(t.size()) != Le0) f.bad("excitation IRF time vectors differ in length between bodies");
        et.insert(et.end(), t.begin(), t.end());
        v = f.numbers(hcx + "excitation/impulse_response_fun/f", &dims);
        sq = squeeze_mid(v, dims, "excitation/impulse_response_fun/f", Le0);
        ef.insert(ef.end(), sq.begin(), sq.end());
    }
    d.rirf_steps = L; d.num_freqs = nw; d.exc_irf_steps = Le0;
    d.rirf_t = rirf_t.data(); d.rirf_K = K.data(); d.lin_matrix = lin.data(); d.inf_added_mass = ainf.data();
    d.disp_vol = vol.data(); d.cg = cg.data(); d.cb = cb.data(); d.w = w.data(); d.exc_mag = mag.data();
    d.exc_phase = ph.data(); d.exc_irf_t = et.data(); d.exc_irf_f = ef.data();
    return tables_from_desc(d);
}

}  // namespace hc

// ---- generic dataset access (results files, fixtures) ----------------------------------------------
extern "C" {

hc_status hc_h5_read_f64(const char* file, const char* dataset, int* rank, uint64_t* dims, double* out, size_t capacity) {
    try {
        hc::H5File f(file);
        std::vector<uint64_t> d;
        hc::dvec v = f.numbers(dataset, &d);
        if (rank) *rank = int(d.size());
        if (dims) for (size_t i = 0; i < d.size() && i < 8; ++i) dims[i] = d[i];
        if (out) {
            if (capacity < v.size()) hc::fail(HC_ERR_INVALID, "hc_h5_read_f64: output buffer too small");
            std::copy(v.begin(), v.end(), out);
        }
        return HC_OK;
    } catch (const hc::StatusError& e) { hc::set_last_error(e.msg); return e.code; }
      catch (const std::exception& e) { hc::set_last_error(e.what()); return HC_ERR_IO; }
}

hc_status hc_h5_read_string(const char* file, const char* dataset, char* out, size_t capacity) {
    try {
        hc::H5File f(file);
        auto d = f.dataset(dataset);
        if (d.type_class != 3) hc::fail(HC_ERR_INVALID, "hc_h5_read_string: not a fixed-length string dataset");
        std::string s(reinterpret_cast<const char*>(&f.buf[d.data_off]), size_t(d.data_size));
        s = s.substr(0, s.find('\0'));
        if (capacity < s.size() + 1) hc::fail(HC_ERR_INVALID, "hc_h5_read_string: output buffer too small");
        std::memcpy(out, s.c_str(), s.size() + 1);
        return HC_OK;
    } catch (const hc::StatusError& e) { hc::set_last_error(e.msg); return e.code; }
      catch (const std::exception& e) { hc::set_last_error(e.what()); return HC_ERR_IO; }
}

// children of a group, '\n'-separated, sorted
hc_status hc_h5_list(const char* file, const char* group, char* out, size_t capacity) {
    try {
        hc::H5File f(file);
        std::string s;
        for (auto& kv : f.children(f.resolve(group))) { if (!s.empty()) s += "\n"; s += kv.first; }
        if (capacity < s.size() + 1) hc::fail(HC_ERR_INVALID, "hc_h5_list: output buffer too small");
        std::memcpy(out, s.c_str(), s.size() + 1);
        return HC_OK;
    } catch (const hc::StatusError& e) { hc::set_last_error(e.msg); return e.code; }
      catch (const std::exception& e) { hc::set_last_error(e.what()); return HC_ERR_IO; }
}

}  // extern "C"
