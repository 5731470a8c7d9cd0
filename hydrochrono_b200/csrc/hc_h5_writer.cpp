// Classic-HDF5 writer without libhdf5: superblock v0, old-style groups (v1 B-tree + symbol nodes + local heap),
// v1 object headers, contiguous little-endian float64 datasets, fixed-length string datasets and scalar
// attributes -- the structures HydroChrono's H5Writer / SimulationExporter output is made of
// (src/h5_writer.cpp, src/simulation_exporter.cpp:181-199,373-391) and the ones BEMIO input files use, so the
// reader in hc_h5.cpp (and any stock HDF5 tool) reads them back.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>

#include "hc_internal.h"

namespace hc {
namespace {

constexpr uint64_t UNDEF = ~uint64_t(0);
constexpr int LEAF_K = 4;        // symbols per SNOD: up to 2 * LEAF_K
constexpr int INTERNAL_K = 16;   // children per B-tree node: up to 2 * INTERNAL_K

struct Attr {
    std::string name;
    bool is_string = false;
    std::string s;
    double d = 0;
};

struct Node {
    bool is_group = true;
    std::map<std::string, std::unique_ptr<Node>> kids;   // sorted by name, as symbol nodes require
    // dataset payload
    bool is_string = false;
    size_t str_size = 0;             // fixed element size of a string ARRAY (0: scalar string, size = all bytes)
    std::vector<uint64_t> dims;
    std::vector<uint8_t> bytes;
    std::vector<Attr> attrs;
    // filled while serialising
    uint64_t header = 0, btree = 0, heap = 0;
};

struct Out {
    std::vector<uint8_t> b;
    uint64_t pos() const { return b.size(); }
    void align8() { while (b.size() % 8) b.push_back(0); }
    void u8(uint8_t v) { b.push_back(v); }
    void u16(uint16_t v) { for (int i = 0; i < 2; ++i) b.push_back(uint8_t(v >> (8 * i))); }
    void u32(uint32_t v) { for (int i = 0; i < 4; ++i) b.push_back(uint8_t(v >> (8 * i))); }
    void u64(uint64_t v) { for (int i = 0; i < 8; ++i) b.push_back(uint8_t(v >> (8 * i))); }
    void raw(const void* p, size_t n) { const uint8_t* q = static_cast<const uint8_t*>(p); b.insert(b.end(), q, q + n); }
    void zeros(size_t n) { b.insert(b.end(), n, 0); }
    void patch64(uint64_t at, uint64_t v) { for (int i = 0; i < 8; ++i) b[at + i] = uint8_t(v >> (8 * i)); }
};

size_t pad8(size_t n) { return (n + 7) & ~size_t(7); }

// ---- message bodies ----------------------------------------------------------------------------
std::vector<uint8_t> msg_dataspace(const std::vector<uint64_t>& dims) {
    Out o;
    o.u8(1); o.u8(uint8_t(dims.size())); o.u8(0); o.u8(0); o.u32(0);
    for (uint64_t d : dims) o.u64(d);
    return o.b;
}
std::vector<uint8_t> msg_type_f64() {
    Out o;
    o.u8(0x11); o.u8(0x20); o.u8(0x3f); o.u8(0x00); o.u32(8);          // class 1 (float) v1, LE, size 8
    o.u16(0); o.u16(64); o.u8(52); o.u8(11); o.u8(0); o.u8(52); o.u32(1023);
    return o.b;
}
std::vector<uint8_t> msg_type_str(size_t size) {
    Out o;
    o.u8(0x13); o.u8(0); o.u8(0); o.u8(0); o.u32(uint32_t(size));       // class 3 (string) v1, null-terminated ASCII
    return o.b;
}
std::vector<uint8_t> msg_fill() { return {2, 2, 2, 1, 0, 0, 0, 0}; }
std::vector<uint8_t> msg_layout(uint64_t addr, uint64_t size) {
    Out o;
    o.u8(3); o.u8(1); o.u64(addr); o.u64(size);
    return o.b;
}
std::vector<uint8_t> msg_attr(const Attr& a) {
    Out o;
    const std::vector<uint8_t> type = a.is_string ? msg_type_str(a.s.size() + 1) : msg_type_f64();
    const std::vector<uint8_t> space = msg_dataspace({});
    o.u8(1); o.u8(0); o.u16(uint16_t(a.name.size() + 1)); o.u16(uint16_t(type.size())); o.u16(uint16_t(space.size()));
    o.raw(a.name.c_str(), a.name.size() + 1); o.zeros(pad8(a.name.size() + 1) - (a.name.size() + 1));
    o.raw(type.data(), type.size()); o.zeros(pad8(type.size()) - type.size());
    o.raw(space.data(), space.size()); o.zeros(pad8(space.size()) - space.size());
    if (a.is_string) o.raw(a.s.c_str(), a.s.size() + 1); else o.raw(&a.d, 8);
    return o.b;
}

// v1 object header with the given messages; returns its address
uint64_t write_header(Out& o, const std::vector<std::pair<uint16_t, std::vector<uint8_t>>>& msgs) {
    o.align8();
    const uint64_t at = o.pos();
    size_t total = 0;
    for (auto& m : msgs) total += 8 + pad8(m.second.size());
    o.u8(1); o.u8(0); o.u16(uint16_t(msgs.size())); o.u32(1); o.u32(uint32_t(total)); o.u32(0);
    for (auto& m : msgs) {
        o.u16(m.first); o.u16(uint16_t(pad8(m.second.size()))); o.u8(0); o.u8(0); o.u8(0); o.u8(0);
        o.raw(m.second.data(), m.second.size());
        o.zeros(pad8(m.second.size()) - m.second.size());
    }
    return at;
}

void write_dataset(Out& o, Node& n) {
    o.align8();
    const uint64_t data_at = o.pos();
    o.raw(n.bytes.data(), n.bytes.size());
    std::vector<std::pair<uint16_t, std::vector<uint8_t>>> msgs;
    msgs.push_back({0x01, msg_dataspace(n.dims)});
    msgs.push_back({0x03, n.is_string ? msg_type_str(n.str_size ? n.str_size : n.bytes.size()) : msg_type_f64()});
    msgs.push_back({0x05, msg_fill()});
    msgs.push_back({0x08, msg_layout(n.bytes.empty() ? UNDEF : data_at, n.bytes.size())});
    for (const Attr& a : n.attrs) msgs.push_back({0x0c, msg_attr(a)});
    n.header = write_header(o, msgs);
}

void write_group(Out& o, Node& g) {
    for (auto& kv : g.kids) {
        if (kv.second->is_group) write_group(o, *kv.second);
        else write_dataset(o, *kv.second);
    }
    if (g.kids.size() > size_t(2 * LEAF_K) * size_t(2 * INTERNAL_K))
        fail(HC_ERR_INVALID, "h5 writer: too many children in one group");
    // local heap: "" at offset 0, then the child names, then one free block
    std::vector<uint64_t> name_off;
    Out heap;
    heap.zeros(8);
    for (auto& kv : g.kids) {
        name_off.push_back(heap.pos());
        heap.raw(kv.first.c_str(), kv.first.size() + 1);
        heap.align8();
    }
    const uint64_t free_off = heap.pos();
    heap.u64(1);    // H5HL_FREE_NULL: last free block
    heap.u64(16);   // its size
    o.align8();
    g.heap = o.pos();
    o.raw("HEAP", 4); o.u8(0); o.u8(0); o.u8(0); o.u8(0);
    o.u64(heap.pos()); o.u64(free_off); o.u64(g.heap + 32);
    o.raw(heap.b.data(), heap.b.size());
    // symbol nodes
    std::vector<uint64_t> snods, last_name;
    size_t idx = 0;
    auto it = g.kids.begin();
    while (it != g.kids.end() || snods.empty()) {
        o.align8();
        snods.push_back(o.pos());
        const size_t n = std::min<size_t>(2 * LEAF_K, g.kids.size() - idx);
        o.raw("SNOD", 4); o.u8(1); o.u8(0); o.u16(uint16_t(n));
        for (size_t k = 0; k < n; ++k, ++it, ++idx) {
            Node& c = *it->second;
            o.u64(name_off[idx]); o.u64(c.header);
            if (c.is_group) { o.u32(1); o.u32(0); o.u64(c.btree); o.u64(c.heap); }
            else { o.u32(0); o.u32(0); o.zeros(16); }
        }
        o.zeros((2 * LEAF_K - n) * 40);
        last_name.push_back(n ? name_off[idx - 1] : 0);
        if (g.kids.empty()) break;
    }
    // B-tree node (level 0): key0 = "", key[i+1] = largest name in child i
    o.align8();
    g.btree = o.pos();
    o.raw("TREE", 4); o.u8(0); o.u8(0); o.u16(uint16_t(g.kids.empty() ? 0 : snods.size()));
    o.u64(UNDEF); o.u64(UNDEF);
    o.u64(0);
    size_t written = 0;
    if (!g.kids.empty())
        for (size_t i = 0; i < snods.size(); ++i, ++written) { o.u64(snods[i]); o.u64(last_name[i]); }
    o.zeros((2 * INTERNAL_K - written) * 16);
    // object header with the symbol-table message
    Out stab;
    stab.u64(g.btree); stab.u64(g.heap);
    std::vector<std::pair<uint16_t, std::vector<uint8_t>>> msgs{{0x11, stab.b}};
    for (const Attr& a : g.attrs) msgs.push_back({0x0c, msg_attr(a)});
    g.header = write_header(o, msgs);
}

}  // namespace
}  // namespace hc

struct hc_h5_writer {
    hc::Node root;
    hc::Node* find(const std::string& path, bool create_groups, bool leaf_is_dataset, std::string* leaf = nullptr) {
        hc::Node* cur = &root;
        size_t i = 0;
        std::vector<std::string> parts;
        while (i < path.size()) {
            size_t j = path.find('/', i);
            if (j == std::string::npos) j = path.size();
            if (j > i) parts.push_back(path.substr(i, j - i));
            i = j + 1;
        }
        if (parts.empty()) return cur;
        for (size_t k = 0; k < parts.size(); ++k) {
            const bool last = k + 1 == parts.size();
            if (!cur->is_group) hc::fail(HC_ERR_INVALID, "h5 writer: '" + path + "' crosses a dataset");
            auto it = cur->kids.find(parts[k]);
            if (it == cur->kids.end()) {
                if (!create_groups) return nullptr;
                auto n = std::make_unique<hc::Node>();
                n->is_group = !(last && leaf_is_dataset);
                it = cur->kids.emplace(parts[k], std::move(n)).first;
            } else if (last && leaf_is_dataset) {
                if (it->second->is_group) hc::fail(HC_ERR_INVALID, "h5 writer: '" + path + "' is a group");
            }
            cur = it->second.get();
        }
        if (leaf) *leaf = parts.back();
        return cur;
    }
};

#define HC_GUARD_BEGIN try {
#define HC_GUARD_END                                                                       \
    }                                                                                      \
    catch (const hc::StatusError& e) { hc::set_last_error(e.msg); return e.code; }         \
    catch (const std::exception& e) { hc::set_last_error(e.what()); return HC_ERR_INVALID; }

extern "C" {

hc_status hc_h5_writer_create(hc_h5_writer** out) {
    HC_GUARD_BEGIN
    *out = new hc_h5_writer();
    return HC_OK;
    HC_GUARD_END
}
void hc_h5_writer_destroy(hc_h5_writer* w) { delete w; }

hc_status hc_h5_writer_put_group(hc_h5_writer* w, const char* path) {
    HC_GUARD_BEGIN
    w->find(path, true, false);
    return HC_OK;
    HC_GUARD_END
}

hc_status hc_h5_writer_put_f64(hc_h5_writer* w, const char* path, int rank, const uint64_t* dims, const double* data) {
    HC_GUARD_BEGIN
    if (rank < 0 || rank > 8) hc::fail(HC_ERR_INVALID, "h5 writer: rank must be 0..8");
    hc::Node* n = w->find(path, true, true);
    n->is_group = false; n->is_string = false;
    n->dims.assign(dims, dims + rank);
    uint64_t count = 1;
    for (int i = 0; i < rank; ++i) count *= dims[i];
    n->bytes.resize(count * 8);
    if (count) std::memcpy(n->bytes.data(), data, count * 8);
    return HC_OK;
    HC_GUARD_END
}

hc_status hc_h5_writer_put_string(hc_h5_writer* w, const char* path, const char* value) {
    HC_GUARD_BEGIN
    hc::Node* n = w->find(path, true, true);
    n->is_group = false; n->is_string = true; n->str_size = 0;
    n->dims.clear();
    const size_t len = std::strlen(value);
    n->bytes.assign(value, value + len + 1);
    return HC_OK;
    HC_GUARD_END
}

// 1-D array of fixed-length, null-padded strings (what H5Writer::WriteStringArray produces in the reference)
hc_status hc_h5_writer_put_string_array(hc_h5_writer* w, const char* path, int count, const char* const* values) {
    HC_GUARD_BEGIN
    if (count < 0 || (count > 0 && !values)) hc::fail(HC_ERR_INVALID, "h5 writer: bad string array");
    hc::Node* n = w->find(path, true, true);
    n->is_group = false; n->is_string = true;
    size_t width = 1;
    for (int i = 0; i < count; ++i) width = std::max(width, std::strlen(values[i]) + 1);
    n->str_size = width;
    n->dims.assign(1, uint64_t(count));
    n->bytes.assign(size_t(count) * width, 0);
    for (int i = 0; i < count; ++i) std::memcpy(n->bytes.data() + size_t(i) * width, values[i], std::strlen(values[i]));
    return HC_OK;
    HC_GUARD_END
}

hc_status hc_h5_writer_attr_string(hc_h5_writer* w, const char* path, const char* name, const char* value) {
    HC_GUARD_BEGIN
    hc::Node* n = w->find(path, false, false);
    if (!n) hc::fail(HC_ERR_INVALID, std::string("h5 writer: no object '") + path + "'");
    hc::Attr a; a.name = name; a.is_string = true; a.s = value;
    n->attrs.push_back(a);
    return HC_OK;
    HC_GUARD_END
}

hc_status hc_h5_writer_attr_f64(hc_h5_writer* w, const char* path, const char* name, double value) {
    HC_GUARD_BEGIN
    hc::Node* n = w->find(path, false, false);
    if (!n) hc::fail(HC_ERR_INVALID, std::string("h5 writer: no object '") + path + "'");
    hc::Attr a; a.name = name; a.d = value;
    n->attrs.push_back(a);
    return HC_OK;
    HC_GUARD_END
}

hc_status hc_h5_writer_save(hc_h5_writer* w, const char* file) {
    HC_GUARD_BEGIN
    hc::Out o;
    static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    o.raw(sig, 8);
    o.u8(0); o.u8(0); o.u8(0); o.u8(0); o.u8(0); o.u8(8); o.u8(8); o.u8(0);
    o.u16(hc::LEAF_K); o.u16(hc::INTERNAL_K); o.u32(0);
    o.u64(0); o.u64(hc::UNDEF);
    const uint64_t eof_at = o.pos();
    o.u64(0); o.u64(hc::UNDEF);
    const uint64_t root_entry = o.pos();
    o.zeros(40);
    hc::write_group(o, w->root);
    o.align8();
    o.patch64(eof_at, o.pos());
    o.patch64(root_entry + 8, w->root.header);
    o.b[root_entry + 16] = 1;   // cache type 1: scratch holds the B-tree and heap addresses
    o.patch64(root_entry + 24, w->root.btree);
    o.patch64(root_entry + 32, w->root.heap);
    FILE* f = std::fopen(file, "wb");
    if (!f) hc::fail(HC_ERR_IO, std::string("h5 writer: cannot open '") + file + "' for writing");
    const size_t put = std::fwrite(o.b.data(), 1, o.b.size(), f);
    std::fclose(f);
    if (put != o.b.size()) hc::fail(HC_ERR_IO, "h5 writer: short write");
    return HC_OK;
    HC_GUARD_END
}

}  // extern "C"
