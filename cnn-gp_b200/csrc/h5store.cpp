// h5store.cpp -- native HDF5 block store behind include/cnngp_h5.h (host code only).
//
// Writes and reads the subset of the HDF5 file format that the reference's h5py calls produce
// (cnn_gp/kernel_save_tools.py:7-23, exp_mnist_resnet/save_kernel.py:26-36, classify_gp.py:45-48,
// merge_h5_files.py:15-30) with libhdf5's default settings:
//   superblock version 0 (1 is read too)        symbol-table root group: version-1 B-tree of
//   version-1 object headers (+ continuations)    type 0, SNOD nodes, local heap
//   dataspace v1 (v2 read), IEEE LE f32/f64      chunked storage: version-1 B-tree of type 1
//   fill-value message v2 (v1-v3, old: read)     layout message v3 (v1/v2 read)
// Everything is written from the format specification; there is no libhdf5 in the image.
//
// Strategy: chunk data is appended to the file as it is written; each dataset's chunk index is
// kept in memory (sorted map) and bulk-written as a B-tree on flush, so no node ever splits.
#include "../../include/cnngp_h5.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <sys/uio.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_err;
int fail(const std::string &m) { g_err = m; return 1; }

constexpr uint64_t UNDEF = ~0ull;
const unsigned char kSig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
constexpr int kMaxRank = CNNGP_H5_MAX_RANK;

inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | p[1] << 8); }
inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
inline uint64_t rd64(const uint8_t *p) { return (uint64_t)rd32(p) | (uint64_t)rd32(p + 4) << 32; }

struct Buf {  // little-endian serialiser
    std::vector<uint8_t> b;
    void u8(unsigned v) { b.push_back((uint8_t)v); }
    void u16(unsigned v) { u8(v & 255); u8(v >> 8 & 255); }
    void u32(uint32_t v) { u16(v & 0xffff); u16(v >> 16); }
    void u64(uint64_t v) { u32((uint32_t)v); u32((uint32_t)(v >> 32)); }
    void raw(const void *p, size_t n) { const uint8_t *q = (const uint8_t *)p; b.insert(b.end(), q, q + n); }
    void zeros(size_t n) { b.insert(b.end(), n, 0); }
    void pad8() { while (b.size() % 8) b.push_back(0); }
    size_t size() const { return b.size(); }
};

struct Dataset {
    std::string name;
    int rank = 0, dtype = -1, esize = 0;
    int64_t shape[kMaxRank] = {0}, maxshape[kMaxRank] = {0}, chunk[kMaxRank] = {0};
    bool chunked = false, compact = false, has_fill = false, filtered = false;
    std::string unsupported;            // why the data cannot be accessed (empty: fine)
    uint8_t fill[8] = {0};
    uint64_t ohdr = UNDEF;
    uint64_t dims_pos = UNDEF;          // address of the dataspace message's dimension sizes
    uint64_t layout_addr_pos = UNDEF;   // address of the layout message's address field
    uint64_t data_addr = UNDEF;         // contiguous storage
    std::vector<uint8_t> compact_data;
    // chunked storage: chunk offset in elements (rank entries, lexicographic) -> file address
    std::map<std::vector<uint64_t>, uint64_t> chunks;
    uint64_t chunk_bytes = 0;
    uint64_t index_addr = UNDEF, index_cap = 0;  // file space owned by the on-disk index
    uint64_t root = UNDEF;
    bool index_dirty = false, dims_dirty = false;
};

struct Link {
    std::string name;
    uint64_t ohdr = UNDEF;
    uint32_t ctype = 0;
    uint8_t scratch[16] = {0};
};

}  // namespace

struct cnngp_h5 {
    int fd = -1;
    bool writable = false, created = false;
    std::string path;
    std::mutex mu;
    uint64_t base = 0;  // user-block size: file position = base + address
    uint64_t eof = 0;   // allocation pointer (address)
    int leaf_k = 4, internal_k = 16, chunk_k = 32, sb_version = 0;
    uint64_t sb_eof_pos = 0;  // file position of the superblock's end-of-file address
    uint64_t root_ohdr = UNDEF, root_btree = UNDEF, root_heap = UNDEF;
    uint64_t heap_data_addr = UNDEF, heap_data_size = 0;
    int group_levels = 0;
    std::vector<uint64_t> snods;
    std::vector<Link> links;  // every link of the root group
    bool group_dirty = false, any_dirty = false;
    std::vector<std::unique_ptr<Dataset>> dsets;
};

namespace {

using File = cnngp_h5;

bool pread_all(const File *f, void *buf, size_t n, uint64_t addr) {
    uint8_t *p = (uint8_t *)buf;
    uint64_t off = f->base + addr;
    while (n) {
        ssize_t r = pread(f->fd, p, n, (off_t)off);
        if (r < 0 && errno == EINTR) continue;
        if (r <= 0) return false;
        p += r; off += (uint64_t)r; n -= (size_t)r;
    }
    return true;
}
bool pwrite_all(const File *f, const void *buf, size_t n, uint64_t addr) {
    const uint8_t *p = (const uint8_t *)buf;
    uint64_t off = f->base + addr;
    while (n) {
        ssize_t r = pwrite(f->fd, p, n, (off_t)off);
        if (r < 0 && errno == EINTR) continue;
        if (r <= 0) return false;
        p += r; off += (uint64_t)r; n -= (size_t)r;
    }
    return true;
}
uint64_t alloc(File *f, uint64_t n) {
    const uint64_t a = (f->eof + 7) & ~7ull;
    f->eof = a + n;
    f->any_dirty = true;
    return a;
}

// ---- object headers (version 1) -------------------------------------------------------------
struct Msg { uint16_t type; uint8_t flags; uint64_t addr; std::vector<uint8_t> body; };

bool read_messages(File *f, uint64_t ohdr, std::vector<Msg> &out, std::string &err) {
    uint8_t h[16];
    if (ohdr == UNDEF || !pread_all(f, h, 16, ohdr)) { err = "object header outside the file"; return false; }
    if (h[0] != 1) { err = "object header version " + std::to_string(h[0]) + " (only version 1 is read; was the file written with libver='latest'?)"; return false; }
    const unsigned nmsg = rd16(h + 2);
    std::vector<std::pair<uint64_t, uint64_t>> blocks{{ohdr + 16, rd32(h + 8)}};
    size_t seen = 0;
    for (size_t bi = 0; bi < blocks.size() && seen < nmsg; ++bi) {
        const uint64_t a = blocks[bi].first, n = blocks[bi].second;
        if (n > (1u << 26)) { err = "object header block too large"; return false; }
        std::vector<uint8_t> blk(n);
        if (n && !pread_all(f, blk.data(), n, a)) { err = "object header block outside the file"; return false; }
        uint64_t p = 0;
        while (p + 8 <= n && seen < nmsg) {
            Msg m;
            m.type = rd16(&blk[p]);
            const unsigned sz = rd16(&blk[p + 2]);
            m.flags = blk[p + 4];
            if (p + 8 + sz > n) { err = "object header message overruns its block"; return false; }
            m.addr = a + p + 8;
            m.body.assign(blk.begin() + p + 8, blk.begin() + p + 8 + sz);
            p += 8 + sz;
            ++seen;
            if (m.type == 0x0010 && sz >= 16) blocks.push_back({rd64(m.body.data()), rd64(m.body.data() + 8)});
            else if (m.type != 0) out.push_back(std::move(m));
        }
    }
    return true;
}

int parse_datatype(const std::vector<uint8_t> &b, int &esize) {
    if (b.size() < 8) return -1;
    const int cls = b[0] & 15;
    esize = (int)rd32(&b[4]);
    if (cls != 1 || b.size() < 20) return -1;
    if (b[1] & 1) return -1;               // big-endian
    if (((b[1] >> 4) & 3) != 2) return -1; // mantissa normalisation must be "msb implied"
    const unsigned boff = rd16(&b[8]), prec = rd16(&b[10]), eloc = b[12], esz = b[13], mloc = b[14], msz = b[15];
    const uint32_t bias = rd32(&b[16]);
    if (esize == 4 && boff == 0 && prec == 32 && eloc == 23 && esz == 8 && mloc == 0 && msz == 23 && bias == 127 && b[2] == 31) return 0;
    if (esize == 8 && boff == 0 && prec == 64 && eloc == 52 && esz == 11 && mloc == 0 && msz == 52 && bias == 1023 && b[2] == 63) return 1;
    return -1;
}

bool walk_chunk_tree(File *f, Dataset *d, uint64_t addr, int depth, std::string &err) {
    if (depth > 16) { err = "chunk B-tree too deep"; return false; }
    const int nd = d->rank + 1;
    const size_t ksz = 8 + 8 * (size_t)nd;
    uint8_t h[24];
    if (!pread_all(f, h, 24, addr) || memcmp(h, "TREE", 4) != 0 || h[4] != 1) { err = "bad chunk B-tree node"; return false; }
    const int level = h[5];
    const unsigned used = rd16(h + 6);
    if (used > 2u * (unsigned)f->chunk_k) { err = "chunk B-tree node over capacity"; return false; }
    std::vector<uint8_t> body(used * (ksz + 8) + ksz);
    if (!pread_all(f, body.data(), body.size(), addr + 24)) { err = "chunk B-tree node outside the file"; return false; }
    for (unsigned e = 0; e < used; ++e) {
        const uint8_t *k = &body[e * (ksz + 8)];
        const uint64_t child = rd64(k + ksz);
        if (level > 0) {
            if (!walk_chunk_tree(f, d, child, depth + 1, err)) return false;
            continue;
        }
        if (rd32(k) != d->chunk_bytes || rd32(k + 4) != 0) { err = "chunk size/filter mask not those of an unfiltered chunk"; return false; }
        std::vector<uint64_t> off(d->rank);
        for (int i = 0; i < d->rank; ++i) off[i] = rd64(k + 8 + 8 * i);
        d->chunks[off] = child;
    }
    return true;
}

// one link target -> Dataset (or nullptr when it is not a dataset); fatal format errors -> err
std::unique_ptr<Dataset> parse_dataset(File *f, const Link &ln, std::string &err) {
    std::vector<Msg> msgs;
    if (!read_messages(f, ln.ohdr, msgs, err)) return nullptr;
    const Msg *space = nullptr, *type = nullptr, *layout = nullptr, *fill_new = nullptr, *fill_old = nullptr;
    bool filtered = false;
    for (const Msg &m : msgs) {
        if (m.type == 0x0001) space = &m;
        else if (m.type == 0x0003) type = &m;
        else if (m.type == 0x0008) layout = &m;
        else if (m.type == 0x0005) fill_new = &m;
        else if (m.type == 0x0004) fill_old = &m;
        else if (m.type == 0x000B) filtered = true;
        else if (m.type == 0x0011) return nullptr;  // a group
    }
    if (!space || !type || !layout) return nullptr;
    std::unique_ptr<Dataset> d(new Dataset);
    d->name = ln.name;
    d->ohdr = ln.ohdr;
    d->filtered = filtered;
    {   // dataspace
        const std::vector<uint8_t> &b = space->body;
        if (b.size() < 4) { err = "short dataspace message"; return nullptr; }
        const int ver = b[0];
        d->rank = b[1];
        const int flags = b[2];
        const size_t off = ver == 1 ? 8 : 4;
        if (ver != 1 && ver != 2) { err = "dataspace message version " + std::to_string(ver); return nullptr; }
        if (d->rank > kMaxRank) { err = "dataset rank above " + std::to_string(kMaxRank); return nullptr; }
        if (b.size() < off + 8 * (size_t)d->rank * ((flags & 1) ? 2 : 1)) { err = "short dataspace message"; return nullptr; }
        d->dims_pos = space->addr + off;
        for (int i = 0; i < d->rank; ++i) {
            d->shape[i] = (int64_t)rd64(&b[off + 8 * i]);
            const uint64_t mx = (flags & 1) ? rd64(&b[off + 8 * (d->rank + i)]) : (uint64_t)d->shape[i];
            d->maxshape[i] = mx == UNDEF ? CNNGP_H5_UNLIMITED : (int64_t)mx;
        }
    }
    if (type->flags & 2) d->unsupported = "shared (committed) datatype";
    else d->dtype = parse_datatype(type->body, d->esize);
    if (d->dtype < 0 && d->unsupported.empty()) d->unsupported = "element type is not little-endian IEEE float32/float64";
    {   // fill value
        const uint8_t *fv = nullptr;
        uint32_t fsz = 0;
        if (fill_new) {
            const std::vector<uint8_t> &b = fill_new->body;
            const int ver = b.empty() ? 0 : b[0];
            if ((ver == 1 || ver == 2) && b.size() >= 4) {
                if ((ver == 1 || b[3]) && b.size() >= 8) { fsz = rd32(&b[4]); if (b.size() >= 8 + (size_t)fsz) fv = &b[8]; }
            } else if (ver == 3 && b.size() >= 2) {
                if ((b[1] & 0x20) && b.size() >= 6) { fsz = rd32(&b[2]); if (b.size() >= 6 + (size_t)fsz) fv = &b[6]; }
            }
        }
        if (!fv && fill_old && fill_old->body.size() >= 4) {
            fsz = rd32(fill_old->body.data());
            if (fill_old->body.size() >= 4 + (size_t)fsz) fv = fill_old->body.data() + 4;
        }
        if (fv && fsz > 0 && (int)fsz == d->esize && fsz <= 8) { memcpy(d->fill, fv, fsz); d->has_fill = true; }
    }
    {   // layout
        const std::vector<uint8_t> &b = layout->body;
        if (b.size() < 3) { err = "short layout message"; return nullptr; }
        const int ver = b[0];
        int cls, nd = 0;
        size_t p;
        uint64_t addr = UNDEF;
        if (ver == 1 || ver == 2) {
            nd = b[1]; cls = b[2]; p = 8;
            if (cls != 0) { if (b.size() < p + 8) { err = "short layout message"; return nullptr; } d->layout_addr_pos = layout->addr + p; addr = rd64(&b[p]); p += 8; }
            if (b.size() < p + 4 * (size_t)nd) { err = "short layout message"; return nullptr; }
            if (cls == 2) for (int i = 0; i < nd - 1 && i < kMaxRank; ++i) d->chunk[i] = rd32(&b[p + 4 * i]);
            p += 4 * (size_t)nd;
            if (cls == 0 && b.size() >= p + 4) { const uint32_t n = rd32(&b[p]); if (b.size() >= p + 4 + n) d->compact_data.assign(b.begin() + p + 4, b.begin() + p + 4 + n); }
        } else if (ver == 3) {
            cls = b[1];
            if (cls == 0) {
                const unsigned n = b.size() >= 4 ? rd16(&b[2]) : 0;
                if (b.size() >= 4 + (size_t)n) d->compact_data.assign(b.begin() + 4, b.begin() + 4 + n);
            } else if (cls == 1) {
                if (b.size() < 18) { err = "short layout message"; return nullptr; }
                d->layout_addr_pos = layout->addr + 2; addr = rd64(&b[2]);
            } else if (cls == 2) {
                nd = b[2];
                if (b.size() < 11 + 4 * (size_t)nd) { err = "short layout message"; return nullptr; }
                d->layout_addr_pos = layout->addr + 3; addr = rd64(&b[3]);
                for (int i = 0; i < nd - 1 && i < kMaxRank; ++i) d->chunk[i] = rd32(&b[11 + 4 * i]);
            }
        } else {
            d->unsupported = "data layout message version " + std::to_string(ver) + " (written with libver='latest'?)";
            return d;
        }
        if (cls == 0) d->compact = true;
        else if (cls == 1) d->data_addr = addr;
        else if (cls == 2) {
            if (nd != d->rank + 1) { err = "chunk dimensionality does not match the dataspace"; return nullptr; }
            d->chunked = true;
            d->chunk_bytes = (uint64_t)d->esize;
            for (int i = 0; i < d->rank; ++i) d->chunk_bytes *= (uint64_t)d->chunk[i];
            d->root = addr;
            if (addr != UNDEF && !filtered && d->unsupported.empty()) {
                d->index_addr = UNDEF;  // nodes of a foreign index are never reused
                if (!walk_chunk_tree(f, d.get(), addr, 0, err)) return nullptr;
            }
        } else { d->unsupported = "layout class " + std::to_string(cls); }
    }
    if (filtered && d->unsupported.empty()) d->unsupported = "filtered (compressed) dataset";
    return d;
}

// ---- root group ------------------------------------------------------------------------------
bool walk_group_tree(File *f, uint64_t addr, const std::vector<uint8_t> &heap, int depth, std::string &err) {
    if (depth > 16) { err = "group B-tree too deep"; return false; }
    uint8_t h[24];
    if (!pread_all(f, h, 24, addr) || memcmp(h, "TREE", 4) != 0 || h[4] != 0) { err = "bad group B-tree node"; return false; }
    const int level = h[5];
    const unsigned used = rd16(h + 6);
    if (depth == 0) f->group_levels = level;
    std::vector<uint8_t> body(used * 16 + 8);
    if (!pread_all(f, body.data(), body.size(), addr + 24)) { err = "group B-tree node outside the file"; return false; }
    for (unsigned e = 0; e < used; ++e) {
        const uint64_t child = rd64(&body[e * 16 + 8]);
        if (level > 0) { if (!walk_group_tree(f, child, heap, depth + 1, err)) return false; continue; }
        uint8_t sh[8];
        if (!pread_all(f, sh, 8, child) || memcmp(sh, "SNOD", 4) != 0) { err = "bad symbol-table node"; return false; }
        const unsigned n = rd16(sh + 6);
        std::vector<uint8_t> ents(40 * (size_t)n);
        if (n && !pread_all(f, ents.data(), ents.size(), child + 8)) { err = "symbol-table node outside the file"; return false; }
        f->snods.push_back(child);
        for (unsigned i = 0; i < n; ++i) {
            const uint8_t *q = &ents[40 * i];
            Link ln;
            const uint64_t noff = rd64(q);
            if (noff >= heap.size()) { err = "link name outside the local heap"; return false; }
            const void *z = memchr(&heap[noff], 0, heap.size() - noff);
            if (!z) { err = "unterminated link name"; return false; }
            ln.name.assign((const char *)&heap[noff]);
            ln.ohdr = rd64(q + 8);
            ln.ctype = rd32(q + 16);
            memcpy(ln.scratch, q + 24, 16);
            f->links.push_back(ln);
        }
    }
    return true;
}

int load_existing(File *f) {
    struct stat st;
    if (fstat(f->fd, &st) != 0) return fail("fstat failed: " + f->path);
    const uint64_t fsize = (uint64_t)st.st_size;
    uint64_t sb = 0;
    uint8_t s[8];
    for (;;) {  // the superblock sits at 0, 512, 1024, ... (user block)
        f->base = 0;
        if (sb + 8 > fsize || !pread_all(f, s, 8, sb)) return fail("not an HDF5 file (no signature): " + f->path);
        if (memcmp(s, kSig, 8) == 0) break;
        sb = sb == 0 ? 512 : sb * 2;
    }
    uint8_t h[136];
    memset(h, 0, sizeof h);
    const size_t avail = (size_t)std::min<uint64_t>(sizeof h, fsize - sb);
    pread_all(f, h, avail, sb);
    f->sb_version = h[8];
    if (f->sb_version > 1) return fail("superblock version " + std::to_string(f->sb_version) + " is not read (file written with libver='latest'?): " + f->path);
    if (h[13] != 8 || h[14] != 8) return fail("only 8-byte offsets and lengths are read");
    f->leaf_k = rd16(h + 16);
    f->internal_k = rd16(h + 18);
    size_t p = 24;
    if (f->sb_version == 1) { f->chunk_k = rd16(h + p); p += 4; }
    const uint64_t base = rd64(h + p), stored_eof = rd64(h + p + 16);
    f->sb_eof_pos = sb + p + 16;
    p += 32;
    if (base != sb) return fail("base address differs from the superblock position");
    if (stored_eof > fsize) return fail("truncated file: end-of-file address " + std::to_string(stored_eof) + " beyond " + std::to_string(fsize) + " bytes");
    f->base = base;
    f->eof = std::max(stored_eof, fsize) - base;
    if (f->writable && base != 0) return fail("files with a user block are read-only here");
    f->root_ohdr = rd64(h + p + 8);
    std::string err;
    std::vector<Msg> msgs;
    if (!read_messages(f, f->root_ohdr, msgs, err)) return fail("root group: " + err);
    const Msg *stab = nullptr;
    for (const Msg &m : msgs) if (m.type == 0x0011 && m.body.size() >= 16) stab = &m;
    if (!stab) return fail("root group has no symbol table (new-style groups are not read)");
    f->root_btree = rd64(stab->body.data());
    f->root_heap = rd64(stab->body.data() + 8);
    uint8_t hh[32];
    if (!pread_all(f, hh, 32, f->root_heap) || memcmp(hh, "HEAP", 4) != 0) return fail("bad local heap");
    f->heap_data_size = rd64(hh + 8);
    f->heap_data_addr = rd64(hh + 24);
    if (f->heap_data_size > (1ull << 30)) return fail("local heap too large");
    std::vector<uint8_t> heap(f->heap_data_size);
    if (f->heap_data_size && !pread_all(f, heap.data(), heap.size(), f->heap_data_addr)) return fail("local heap data outside the file");
    if (!walk_group_tree(f, f->root_btree, heap, 0, err)) return fail(err);
    std::sort(f->links.begin(), f->links.end(), [](const Link &a, const Link &b) { return a.name < b.name; });
    for (const Link &ln : f->links) {
        err.clear();
        std::unique_ptr<Dataset> d = parse_dataset(f, ln, err);
        if (!d && !err.empty()) return fail("dataset '" + ln.name + "': " + err);
        if (d) f->dsets.push_back(std::move(d));
    }
    return 0;
}

bool write_group(File *f) {
    std::sort(f->links.begin(), f->links.end(), [](const Link &a, const Link &b) { return a.name < b.name; });
    const size_t per = 2 * (size_t)f->leaf_k, n = f->links.size();
    const size_t n_snod = (n + per - 1) / per;
    if (f->group_levels > 0 || n_snod > 2 * (size_t)f->internal_k) { g_err = "root group too large to modify"; return false; }
    // local heap: "" at offset 0, then every name, each padded to 8 bytes
    Buf heap;
    heap.zeros(8);
    std::vector<uint64_t> noff(n);
    for (size_t i = 0; i < n; ++i) {
        noff[i] = heap.size();
        heap.raw(f->links[i].name.c_str(), f->links[i].name.size() + 1);
        heap.pad8();
    }
    if (heap.size() > f->heap_data_size) {
        uint64_t sz = std::max<uint64_t>(256, f->heap_data_size);
        while (sz < heap.size()) sz *= 2;
        f->heap_data_addr = alloc(f, sz);
        f->heap_data_size = sz;
    }
    uint64_t free_head = 1;  // H5HL_FREE_NULL
    const uint64_t used = heap.size();
    if (f->heap_data_size - used >= 16) {
        free_head = used;
        heap.u64(1);
        heap.u64(f->heap_data_size - used);
    }
    heap.zeros(f->heap_data_size - heap.size());
    Buf hh;
    hh.raw("HEAP", 4); hh.u8(0); hh.zeros(3);
    hh.u64(f->heap_data_size); hh.u64(free_head); hh.u64(f->heap_data_addr);
    if (!pwrite_all(f, hh.b.data(), hh.size(), f->root_heap) || !pwrite_all(f, heap.b.data(), heap.size(), f->heap_data_addr)) return false;
    // symbol-table nodes
    const size_t snod_bytes = 8 + per * 40;
    while (f->snods.size() < n_snod) f->snods.push_back(alloc(f, snod_bytes));
    for (size_t s = 0; s < n_snod; ++s) {
        const size_t lo = s * per, hi = std::min(n, lo + per);
        Buf b;
        b.raw("SNOD", 4); b.u8(1); b.u8(0); b.u16((unsigned)(hi - lo));
        for (size_t i = lo; i < hi; ++i) {
            b.u64(noff[i]); b.u64(f->links[i].ohdr); b.u32(f->links[i].ctype); b.u32(0);
            b.raw(f->links[i].scratch, 16);
        }
        b.zeros(snod_bytes - b.size());
        if (!pwrite_all(f, b.b.data(), b.size(), f->snods[s])) return false;
    }
    // the B-tree node above them
    Buf t;
    t.raw("TREE", 4); t.u8(0); t.u8(0); t.u16((unsigned)n_snod); t.u64(UNDEF); t.u64(UNDEF);
    t.u64(0);
    for (size_t s = 0; s < n_snod; ++s) {
        t.u64(f->snods[s]);
        t.u64(noff[std::min(n, (s + 1) * per) - 1]);
    }
    const size_t node_bytes = 24 + 2 * (size_t)f->internal_k * 16 + 8;
    t.zeros(node_bytes - t.size());
    if (!pwrite_all(f, t.b.data(), t.size(), f->root_btree)) return false;
    f->group_dirty = false;
    return true;
}

void superblock_bytes(const File *f, Buf &b) {
    b.raw(kSig, 8);
    b.u8(0); b.u8(0); b.u8(0); b.u8(0); b.u8(0);  // versions: superblock, free space, root entry, reserved, shared header
    b.u8(8); b.u8(8); b.u8(0);                    // sizes of offsets and lengths
    b.u16((unsigned)f->leaf_k); b.u16((unsigned)f->internal_k);
    b.u32(0);                                      // consistency flags
    b.u64(0); b.u64(UNDEF); b.u64(f->eof); b.u64(UNDEF);  // base, free-space info, end of file, driver info
    b.u64(0); b.u64(f->root_ohdr); b.u32(1); b.u32(0);    // root entry: cached symbol-table addresses
    b.u64(f->root_btree); b.u64(f->root_heap);
}

int create_new(File *f) {
    f->base = 0;
    f->created = true;
    f->sb_eof_pos = 40;
    f->eof = 96;
    f->root_ohdr = alloc(f, 40);
    f->root_btree = alloc(f, 24 + 2 * (size_t)f->internal_k * 16 + 8);
    f->root_heap = alloc(f, 32);
    f->heap_data_size = 256;
    f->heap_data_addr = alloc(f, f->heap_data_size);
    Buf oh;  // root object header: one symbol-table message
    oh.u8(1); oh.u8(0); oh.u16(1); oh.u32(1); oh.u32(24); oh.u32(0);
    oh.u16(0x0011); oh.u16(16); oh.u8(0); oh.zeros(3);
    oh.u64(f->root_btree); oh.u64(f->root_heap);
    if (!pwrite_all(f, oh.b.data(), oh.size(), f->root_ohdr)) return fail("write failed: " + f->path);
    f->group_dirty = true;
    return 0;
}

// ---- chunk index: bulk-written version-1 B-tree ---------------------------------------------
bool write_index(File *f, Dataset *d) {
    const int nd = d->rank + 1;
    const size_t ksz = 8 + 8 * (size_t)nd, cap = 2 * (size_t)f->chunk_k;
    const size_t node_bytes = 24 + (cap + 1) * ksz + cap * 8;
    const size_t n = d->chunks.size();
    if (n == 0) {
        d->root = UNDEF;
    } else {
        // level sizes, leaves first
        std::vector<size_t> level_nodes;
        for (size_t m = n;;) { m = (m + cap - 1) / cap; level_nodes.push_back(m); if (m == 1) break; }
        size_t total = 0;
        for (size_t m : level_nodes) total += m;
        if (d->index_addr == UNDEF || total * node_bytes > d->index_cap) {
            d->index_cap = std::max<size_t>(2 * total, 4) * node_bytes;  // room to grow between flushes
            d->index_addr = alloc(f, d->index_cap);
        }
        auto put_key = [&](Buf &b, uint32_t size, const uint64_t *off) {
            b.u32(size); b.u32(0);
            for (int i = 0; i < d->rank; ++i) b.u64(off[i]);
            b.u64(0);
        };
        // the key to the right of everything: one chunk past the last one in every dimension
        std::vector<uint64_t> last = d->chunks.rbegin()->first;
        Buf final_key;
        final_key.u32(0); final_key.u32(0);
        for (int i = 0; i < d->rank; ++i) final_key.u64(last[i] + (uint64_t)d->chunk[i]);
        final_key.u64((uint64_t)d->esize);
        // children of the level being written: (first chunk offset, address)
        std::vector<std::pair<const std::vector<uint64_t> *, uint64_t>> kids;
        kids.reserve(n);
        for (const auto &kv : d->chunks) kids.push_back({&kv.first, kv.second});
        uint64_t level_base = d->index_addr;
        std::vector<uint8_t> image;
        for (size_t lv = 0; lv < level_nodes.size(); ++lv) {
            const size_t m = level_nodes[lv];
            std::vector<std::pair<const std::vector<uint64_t> *, uint64_t>> parents;
            image.clear();
            for (size_t i = 0; i < m; ++i) {
                const size_t lo = i * cap, hi = std::min(kids.size(), lo + cap);
                const uint64_t addr = level_base + i * node_bytes;
                Buf b;
                b.raw("TREE", 4); b.u8(1); b.u8((unsigned)lv); b.u16((unsigned)(hi - lo));
                b.u64(i == 0 ? UNDEF : addr - node_bytes);
                b.u64(i + 1 == m ? UNDEF : addr + node_bytes);
                for (size_t e = lo; e < hi; ++e) {
                    put_key(b, (uint32_t)d->chunk_bytes, kids[e].first->data());
                    b.u64(kids[e].second);
                }
                if (hi < kids.size()) put_key(b, (uint32_t)d->chunk_bytes, kids[hi].first->data());
                else b.raw(final_key.b.data(), final_key.size());
                b.zeros(node_bytes - b.size());
                image.insert(image.end(), b.b.begin(), b.b.end());
                parents.push_back({kids[lo].first, addr});
            }
            if (!pwrite_all(f, image.data(), image.size(), level_base)) return false;
            level_base += m * node_bytes;
            kids.swap(parents);
        }
        d->root = kids[0].second;
    }
    uint8_t a[8];
    for (int i = 0; i < 8; ++i) a[i] = (uint8_t)(d->root >> (8 * i));
    if (!pwrite_all(f, a, 8, d->layout_addr_pos)) return false;
    d->index_dirty = false;
    return true;
}

bool write_dims(File *f, Dataset *d) {
    Buf b;
    for (int i = 0; i < d->rank; ++i) b.u64((uint64_t)d->shape[i]);
    if (!pwrite_all(f, b.b.data(), b.size(), d->dims_pos)) return false;
    d->dims_dirty = false;
    return true;
}

int flush_locked(File *f) {
    if (!f->writable) return 0;
    for (auto &d : f->dsets) {
        if (d->index_dirty && !write_index(f, d.get())) return fail("writing the chunk index failed: " + f->path);
        if (d->dims_dirty && !write_dims(f, d.get())) return fail("writing the dataspace failed: " + f->path);
    }
    if (f->group_dirty && !write_group(f)) return fail("writing the root group failed: " + g_err);
    if (f->any_dirty) {
        // the file must be at least as long as the end-of-file address says
        struct stat st;
        if (fstat(f->fd, &st) == 0 && (uint64_t)st.st_size < f->base + f->eof && ftruncate(f->fd, (off_t)(f->base + f->eof)) != 0)
            return fail("ftruncate failed: " + f->path);
        if (f->created) {
            Buf sb;
            superblock_bytes(f, sb);
            if (!pwrite_all(f, sb.b.data(), sb.size(), 0)) return fail("writing the superblock failed: " + f->path);
        } else {
            uint8_t a[8];
            const uint64_t v = f->base + f->eof;
            for (int i = 0; i < 8; ++i) a[i] = (uint8_t)(v >> (8 * i));
            if (pwrite(f->fd, a, 8, (off_t)f->sb_eof_pos) != 8) return fail("writing the end-of-file address failed: " + f->path);
        }
        f->any_dirty = false;
    }
    return 0;
}

// ---- hyperslab transfer ------------------------------------------------------------------------
// Copy the box [lo, hi) (dataset coordinates) between a chunk-shaped buffer whose origin is `corg`
// and a selection-shaped buffer whose origin is `sorg`.  to_user: chunk -> user, else user -> chunk.
void copy_box(int rank, const int64_t *lo, const int64_t *hi, const int64_t *corg, const int64_t *cdim,
              uint8_t *cbuf, const int64_t *sorg, const int64_t *sstr, uint8_t *ubuf, int esize, bool to_user) {
    int64_t cstr[kMaxRank], idx[kMaxRank];
    cstr[rank - 1] = esize;
    for (int i = rank - 2; i >= 0; --i) cstr[i] = cstr[i + 1] * cdim[i + 1];
    for (int i = 0; i < rank; ++i) { if (hi[i] <= lo[i]) return; idx[i] = lo[i]; }
    const size_t row = (size_t)(hi[rank - 1] - lo[rank - 1]) * (size_t)esize;
    for (;;) {
        int64_t co = 0, so = 0;
        for (int i = 0; i < rank; ++i) { co += (idx[i] - corg[i]) * cstr[i]; so += (idx[i] - sorg[i]) * sstr[i]; }
        if (to_user) memcpy(ubuf + so, cbuf + co, row); else memcpy(cbuf + co, ubuf + so, row);
        int k = rank - 2;
        for (; k >= 0; --k) { if (++idx[k] < hi[k]) break; idx[k] = lo[k]; }
        if (k < 0) break;
    }
}

// A chunk that lies wholly inside the selection is one contiguous file range and a set of equally
// long rows in the user's buffer: move it with preadv / pwritev, no staging copy.
bool whole_chunk_io(const File *f, int rank, const int64_t *corg, const int64_t *cdim, const int64_t *sorg,
                    const int64_t *sstr, uint8_t *ubuf, int esize, uint64_t addr, bool write) {
    int64_t idx[kMaxRank];
    for (int i = 0; i < rank; ++i) idx[i] = corg[i];
    const size_t row = (size_t)cdim[rank - 1] * (size_t)esize;
    std::vector<iovec> iov;
    iov.reserve(1024);
    uint64_t off = f->base + addr;
    auto flush = [&]() -> bool {
        size_t done = 0, want = iov.size() * row;
        size_t first = 0;
        while (done < want) {
            const ssize_t r = write ? pwritev(f->fd, iov.data() + first, (int)(iov.size() - first), (off_t)(off + done))
                                    : preadv(f->fd, iov.data() + first, (int)(iov.size() - first), (off_t)(off + done));
            if (r < 0 && errno == EINTR) continue;
            if (r <= 0) return false;
            done += (size_t)r;
            size_t rem = (size_t)r;  // advance past what was transferred (short transfers are rare)
            while (rem && first < iov.size()) {
                if (rem >= iov[first].iov_len) { rem -= iov[first].iov_len; ++first; }
                else { iov[first].iov_base = (uint8_t *)iov[first].iov_base + rem; iov[first].iov_len -= rem; rem = 0; }
            }
        }
        off += want;
        iov.clear();
        return true;
    };
    for (;;) {
        int64_t so = 0;
        for (int i = 0; i < rank; ++i) so += (idx[i] - sorg[i]) * sstr[i];
        iov.push_back(iovec{ubuf + so, row});
        if (iov.size() == 1024 && !flush()) return false;
        int k = rank - 2;
        for (; k >= 0; --k) { if (++idx[k] < corg[k] + cdim[k]) break; idx[k] = corg[k]; }
        if (k < 0) break;
    }
    return iov.empty() || flush();
}

void fill_pattern(const Dataset *d, uint8_t *buf, size_t bytes) {
    bool zero = true;
    for (int i = 0; i < d->esize; ++i) zero = zero && d->fill[i] == 0;
    if (!d->has_fill || zero) { memset(buf, 0, bytes); return; }
    for (size_t o = 0; o + (size_t)d->esize <= bytes; o += (size_t)d->esize) memcpy(buf + o, d->fill, (size_t)d->esize);
}

struct ChunkJob { std::vector<uint64_t> off; uint64_t addr; bool existed; };

template <class Fn>
bool run_jobs(size_t n, Fn fn) {  // fn(job index, scratch id) -> bool; a few I/O threads for large selections
    unsigned nt = std::min<unsigned>(8, std::max(1u, std::thread::hardware_concurrency()));
    if (n < 8) nt = 1;
    if (nt == 1) { for (size_t i = 0; i < n; ++i) if (!fn(i)) return false; return true; }
    std::vector<std::thread> th;
    std::vector<char> ok(nt, 1);
    for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([&, t] { for (size_t i = t; i < n; i += nt) if (!fn(i)) { ok[t] = 0; return; } });
    for (auto &x : th) x.join();
    for (char c : ok) if (!c) return false;
    return true;
}

int check_sel(const Dataset *d, const int64_t *start, const int64_t *count) {
    for (int i = 0; i < d->rank; ++i)
        if (start[i] < 0 || count[i] < 0 || start[i] + count[i] > d->shape[i])
            return fail("selection out of range in dimension " + std::to_string(i) + " of '" + d->name + "'");
    return 0;
}

// ustride: byte strides of the caller's array per dimension (the last one must be the element size),
// or NULL for a C-contiguous array of shape count[]
int transfer(File *f, Dataset *d, const int64_t *start, const int64_t *count, uint8_t *user, bool write,
             const int64_t *ustride = nullptr) {
    if (!d->unsupported.empty()) return fail("dataset '" + d->name + "': " + d->unsupported);
    if (check_sel(d, start, count)) return 1;
    const int r = d->rank;
    int64_t sstr[kMaxRank] = {0};
    if (r > 0) {
        sstr[r - 1] = d->esize;
        for (int i = r - 2; i >= 0; --i) sstr[i] = sstr[i + 1] * count[i + 1];
        if (ustride) {
            if (ustride[r - 1] != d->esize) return fail("strided transfer: the last dimension must be contiguous");
            for (int i = 0; i < r; ++i) sstr[i] = ustride[i];
        }
    }
    int64_t total = 1;
    for (int i = 0; i < r; ++i) total *= count[i];
    if (total == 0) return 0;
    if (r == 0) return fail("scalar datasets are not supported");
    if (!d->chunked) {
        int64_t lo[kMaxRank], hi[kMaxRank], zero[kMaxRank] = {0};
        for (int i = 0; i < r; ++i) { lo[i] = start[i]; hi[i] = start[i] + count[i]; }
        if (d->compact) {
            if (write) return fail("compact datasets are read-only here");
            copy_box(r, lo, hi, zero, d->shape, d->compact_data.data(), start, sstr, user, d->esize, true);
            return 0;
        }
        // row by row straight to / from the file (storage never allocated: rows of the fill value)
        int64_t str[kMaxRank], idx[kMaxRank];
        str[r - 1] = d->esize;
        for (int i = r - 2; i >= 0; --i) str[i] = str[i + 1] * d->shape[i + 1];
        for (int i = 0; i < r; ++i) idx[i] = lo[i];
        const size_t row = (size_t)count[r - 1] * (size_t)d->esize;
        if (d->data_addr == UNDEF && write) return fail("contiguous dataset without allocated storage");
        for (;;) {
            int64_t fo = 0, so = 0;
            for (int i = 0; i < r; ++i) { fo += idx[i] * str[i]; so += (idx[i] - start[i]) * sstr[i]; }
            bool ok = true;
            if (d->data_addr == UNDEF) fill_pattern(d, user + so, row);
            else ok = write ? pwrite_all(f, user + so, row, d->data_addr + (uint64_t)fo)
                            : pread_all(f, user + so, row, d->data_addr + (uint64_t)fo);
            if (!ok) return fail("I/O error on '" + d->name + "'");
            int k = r - 2;
            for (; k >= 0; --k) { if (++idx[k] < hi[k]) break; idx[k] = lo[k]; }
            if (k < 0) break;
        }
        return 0;
    }
    // chunked: enumerate the chunks the selection touches
    int64_t c0[kMaxRank], c1[kMaxRank], ci[kMaxRank];
    for (int i = 0; i < r; ++i) { c0[i] = start[i] / d->chunk[i]; c1[i] = (start[i] + count[i] - 1) / d->chunk[i]; ci[i] = c0[i]; }
    std::vector<ChunkJob> jobs;
    for (;;) {
        ChunkJob j;
        j.off.resize(r);
        for (int i = 0; i < r; ++i) j.off[i] = (uint64_t)(ci[i] * d->chunk[i]);
        auto it = d->chunks.find(j.off);
        j.existed = it != d->chunks.end();
        j.addr = j.existed ? it->second : UNDEF;
        if (write && !j.existed) {
            j.addr = alloc(f, d->chunk_bytes);
            d->chunks[j.off] = j.addr;
            d->index_dirty = true;
        }
        jobs.push_back(std::move(j));
        int k = r - 1;
        for (; k >= 0; --k) { if (++ci[k] <= c1[k]) break; ci[k] = c0[k]; }
        if (k < 0) break;
    }
    const bool ok = run_jobs(jobs.size(), [&](size_t ji) -> bool {
        const ChunkJob &j = jobs[ji];
        int64_t corg[kMaxRank], lo[kMaxRank], hi[kMaxRank];
        bool covers = true;  // does the selection cover the whole of the chunk that lies inside the extent?
        bool edge = false;
        for (int i = 0; i < r; ++i) {
            corg[i] = (int64_t)j.off[i];
            const int64_t cend = std::min(corg[i] + d->chunk[i], d->shape[i]);
            lo[i] = std::max(corg[i], start[i]);
            hi[i] = std::min(cend, start[i] + count[i]);
            if (lo[i] > corg[i] || hi[i] < cend) covers = false;
            if (cend < corg[i] + d->chunk[i]) edge = true;
        }
        if (!write && !j.existed) {  // never written: the fill value
            std::vector<uint8_t> pat((size_t)(hi[r - 1] - lo[r - 1]) * (size_t)d->esize);
            fill_pattern(d, pat.data(), pat.size());
            // one row of fill values, reused for every row of the box
            int64_t idx[kMaxRank];
            for (int i = 0; i < r; ++i) idx[i] = lo[i];
            for (;;) {
                int64_t so = 0;
                for (int i = 0; i < r; ++i) so += (idx[i] - start[i]) * sstr[i];
                memcpy(user + so, pat.data(), pat.size());
                int k = r - 2;
                for (; k >= 0; --k) { if (++idx[k] < hi[k]) break; idx[k] = lo[k]; }
                if (k < 0) break;
            }
            return true;
        }
        if (covers && !edge) return whole_chunk_io(f, r, corg, d->chunk, start, sstr, user, d->esize, j.addr, write);
        std::vector<uint8_t> buf(d->chunk_bytes);
        if (!write) {
            if (!pread_all(f, buf.data(), buf.size(), j.addr)) return false;
            copy_box(r, lo, hi, corg, d->chunk, buf.data(), start, sstr, user, d->esize, true);
            return true;
        }
        if (!covers && j.existed) { if (!pread_all(f, buf.data(), buf.size(), j.addr)) return false; }
        else if (!covers || edge) fill_pattern(d, buf.data(), buf.size());
        copy_box(r, lo, hi, corg, d->chunk, buf.data(), start, sstr, user, d->esize, false);
        return pwrite_all(f, buf.data(), buf.size(), j.addr);
    });
    if (!ok) return fail("I/O error on '" + d->name + "' (" + f->path + ")");
    return 0;
}

Dataset *get(File *f, int id) {
    if (!f || id < 0 || (size_t)id >= f->dsets.size()) { g_err = "bad dataset id"; return nullptr; }
    return f->dsets[id].get();
}

std::vector<int> name_order(File *f) {
    std::vector<int> o(f->dsets.size());
    for (size_t i = 0; i < o.size(); ++i) o[i] = (int)i;
    std::sort(o.begin(), o.end(), [&](int a, int b) { return f->dsets[a]->name < f->dsets[b]->name; });
    return o;
}

}  // namespace

extern "C" {

const char *cnngp_h5_last_error(void) { return g_err.c_str(); }

int cnngp_h5_open(const char *path, const char *mode, cnngp_h5 **out) {
    if (!path || !mode || !out) return fail("cnngp_h5_open: NULL argument");
    *out = nullptr;
    const std::string m(mode);
    int flags;
    bool create = false;
    struct stat st;
    const bool exists = stat(path, &st) == 0;
    if (m == "r") flags = O_RDONLY;
    else if (m == "r+") flags = O_RDWR;
    else if (m == "w") { flags = O_RDWR | O_CREAT | O_TRUNC; create = true; }
    else if (m == "w-" || m == "x") { if (exists) return fail(std::string("file exists: ") + path); flags = O_RDWR | O_CREAT | O_EXCL; create = true; }
    else if (m == "a") { flags = O_RDWR | O_CREAT; create = !exists || st.st_size == 0; }
    else return fail("cnngp_h5_open: mode must be r, r+, w, w-, x or a");
    std::unique_ptr<cnngp_h5> f(new cnngp_h5);
    f->path = path;
    f->writable = m != "r";
    f->fd = open(path, flags | O_CLOEXEC, 0644);
    if (f->fd < 0) return fail(std::string("cannot open ") + path + ": " + strerror(errno));
    int rc = create ? create_new(f.get()) : load_existing(f.get());
    if (rc == 0 && create) rc = flush_locked(f.get());
    if (rc) { close(f->fd); return rc; }
    *out = f.release();
    return 0;
}

int cnngp_h5_flush(cnngp_h5 *f) {
    if (!f) return fail("NULL file");
    std::lock_guard<std::mutex> g(f->mu);
    return flush_locked(f);
}

int cnngp_h5_close(cnngp_h5 *f) {
    if (!f) return 0;
    int rc;
    {
        std::lock_guard<std::mutex> g(f->mu);
        rc = flush_locked(f);
        if (close(f->fd) != 0 && rc == 0) rc = fail("close failed: " + f->path);
    }
    delete f;
    return rc;
}

int cnngp_h5_count(cnngp_h5 *f) { return f ? (int)f->dsets.size() : 0; }

int cnngp_h5_name(cnngp_h5 *f, int index, char *buf, int cap) {
    if (!f) return 0;
    std::lock_guard<std::mutex> g(f->mu);
    const std::vector<int> o = name_order(f);
    if (index < 0 || (size_t)index >= o.size()) return 0;
    const std::string &s = f->dsets[o[index]]->name;
    if (buf && cap > 0) { const size_t n = std::min<size_t>(s.size(), (size_t)cap - 1); memcpy(buf, s.data(), n); buf[n] = 0; }
    return (int)s.size() + 1;
}

int cnngp_h5_find(cnngp_h5 *f, const char *name) {
    if (!f || !name) return -1;
    std::lock_guard<std::mutex> g(f->mu);
    for (size_t i = 0; i < f->dsets.size(); ++i) if (f->dsets[i]->name == name) return (int)i;
    return -1;
}

int cnngp_h5_create_dataset(cnngp_h5 *f, const char *name, int rank, const int64_t *shape, const int64_t *maxshape,
                            const int64_t *chunks, int dtype, const void *fill, int *id) {
    if (!f || !name || !shape || !id) return fail("cnngp_h5_create_dataset: NULL argument");
    std::lock_guard<std::mutex> g(f->mu);
    if (!f->writable) return fail("file is open read-only: " + f->path);
    if (rank < 1 || rank > kMaxRank) return fail("rank must be 1.." + std::to_string(kMaxRank));
    if (dtype != 0 && dtype != 1) return fail("dtype must be 0 (float32) or 1 (float64)");
    const std::string nm(name);
    if (nm.empty() || nm.find('/') != std::string::npos) return fail("dataset names are single path components");
    for (const Link &ln : f->links) if (ln.name == nm) return fail("name already exists: " + nm);
    std::unique_ptr<Dataset> d(new Dataset);
    d->name = nm; d->rank = rank; d->dtype = dtype; d->esize = dtype == 0 ? 4 : 8;
    d->chunked = chunks != nullptr;
    uint64_t nelem = 1, cbytes = (uint64_t)d->esize;
    for (int i = 0; i < rank; ++i) {
        d->shape[i] = shape[i];
        d->maxshape[i] = maxshape ? maxshape[i] : shape[i];
        if (shape[i] < 0) return fail("negative extent");
        if (d->maxshape[i] != CNNGP_H5_UNLIMITED && d->maxshape[i] < shape[i]) return fail("maxshape below shape");
        if (!chunks && d->maxshape[i] != shape[i]) return fail("only chunked datasets can be resizable");
        if (chunks) {
            if (chunks[i] < 1 || chunks[i] > 0xffffffffLL) return fail("bad chunk shape");
            if (d->maxshape[i] != CNNGP_H5_UNLIMITED && chunks[i] > d->maxshape[i] && d->maxshape[i] > 0)
                return fail("chunk larger than the maximum extent");
            d->chunk[i] = chunks[i];
            cbytes *= (uint64_t)chunks[i];
        }
        nelem *= (uint64_t)shape[i];
    }
    if (chunks && cbytes > 0xffffffffull) return fail("chunks must be smaller than 4 GiB");
    d->chunk_bytes = chunks ? cbytes : 0;
    if (fill) { memcpy(d->fill, fill, (size_t)d->esize); d->has_fill = true; }
    // object header: dataspace, datatype, fill value, layout
    Buf msgs;
    auto begin_msg = [&](unsigned type, unsigned flags) -> size_t {
        msgs.u16(type); msgs.u16(0); msgs.u8(flags); msgs.zeros(3);
        return msgs.size();
    };
    auto end_msg = [&](size_t body) {
        msgs.pad8();
        const size_t n = msgs.size() - body;
        msgs.b[body - 6] = (uint8_t)(n & 255); msgs.b[body - 5] = (uint8_t)(n >> 8);
    };
    size_t b0 = begin_msg(0x0001, 0);
    msgs.u8(1); msgs.u8((unsigned)rank); msgs.u8(1); msgs.u8(0); msgs.u32(0);
    const size_t dims_rel = msgs.size();
    for (int i = 0; i < rank; ++i) msgs.u64((uint64_t)shape[i]);
    for (int i = 0; i < rank; ++i) msgs.u64(d->maxshape[i] == CNNGP_H5_UNLIMITED ? UNDEF : (uint64_t)d->maxshape[i]);
    end_msg(b0);
    b0 = begin_msg(0x0003, 1);
    msgs.u8(0x11); msgs.u8(0x20); msgs.u8(dtype == 0 ? 31 : 63); msgs.u8(0); msgs.u32((uint32_t)d->esize);
    msgs.u16(0); msgs.u16(dtype == 0 ? 32 : 64);
    msgs.u8(dtype == 0 ? 23 : 52); msgs.u8(dtype == 0 ? 8 : 11); msgs.u8(0); msgs.u8(dtype == 0 ? 23 : 52);
    msgs.u32(dtype == 0 ? 127 : 1023);
    end_msg(b0);
    b0 = begin_msg(0x0005, 1);
    msgs.u8(2); msgs.u8(chunks ? 3 : 1); msgs.u8(2); msgs.u8(1);  // version 2; allocation incremental / early; fill time "if set"; defined
    msgs.u32(fill ? (uint32_t)d->esize : 0);
    if (fill) msgs.raw(d->fill, (size_t)d->esize);
    end_msg(b0);
    b0 = begin_msg(0x0008, 0);
    size_t addr_rel;
    if (chunks) {
        msgs.u8(3); msgs.u8(2); msgs.u8((unsigned)rank + 1);
        addr_rel = msgs.size();
        msgs.u64(UNDEF);
        for (int i = 0; i < rank; ++i) msgs.u32((uint32_t)chunks[i]);
        msgs.u32((uint32_t)d->esize);
    } else {
        msgs.u8(3); msgs.u8(1);
        addr_rel = msgs.size();
        msgs.u64(UNDEF); msgs.u64(nelem * (uint64_t)d->esize);
    }
    end_msg(b0);
    Buf oh;
    oh.u8(1); oh.u8(0); oh.u16(4); oh.u32(1); oh.u32((uint32_t)msgs.size()); oh.u32(0);
    d->ohdr = alloc(f, 16 + msgs.size());
    d->dims_pos = d->ohdr + 16 + dims_rel;
    d->layout_addr_pos = d->ohdr + 16 + addr_rel;
    if (!chunks) {
        d->data_addr = alloc(f, nelem * (uint64_t)d->esize);
        for (int i = 0; i < 8; ++i) msgs.b[addr_rel + i] = (uint8_t)(d->data_addr >> (8 * i));
        // early allocation: the storage holds the fill value from the start
        const size_t blk = 1 << 20;
        std::vector<uint8_t> pat(std::min<uint64_t>(blk, nelem * (uint64_t)d->esize));
        fill_pattern(d.get(), pat.data(), pat.size());
        for (uint64_t o = 0, tot = nelem * (uint64_t)d->esize; o < tot; o += pat.size())
            if (!pwrite_all(f, pat.data(), (size_t)std::min<uint64_t>(pat.size(), tot - o), d->data_addr + o)) return fail("write failed: " + f->path);
    }
    oh.raw(msgs.b.data(), msgs.size());
    if (!pwrite_all(f, oh.b.data(), oh.size(), d->ohdr)) return fail("write failed: " + f->path);
    Link ln;
    ln.name = nm; ln.ohdr = d->ohdr;
    f->links.push_back(ln);
    f->group_dirty = true;
    f->dsets.push_back(std::move(d));
    *id = (int)f->dsets.size() - 1;
    return flush_locked(f);  // the new dataset is visible in the file right away
}

int cnngp_h5_dataset_info(cnngp_h5 *f, int id, cnngp_h5_info *info) {
    if (!f || !info) return fail("NULL argument");
    std::lock_guard<std::mutex> g(f->mu);
    Dataset *d = get(f, id);
    if (!d) return 1;
    memset(info, 0, sizeof *info);
    info->rank = d->rank; info->dtype = d->unsupported.empty() ? d->dtype : -1;
    info->chunked = d->chunked; info->has_fill = d->has_fill;
    for (int i = 0; i < d->rank; ++i) { info->shape[i] = d->shape[i]; info->maxshape[i] = d->maxshape[i]; info->chunks[i] = d->chunk[i]; }
    if (d->has_fill) {
        if (d->esize == 4) { float v; memcpy(&v, d->fill, 4); info->fill = v; }
        else if (d->esize == 8) { double v; memcpy(&v, d->fill, 8); info->fill = v; }
    }
    info->n_chunks_stored = (int64_t)d->chunks.size();
    return 0;
}

int cnngp_h5_write(cnngp_h5 *f, int id, const int64_t *start, const int64_t *count, const void *data) {
    if (!f || !start || !count || !data) return fail("NULL argument");
    std::lock_guard<std::mutex> g(f->mu);
    Dataset *d = get(f, id);
    if (!d) return 1;
    if (!f->writable) return fail("file is open read-only: " + f->path);
    return transfer(f, d, start, count, (uint8_t *)const_cast<void *>(data), true);
}

int cnngp_h5_read(cnngp_h5 *f, int id, const int64_t *start, const int64_t *count, void *data) {
    if (!f || !start || !count || !data) return fail("NULL argument");
    std::lock_guard<std::mutex> g(f->mu);
    Dataset *d = get(f, id);
    if (!d) return 1;
    return transfer(f, d, start, count, (uint8_t *)data, false);
}

int cnngp_h5_write_strided(cnngp_h5 *f, int id, const int64_t *start, const int64_t *count, const void *data,
                           const int64_t *stride_bytes) {
    if (!f || !start || !count || !data || !stride_bytes) return fail("NULL argument");
    std::lock_guard<std::mutex> g(f->mu);
    Dataset *d = get(f, id);
    if (!d) return 1;
    if (!f->writable) return fail("file is open read-only: " + f->path);
    return transfer(f, d, start, count, (uint8_t *)const_cast<void *>(data), true, stride_bytes);
}

int cnngp_h5_read_strided(cnngp_h5 *f, int id, const int64_t *start, const int64_t *count, void *data,
                          const int64_t *stride_bytes) {
    if (!f || !start || !count || !data || !stride_bytes) return fail("NULL argument");
    std::lock_guard<std::mutex> g(f->mu);
    Dataset *d = get(f, id);
    if (!d) return 1;
    return transfer(f, d, start, count, (uint8_t *)data, false, stride_bytes);
}

int cnngp_h5_resize(cnngp_h5 *f, int id, const int64_t *new_shape) {
    if (!f || !new_shape) return fail("NULL argument");
    std::lock_guard<std::mutex> g(f->mu);
    Dataset *d = get(f, id);
    if (!d) return 1;
    if (!f->writable) return fail("file is open read-only: " + f->path);
    if (!d->chunked) return fail("only chunked datasets can be resized");
    for (int i = 0; i < d->rank; ++i)
        if (new_shape[i] < 0 || (d->maxshape[i] != CNNGP_H5_UNLIMITED && new_shape[i] > d->maxshape[i]))
            return fail("new extent outside maxshape in dimension " + std::to_string(i));
    for (int i = 0; i < d->rank; ++i) d->shape[i] = new_shape[i];
    for (auto it = d->chunks.begin(); it != d->chunks.end();) {  // chunks wholly outside the extent leave the index
        bool outside = false;
        for (int i = 0; i < d->rank; ++i) outside = outside || (int64_t)it->first[i] >= d->shape[i];
        if (outside) { it = d->chunks.erase(it); d->index_dirty = true; } else ++it;
    }
    d->dims_dirty = true;
    f->any_dirty = true;
    return 0;
}

int cnngp_h5_merge_nan(cnngp_h5 *dest, int dest_id, cnngp_h5 *src, int src_id) {
    if (!dest || !src) return fail("NULL argument");
    if (dest == src) return fail("merge: source and destination are the same file handle");
    std::lock_guard<std::mutex> g1(dest->mu);
    std::lock_guard<std::mutex> g2(src->mu);
    Dataset *d = get(dest, dest_id), *s = get(src, src_id);
    if (!d || !s) return 1;
    if (!dest->writable) return fail("file is open read-only: " + dest->path);
    if (!d->unsupported.empty() || !s->unsupported.empty()) return fail("merge: unsupported dataset");
    if (!d->chunked || !s->chunked || d->rank != s->rank || d->dtype != s->dtype) return fail("merge: datasets differ in layout or type");
    for (int i = 0; i < d->rank; ++i)
        if (d->shape[i] != s->shape[i] || d->chunk[i] != s->chunk[i]) return fail("merge: datasets differ in shape or chunk shape");
    bool dest_fill_nan = false;
    if (d->has_fill) {
        if (d->esize == 4) { float v; memcpy(&v, d->fill, 4); dest_fill_nan = std::isnan(v); }
        else { double v; memcpy(&v, d->fill, 8); dest_fill_nan = std::isnan(v); }
    }
    std::vector<ChunkJob> jobs;
    for (const auto &kv : s->chunks) {
        ChunkJob j;
        j.off = kv.first;
        auto it = d->chunks.find(kv.first);
        j.existed = it != d->chunks.end();
        if (!j.existed) {
            if (!dest_fill_nan) continue;  // dest holds a non-NaN fill value there: nothing to take
            j.addr = alloc(dest, d->chunk_bytes);
            d->chunks[j.off] = j.addr;
            d->index_dirty = true;
        } else {
            j.addr = it->second;
        }
        jobs.push_back(std::move(j));
    }
    const bool ok = run_jobs(jobs.size(), [&](size_t ji) -> bool {
        const ChunkJob &j = jobs[ji];
        std::vector<uint8_t> sb(d->chunk_bytes), db;
        if (!pread_all(src, sb.data(), sb.size(), s->chunks.find(j.off)->second)) return false;
        if (!j.existed) return pwrite_all(dest, sb.data(), sb.size(), j.addr);
        db.resize(d->chunk_bytes);
        if (!pread_all(dest, db.data(), db.size(), j.addr)) return false;
        bool changed = false;
        if (d->esize == 4) {
            float *a = (float *)db.data();
            const float *b = (const float *)sb.data();
            for (size_t k = 0, n = db.size() / 4; k < n; ++k) if (std::isnan(a[k])) { a[k] = b[k]; changed = true; }
        } else {
            double *a = (double *)db.data();
            const double *b = (const double *)sb.data();
            for (size_t k = 0, n = db.size() / 8; k < n; ++k) if (std::isnan(a[k])) { a[k] = b[k]; changed = true; }
        }
        return !changed || pwrite_all(dest, db.data(), db.size(), j.addr);
    });
    if (!ok) return fail("merge: I/O error");
    return 0;
}

}  // extern "C"
