// unik.hpp -- host-side .unik v5 container codec (stays on the host by the north star).
//
// Restates github.com/shenwei356/unik/v5 v5.0.1 (go.mod:17) as used by unikmer/cmd: NewReader /
// ReadCodeWithTaxid and NewWriter / WriteCode / WriteCodeWithTaxid / SetMaxTaxid / SetGlobalTaxid /
// SetScale / Flush (37 read and 76 write call sites).  The module source is not in the reference tree,
// so the byte layout follows SURVEY.md Appendix A.4 [RECALL] -- FORMAT PARITY IS UNPINNED: files written
// here are self-consistent; interchange with real unikmer files is unverified.  The three uncertain
// items are the named constants below, in this one place.
//
// Batch oriented: a whole payload is decoded into / encoded from arrays (the per-k-mer Read/Write calls
// of the reference are the end-to-end bottleneck).
#pragma once
#include <stdint.h>
#include <zlib.h>

#include <stdexcept>
#include <string>
#include <vector>

namespace unik {

// flags (iota order used at sort.go:205-214, count.go:449-462)
enum : uint32_t { Compact = 1, Canonical = 2, Sorted = 4, IncludeTaxID = 8, Hashed = 16, Scaled = 32 };

constexpr uint8_t MainVersion = 5, MinorVersion = 0;
// ---- uncertain layout items ------------------------------------------
constexpr int kDescLenBytes = 2;   // description length field: u16 BE (v5 "fix reading long description")
constexpr int kReservedBytes = 64; // zero block closing the header
constexpr bool kTaxidLenInHeader = true;  // u8 taxid byte length after the global taxid
// -------------------------------------------------------------------------------------------------------

struct Header {
    int k = 0;
    uint32_t flag = 0;
    uint64_t number = 0;  // 0 = unknown
    uint32_t global_taxid = 0;
    uint8_t taxid_bytes = 4;  // from SetMaxTaxid
    std::string description;
    uint32_t scale = 1;
    uint64_t max_hash = 0;

    bool is(uint32_t f) const { return (flag & f) != 0; }
    bool has_global_taxid() const { return global_taxid != 0; }
    bool has_taxid_info() const { return is(IncludeTaxID) || has_global_taxid(); }
};

inline uint8_t taxid_byte_length(uint32_t max_taxid) {  // SetMaxTaxid; inverse of maxUint32N (util.go:340-342)
    if (max_taxid <= 0xffu) return 1;
    if (max_taxid <= 0xffffu) return 2;
    if (max_taxid <= 0xffffffu) return 3;
    return 4;
}

struct File {
    Header h;
    std::vector<uint64_t> codes;
    std::vector<uint32_t> taxids;  // per code; filled with the global taxid when the file has one; empty if no taxid info
};

// ---- big-endian helpers ------------------------------------------------------------------------------------
inline void put_be(std::vector<uint8_t>& o, uint64_t v, int nbytes) {
    for (int i = nbytes - 1; i >= 0; --i) o.push_back((uint8_t)(v >> (8 * i)));
}
inline uint64_t get_be(const uint8_t* p, int nbytes) {
    uint64_t v = 0;
    for (int i = 0; i < nbytes; ++i) v = (v << 8) | p[i];
    return v;
}
inline int byte_len(uint64_t v) {
    int n = 1;
    while (v >>= 8) ++n;
    return n;
}

inline void encode_header(const Header& h, std::vector<uint8_t>& o) {
    const char magic[8] = {'.', 'u', 'n', 'i', 'k', 'm', 'e', 'r'};
    o.insert(o.end(), magic, magic + 8);
    o.push_back(MainVersion);
    o.push_back(MinorVersion);
    o.push_back((uint8_t)h.k);
    o.push_back(0);
    put_be(o, h.flag, 4);
    put_be(o, h.number, 8);
    put_be(o, h.global_taxid, 4);
    if (kTaxidLenInHeader) o.push_back(h.taxid_bytes);
    if (h.description.size() > 1024) throw std::runtime_error("unik: description longer than 1024 bytes");
    put_be(o, h.description.size(), kDescLenBytes);
    o.insert(o.end(), h.description.begin(), h.description.end());
    put_be(o, h.scale, 4);
    put_be(o, h.max_hash, 8);
    o.insert(o.end(), kReservedBytes, 0);
}

// Writer: whole payload at once.  `taxids` may be empty (dropped silently when IncludeTaxID is off, as
// WriteCodeWithTaxid does: diff.go:593,597).  Sorted payloads must be non-decreasing.
inline std::vector<uint8_t> encode(const Header& h, const uint64_t* codes, const uint32_t* taxids, size_t n) {
    std::vector<uint8_t> o;
    o.reserve(128 + h.description.size() + n * 5);
    encode_header(h, o);
    const bool tx = h.is(IncludeTaxID) && taxids != nullptr;
    if (h.is(IncludeTaxID) && n && !taxids) throw std::runtime_error("unik: IncludeTaxID set but no taxids given");
    const int tb = h.taxid_bytes;
    if (h.is(Sorted)) {
        uint64_t prev = 0;
        size_t i = 0;
        for (; i + 1 < n; i += 2) {
            const uint64_t d1 = codes[i] - prev, d2 = codes[i + 1] - codes[i];
            const int l1 = byte_len(d1), l2 = byte_len(d2);
            o.push_back((uint8_t)(((l1 - 1) << 3) | (l2 - 1)));
            put_be(o, d1, l1);
            put_be(o, d2, l2);
            if (tx) {
                put_be(o, taxids[i], tb);
                put_be(o, taxids[i + 1], tb);
            }
            prev = codes[i + 1];
        }
        if (i < n) {  // trailing odd code flushed by Flush(): marker byte + full code
            o.push_back(128);
            put_be(o, codes[i], 8);
            if (tx) put_be(o, taxids[i], tb);
        }
    } else if (h.is(Compact) && !h.is(Hashed)) {
        const int nb = (h.k + 3) / 4;
        for (size_t i = 0; i < n; ++i) {
            put_be(o, codes[i], nb);
            if (tx) put_be(o, taxids[i], tb);
        }
    } else {
        for (size_t i = 0; i < n; ++i) {
            put_be(o, codes[i], 8);
            if (tx) put_be(o, taxids[i], tb);
        }
    }
    return o;
}

inline size_t decode_header(const uint8_t* p, size_t len, Header& h) {
    const size_t fixed = 8 + 4 + 4 + 8 + 4 + (kTaxidLenInHeader ? 1 : 0) + kDescLenBytes;
    if (len < fixed) throw std::runtime_error("unik: truncated header");
    if (std::string((const char*)p, 8) != ".unikmer") throw std::runtime_error("unik: invalid binary format (magic)");
    if (p[8] != MainVersion) throw std::runtime_error("unik: version mismatch");
    h.k = p[10];
    size_t q = 12;
    h.flag = (uint32_t)get_be(p + q, 4); q += 4;
    h.number = get_be(p + q, 8); q += 8;
    h.global_taxid = (uint32_t)get_be(p + q, 4); q += 4;
    if (kTaxidLenInHeader) h.taxid_bytes = p[q++];
    const size_t dl = (size_t)get_be(p + q, kDescLenBytes); q += kDescLenBytes;
    if (len < q + dl + 12 + kReservedBytes) throw std::runtime_error("unik: truncated header");
    h.description.assign((const char*)p + q, dl); q += dl;
    h.scale = (uint32_t)get_be(p + q, 4); q += 4;
    h.max_hash = get_be(p + q, 8); q += 8;
    q += kReservedBytes;
    if (h.taxid_bytes < 1 || h.taxid_bytes > 4) throw std::runtime_error("unik: bad taxid byte length");
    return q;
}

// Reader: the whole stream (already gunzipped) -> arrays, i.e. ReadCodeWithTaxid until io.EOF.
inline File decode(const uint8_t* p, size_t len, bool ignore_taxid = false) {
    File f;
    size_t q = decode_header(p, len, f.h);
    const Header& h = f.h;
    const bool tx = h.is(IncludeTaxID);
    const int tb = h.taxid_bytes;
    const bool want_tax = !ignore_taxid && h.has_taxid_info();
    if (h.number) f.codes.reserve(h.number);
    auto need = [&](size_t nbytes) {
        if (q + nbytes > len) throw std::runtime_error("unik: truncated payload");
    };
    if (h.is(Sorted)) {
        uint64_t prev = 0;
        while (q < len) {
            const uint8_t ctrl = p[q++];
            if (ctrl & 128) {
                need(8 + (tx ? tb : 0));
                f.codes.push_back(get_be(p + q, 8)); q += 8;
                if (tx) { if (want_tax) f.taxids.push_back((uint32_t)get_be(p + q, tb)); q += tb; }
                continue;
            }
            const int l1 = ((ctrl >> 3) & 7) + 1, l2 = (ctrl & 7) + 1;
            need(l1 + l2 + (tx ? 2 * tb : 0));
            const uint64_t c1 = prev + get_be(p + q, l1); q += l1;
            const uint64_t c2 = c1 + get_be(p + q, l2); q += l2;
            f.codes.push_back(c1);
            f.codes.push_back(c2);
            if (tx) {
                if (want_tax) { f.taxids.push_back((uint32_t)get_be(p + q, tb)); f.taxids.push_back((uint32_t)get_be(p + q + tb, tb)); }
                q += 2 * tb;
            }
            prev = c2;
        }
    } else {
        const int nb = (h.is(Compact) && !h.is(Hashed)) ? (h.k + 3) / 4 : 8;
        const size_t rec = nb + (tx ? tb : 0);
        if ((len - q) % rec) throw std::runtime_error("unik: payload is not a whole number of records");
        const size_t n = (len - q) / rec;
        f.codes.resize(n);
        if (tx && want_tax) f.taxids.resize(n);
        for (size_t i = 0; i < n; ++i, q += rec) {
            f.codes[i] = get_be(p + q, nb);
            if (tx && want_tax) f.taxids[i] = (uint32_t)get_be(p + q + nb, tb);
        }
    }
    if (want_tax && !tx) f.taxids.assign(f.codes.size(), h.global_taxid);  // ReadCodeWithTaxid returns the global taxid
    return f;
}

// ---- files: gzip sniffing like inStream (util-io.go:68-101); writing like outStream (util-io.go:37-66) ----
inline std::vector<uint8_t> slurp(const std::string& path) {
    gzFile g = path == "-" ? gzdopen(0, "rb") : gzopen(path.c_str(), "rb");  // transparently reads plain files too
    if (!g) throw std::runtime_error("cannot open " + path);
    gzbuffer(g, 1 << 20);
    std::vector<uint8_t> buf;
    std::vector<uint8_t> chunk(1 << 22);
    for (;;) {
        int r = gzread(g, chunk.data(), (unsigned)chunk.size());
        if (r < 0) { gzclose(g); throw std::runtime_error("read error in " + path); }
        if (r == 0) break;
        buf.insert(buf.end(), chunk.begin(), chunk.begin() + r);
    }
    gzclose(g);
    return buf;
}

inline File read_file(const std::string& path, bool ignore_taxid = false) {
    std::vector<uint8_t> raw = slurp(path);
    try {
        return decode(raw.data(), raw.size(), ignore_taxid);
    } catch (const std::exception& e) {
        throw std::runtime_error(path + ": " + e.what());
    }
}

inline void spill(const std::string& path, const std::vector<uint8_t>& bytes, bool compress, int level) {
    if (compress) {
        char mode[8];
        snprintf(mode, sizeof mode, "wb%d", level < 0 ? 6 : (level > 9 ? 9 : level));
        gzFile g = path == "-" ? gzdopen(1, mode) : gzopen(path.c_str(), mode);
        if (!g) throw std::runtime_error("cannot create " + path);
        gzbuffer(g, 1 << 20);
        size_t off = 0;
        while (off < bytes.size()) {
            unsigned n = (unsigned)std::min<size_t>(bytes.size() - off, 1u << 30);
            if (gzwrite(g, bytes.data() + off, n) != (int)n) { gzclose(g); throw std::runtime_error("write error in " + path); }
            off += n;
        }
        gzclose(g);
    } else {
        FILE* fh = path == "-" ? stdout : fopen(path.c_str(), "wb");
        if (!fh) throw std::runtime_error("cannot create " + path);
        if (fwrite(bytes.data(), 1, bytes.size(), fh) != bytes.size()) throw std::runtime_error("write error in " + path);
        if (fh != stdout) fclose(fh);
    }
}

inline void write_file(const std::string& path, const Header& h, const uint64_t* codes, const uint32_t* taxids, size_t n,
                       bool compress, int level) {
    spill(path, encode(h, codes, taxids, n), compress, level);
}

}  // namespace unik
