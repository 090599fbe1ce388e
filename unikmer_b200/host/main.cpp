// unikmer-b200 -- host program over libukm.so with the CLI surface of
// `unikmer {count,sort,union,inter,diff,common}` (+ `view`, `info` for inspection).
//
// The reference's host side is Go (cobra commands in unikmer/cmd/*.go); no Go toolchain exists in this
// image, so the same command layer is written in C++ over the C ABI.  It keeps, per command, the flag
// names, input checks, header flags (`mode`), `Number` and `MaxTaxid` rules of the reference (SURVEY.md
// Appendix D; lines cited at each command) and reads/writes `.unik` through host/unik.hpp.  Compute goes
// through libukm only (no CPU implementation of any operation here).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <dirent.h>
#include <errno.h>
#include <sys/stat.h>

#include <algorithm>
#include <fstream>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "../../include/ukm.h"
#include "fastx.hpp"
#include "unik.hpp"

namespace {

[[noreturn]] void die(const char* fmt, ...) {  // checkError: log + os.Exit(-1) (util-cli.go:39-44)
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "[ERRO] ");
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
    exit(255);
}

struct Options {  // root.go:98-111 persistent flags
    int threads = 4;
    bool verbose = false;
    bool compress = true;
    int compression_level = -1;
    bool compact = false;
    std::string infile_list;
    uint32_t max_taxid = 4294967295u;
    bool ignore_taxid = false;
    std::string data_dir;
    int device = 0;
    std::string out = "-";
    // per-command
    int k = 0;
    bool canonical = false, hashed = false, sorted = false, circular = false, unique = false, repeated = false;
    uint32_t scale = 1, taxid = 0;
    int minimizer_w = 0;  // count -W
    size_t chunk_size = 0;  // sort / split -m (k-mers per chunk; util.go:291 ParseByteSize)
    std::string out_dir;    // split -O
    bool is_dir = false;    // merge -D
    bool force = false;     // split --force
    bool mix_taxid = false, compare_taxid = false;
    int number = 0;
    double proportion = 1.0;
    bool show_taxid = false, show_code = false;
    std::vector<std::string> files;
};

void logi(const Options& o, const char* fmt, ...) {
    if (!o.verbose) return;
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "[INFO] ");
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
}

// util.go:291-335 ParseByteSize: a number with an optional B/K/M/G suffix (powers of 1024)
size_t parse_byte_size(const char* v) {
    std::string val(v);
    while (!val.empty() && isspace((unsigned char)val.back())) val.pop_back();
    if (val.empty()) return 0;
    double unit = 1;
    bool has_unit = true;
    switch (val.back()) {
        case 'B': case 'b': unit = 1; break;
        case 'K': case 'k': unit = 1024.0; break;
        case 'M': case 'm': unit = 1024.0 * 1024.0; break;
        case 'G': case 'g': unit = 1024.0 * 1024.0 * 1024.0; break;
        default: has_unit = false;
    }
    if (has_unit) val.pop_back();
    if (val.empty()) return 0;
    char* end = nullptr;
    const double x = strtod(val.c_str(), &end);
    if (end == val.c_str() || *end) die("invalid byte size: %s", v);
    return x < 0 ? 0 : (size_t)(x * unit);
}

// ---- argument parsing: short/long flags of the six commands ------------------------------------------
Options parse(int argc, char** argv, int first) {
    Options o;
    const char* env = getenv("UNIKMER_DB");  // util.go:74-83
    if (env) o.data_dir = env;
    else if (getenv("HOME")) o.data_dir = std::string(getenv("HOME")) + "/.unikmer";
    auto need = [&](int& i) -> const char* {
        if (i + 1 >= argc) die("flag needs an argument: %s", argv[i]);
        return argv[++i];
    };
    for (int i = first; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "-j" || a == "--threads") o.threads = atoi(need(i));
        else if (a == "--verbose") o.verbose = true;
        else if (a == "-C" || a == "--no-compress") o.compress = false;
        else if (a == "--compression-level") o.compression_level = atoi(need(i));
        else if (a == "-c" || a == "--compact") o.compact = true;
        else if (a == "-i" || a == "--infile-list") o.infile_list = need(i);
        else if (a == "--max-taxid") o.max_taxid = (uint32_t)strtoul(need(i), nullptr, 10);
        else if (a == "-I" || a == "--ignore-taxid") o.ignore_taxid = true;
        else if (a == "--data-dir") o.data_dir = need(i);
        else if (a == "--device") o.device = atoi(need(i));
        else if (a == "-o" || a == "--out-prefix" || a == "--out-file") o.out = need(i);
        else if (a == "-k" || a == "--kmer-len") o.k = atoi(need(i));
        else if (a == "-K" || a == "--canonical") o.canonical = true;
        else if (a == "-H" || a == "--hash") o.hashed = true;
        else if (a == "-s" || a == "--sort") o.sorted = true;
        else if (a == "--circular") o.circular = true;
        else if (a == "-u" || a == "--unique") o.unique = true;
        else if (a == "-d" || a == "--repeated") o.repeated = true;
        else if ((a == "-D" && std::string(argv[1]) != "merge") || a == "--scale") o.scale = (uint32_t)strtoul(need(i), nullptr, 10);
        else if (a == "-W" || a == "--minimizer-w") o.minimizer_w = atoi(need(i));
        else if (a == "-S" || a == "--syncmer-s") die("count -S (closed syncmers) is not supported by this build: bio/sketches is not pinned by the reference tree");
        else if (a == "-t" || a == "--taxid" || a == "--compare-taxid" || a == "--show-taxid") {
            // `count -t <taxid>` takes a value; `diff -t` / `view -t` are switches
            const std::string cmd = argv[1];
            if (cmd == "count") o.taxid = (uint32_t)strtoul(need(i), nullptr, 10);
            else if (cmd == "diff") o.compare_taxid = true;
            else o.show_taxid = true;
        } else if (a == "-m" && (std::string(argv[1]) == "sort" || std::string(argv[1]) == "split")) o.chunk_size = parse_byte_size(need(i));
        else if (a == "--chunk-size") o.chunk_size = parse_byte_size(need(i));
        else if (a == "-O" || a == "--out-dir") o.out_dir = need(i);
        else if (a == "--is-dir" || (a == "-D" && std::string(argv[1]) == "merge")) o.is_dir = true;
        else if (a == "--force") o.force = true;
        else if (a == "-M" || a == "--max-open-files") (void)need(i);  // merge: no open-file limit on the device path
        else if (a == "-m" || a == "--mix-taxid") o.mix_taxid = true;
        else if (a == "-n" || a == "--number") o.number = atoi(need(i));
        else if (a == "-p" || a == "--proportion") o.proportion = atof(need(i));
        else if (a == "-N" || a == "--show-code") o.show_code = true;
        else if (a.size() > 1 && a[0] == '-' && a != "-") die("unknown flag: %s", a.c_str());
        else o.files.push_back(a);
    }
    if (!o.infile_list.empty()) {  // util-cli.go:192-264
        std::ifstream fh(o.infile_list);
        if (!fh) die("cannot read file list %s", o.infile_list.c_str());
        std::string line;
        while (std::getline(fh, line))
            if (!line.empty()) o.files.push_back(line);
    }
    if (o.files.empty()) o.files.push_back("-");
    return o;
}

std::string out_name(const std::string& prefix) {  // sort.go:110-113 etc.
    if (prefix == "-") return prefix;
    const std::string ext = ".unik";
    if (prefix.size() >= ext.size() && prefix.compare(prefix.size() - ext.size(), ext.size(), ext) == 0) return prefix;
    return prefix + ext;
}

// ---- taxonomy: loadTaxonomy (util.go:119-171) ------------------------------------------------------------------
struct Taxonomy {
    std::vector<uint32_t> parent, merged_from, merged_to;
    uint32_t max_taxid = 0;
};

Taxonomy load_taxonomy(Options& o, ukm_ctx* ctx) {
    if (o.data_dir.empty()) die("taxonomy data directory not set (--data-dir or $UNIKMER_DB)");
    const std::string nodes = o.data_dir + "/nodes.dmp";
    std::ifstream fh(nodes);
    if (!fh) die("taxonomy file not found: %s", nodes.c_str());
    logi(o, "loading Taxonomy from: %s", o.data_dir.c_str());
    Taxonomy t;
    std::vector<std::pair<uint32_t, uint32_t>> edges;
    std::string line;
    while (std::getline(fh, line)) {
        // "child\t|\tparent\t|\trank..."
        size_t p1 = line.find("\t|\t");
        if (p1 == std::string::npos) continue;
        size_t p2 = line.find("\t|", p1 + 3);
        uint32_t child = (uint32_t)strtoul(line.substr(0, p1).c_str(), nullptr, 10);
        uint32_t par = (uint32_t)strtoul(line.substr(p1 + 3, p2 == std::string::npos ? std::string::npos : p2 - p1 - 3).c_str(), nullptr, 10);
        edges.emplace_back(child, par);
        t.max_taxid = std::max(t.max_taxid, child);
    }
    t.parent.assign((size_t)t.max_taxid + 1, 0);
    for (auto& e : edges)
        if (e.second <= t.max_taxid) t.parent[e.first] = e.second;
    std::ifstream mh(o.data_dir + "/merged.dmp");  // loaded only if it exists (util.go:150-159)
    while (mh && std::getline(mh, line)) {
        size_t p1 = line.find("\t|\t");
        if (p1 == std::string::npos) continue;
        t.merged_from.push_back((uint32_t)strtoul(line.substr(0, p1).c_str(), nullptr, 10));
        t.merged_to.push_back((uint32_t)strtoul(line.substr(p1 + 3).c_str(), nullptr, 10));
    }
    if (ukm_set_taxonomy(ctx, t.parent.data(), t.parent.size(), t.merged_from.data(), t.merged_to.data(), t.merged_from.size()) != UKM_OK)
        die("%s", ukm_last_error(ctx));
    o.max_taxid = t.max_taxid;  // util.go:169: opt.MaxTaxid = t.MaxTaxid()
    logi(o, "%zu nodes loaded", edges.size());
    return t;
}

// ---- FASTA/FASTQ: bio/seqio/fastx view (line breaks stripped, case kept) ------------------------------------------
void read_fastx(const std::string& path, std::vector<uint8_t>& bases, std::vector<uint64_t>& rec_off) {
    std::vector<uint8_t> raw = unik::slurp(path);
    const std::string err = fastx::parse(raw.data(), raw.size(), bases, rec_off);
    if (!err.empty()) die("%s: %s", path.c_str(), err.c_str());
}

// ---- shared command plumbing -----------------------------------------------------------------------------------------
struct Inputs {
    std::vector<unik::File> files;
    int k = 0;
    bool canonical = false, hashed = false, has_taxid = false;
};

void check_compat(const unik::Header& a, const unik::Header& b, const std::string& file) {  // util-binary-file.go:31-44
    if (a.k != b.k) die("K (%d) of binary file '%s' not equal to previous K (%d)", b.k, file.c_str(), a.k);
    if (a.is(unik::Canonical) != b.is(unik::Canonical)) die("'canonical' flags not consistent: %s", file.c_str());
    if (a.is(unik::Hashed) != b.is(unik::Hashed)) die("'hashed' flags not consistent: %s", file.c_str());
    if (a.is(unik::Scaled) != b.is(unik::Scaled)) die("'scaled' flags not consistent: %s", file.c_str());
}

Inputs load_inputs(const Options& o, bool require_sorted, bool require_first_sorted, bool same_taxid_presence) {
    Inputs in;
    for (size_t i = 0; i < o.files.size(); ++i) {
        logi(o, "reading file (%zu/%zu): %s", i + 1, o.files.size(), o.files[i].c_str());
        in.files.push_back(unik::read_file(o.files[i], o.ignore_taxid));
        const unik::Header& h = in.files.back().h;
        if ((require_sorted || (require_first_sorted && i == 0)) && !h.is(unik::Sorted))
            die("input should be sorted: %s", o.files[i].c_str());  // inter.go:139, diff.go:115, common.go:166
        if (i == 0) {
            in.k = h.k;
            in.canonical = h.is(unik::Canonical);
            in.hashed = h.is(unik::Hashed);
            in.has_taxid = !o.ignore_taxid && h.has_taxid_info();
        } else {
            check_compat(in.files[0].h, h, o.files[i]);
            if (same_taxid_presence && !o.ignore_taxid && h.has_taxid_info() != in.has_taxid)
                die(h.has_taxid_info() ? "taxid information not found in previous files, but found in this: %s"
                                       : "taxid information found in previous files, but missing in this: %s",
                    o.files[i].c_str());
        }
    }
    return in;
}

std::vector<ukm_span> spans_of(Inputs& in, bool with_taxid) {
    std::vector<ukm_span> sp(in.files.size());
    for (size_t i = 0; i < in.files.size(); ++i) {
        unik::File& f = in.files[i];
        memset(&sp[i], 0, sizeof sp[i]);
        sp[i].keys = f.codes.data();
        sp[i].taxids = (with_taxid && f.taxids.size() == f.codes.size() && !f.codes.empty()) ? f.taxids.data() : nullptr;
        sp[i].n = sp[i].cap = f.codes.size();
        sp[i].where = UKM_HOST;
        sp[i].sorted = f.h.is(unik::Sorted) ? 1 : 0;
    }
    return sp;
}

struct Result {
    std::vector<uint64_t> codes;
    std::vector<uint32_t> taxids;
    ukm_span span;
    Result(size_t cap, bool tax) : codes(cap + 2), taxids(tax ? cap + 2 : 0) {
        memset(&span, 0, sizeof span);
        span.keys = codes.data();
        span.taxids = tax ? taxids.data() : nullptr;
        span.cap = cap + 2;
        span.where = UKM_HOST;
    }
    size_t n() const { return span.n; }
};

uint32_t base_mode(const Inputs& in, bool sorted, bool include_taxid) {
    uint32_t m = 0;
    if (sorted) m |= unik::Sorted;
    if (in.canonical) m |= unik::Canonical;
    if (include_taxid) m |= unik::IncludeTaxID;
    if (in.hashed) m |= unik::Hashed;
    return m;
}

void write_result(const Options& o, int k, uint32_t mode, uint64_t number, uint32_t max_taxid, const uint64_t* codes,
                  const uint32_t* taxids, size_t n, uint32_t global_taxid = 0, uint32_t scale = 1, uint64_t max_hash = 0) {
    unik::Header h;
    h.k = k;
    h.flag = mode;
    h.number = number;
    h.global_taxid = global_taxid;
    h.taxid_bytes = unik::taxid_byte_length(max_taxid);
    h.scale = scale;
    h.max_hash = max_hash;
    const std::string out = out_name(o.out);
    unik::write_file(out, h, codes, (mode & unik::IncludeTaxID) ? taxids : nullptr, n, o.compress && out != "-" ? true : (o.compress && out == "-"),
                     o.compression_level);
    logi(o, "%zu k-mers saved to %s", n, out.c_str());
}

void copy_single(const Options& o) {  // union.go:97-112, inter.go:96-120, common.go:123-147: byte copy, re-compressed
    std::vector<uint8_t> raw = unik::slurp(o.files[0]);
    unik::spill(out_name(o.out), raw, o.compress, o.compression_level);
}

ukm_ctx* open_ctx(const Options& o) {
    ukm_ctx* ctx = ukm_create(o.device);
    if (!ctx) die("%s", ukm_last_error(nullptr));
    return ctx;
}

#define CHECK(ctx, call)                                \
    do {                                                \
        if ((call) != UKM_OK) die("%s", ukm_last_error(ctx)); \
    } while (0)

// ---- commands --------------------------------------------------------------------------------------------------------------
int cmd_count(Options o) {  // count.go:56-602
    if (o.k < 1) die("k-mer length (-k) needed");
    if (o.unique || o.repeated) die("count -u/-d is not supported by this build");
    bool hashed = o.hashed, scaled = false;
    uint64_t max_hash = 0;
    if (o.k > 32 && !hashed) hashed = true;  // count.go:81-87: k > 32 switches hashing on
    if (o.k > 64) die("k-mer size (%d) should be <= 64", o.k);
    if (o.scale > 1) {  // count.go:96-99
        hashed = true;
        scaled = true;
        max_hash = (uint64_t)((double)(~0ull) / (double)o.scale);
    }
    std::vector<uint8_t> bases;
    std::vector<uint64_t> rec_off{0};
    for (auto& f : o.files) {
        logi(o, "reading sequence file: %s", f.c_str());
        read_fastx(f, bases, rec_off);
    }
    if (o.minimizer_w > 1) hashed = true;  // count.go:105-109: -W switches -H on
    ukm_ctx* ctx = open_ctx(o);
    unsigned flags = (o.canonical ? UKM_F_CANONICAL : 0) | (hashed ? UKM_F_HASHED : 0) | (o.circular ? UKM_F_CIRCULAR : 0) |
                     (scaled ? UKM_F_SCALED : 0);
    Result r(bases.size() + 1, false);
    if (o.minimizer_w > 0)  // count.go:316-317, 358-359
        CHECK(ctx, ukm_count_minimizer(ctx, bases.data(), rec_off.data(), rec_off.size() - 1, o.k, o.minimizer_w, flags, max_hash, UKM_HOST,
                                       &r.span));
    else
        CHECK(ctx, ukm_count_seq(ctx, bases.data(), rec_off.data(), rec_off.size() - 1, o.k, flags, max_hash, UKM_HOST, &r.span));
    uint32_t mode = 0;  // count.go:449-462
    if (o.sorted) mode |= unik::Sorted;
    else if (o.compact && !hashed) mode |= unik::Compact;
    if (o.canonical) mode |= unik::Canonical;
    if (hashed) mode |= unik::Hashed;
    if (scaled) mode |= unik::Scaled;  // SetScale (count.go:469-471)
    write_result(o, o.k, mode, r.n(), o.max_taxid, r.codes.data(), nullptr, r.n(), o.taxid, scaled ? o.scale : 1, max_hash);
    ukm_destroy(ctx);
    return 0;
}

int cmd_sort(Options o) {  // sort.go:64-580 (in-memory path; -m chunks are a host memory knob the device path does not need)
    if (o.unique && o.repeated) die("flag -u/--unique overides -d/--repeated");
    Inputs in = load_inputs(o, false, false, false);
    ukm_ctx* ctx = open_ctx(o);
    const bool tax = in.has_taxid;
    if (tax && (o.unique || o.repeated)) load_taxonomy(o, ctx);  // sort.go:198-200
    size_t total = 0;
    for (auto& f : in.files) total += f.codes.size();
    std::vector<uint64_t> keys;
    std::vector<uint32_t> tx;
    keys.reserve(total);
    for (auto& f : in.files) {
        keys.insert(keys.end(), f.codes.begin(), f.codes.end());
        if (tax) {
            if (f.taxids.size() == f.codes.size()) tx.insert(tx.end(), f.taxids.begin(), f.taxids.end());
            else tx.insert(tx.end(), f.codes.size(), 0u);
        }
    }
    logi(o, "sorting %zu k-mers", total);
    const int key_bits = in.hashed ? 64 : 2 * in.k;
    if (tax) CHECK(ctx, ukm_sort_pairs(ctx, keys.data(), tx.data(), total, key_bits, UKM_HOST));
    else CHECK(ctx, ukm_sort_u64(ctx, keys.data(), total, key_bits, UKM_HOST));
    const int mode_fold = o.unique ? UKM_FOLD_UNIQUE : (o.repeated ? UKM_FOLD_REPEATED_FINAL : UKM_FOLD_PLAIN);
    const uint32_t mode = base_mode(in, true, tax);  // sort.go:205-214
    if (mode_fold == UKM_FOLD_PLAIN) {
        write_result(o, in.k, mode, total, o.max_taxid, keys.data(), tax ? tx.data() : nullptr, total);  // Number set: sort.go:534,567
    } else {
        ukm_span s;
        memset(&s, 0, sizeof s);
        s.keys = keys.data();
        s.taxids = tax ? tx.data() : nullptr;
        s.n = s.cap = total;
        s.where = UKM_HOST;
        s.sorted = 1;
        Result r(total, tax);
        CHECK(ctx, ukm_fold_sorted(ctx, mode_fold, &s, tax ? UKM_F_TAXID : 0, &r.span));
        write_result(o, in.k, mode, 0, o.max_taxid, r.codes.data(), tax ? r.taxids.data() : nullptr, r.n());  // Number unset
    }
    ukm_destroy(ctx);
    return 0;
}

// concatenated codes (+ taxids) of all inputs, unsorted: what sort / split accumulate before sorting (sort.go:226-239)
void gather(const Inputs& in, bool tax, std::vector<uint64_t>& keys, std::vector<uint32_t>& tx) {
    size_t total = 0;
    for (auto& f : in.files) total += f.codes.size();
    keys.reserve(total);
    for (auto& f : in.files) {
        keys.insert(keys.end(), f.codes.begin(), f.codes.end());
        if (tax) {
            if (f.taxids.size() == f.codes.size()) tx.insert(tx.end(), f.taxids.begin(), f.taxids.end());
            else tx.insert(tx.end(), f.codes.size(), 0u);
        }
    }
}

// split (split.go:56-410): stage 1 of the external sort as a command -- chunks of -m k-mers, each sorted on the device
// and written with the chunk fold of dumpCodes2File / dumpCodesTaxids2File (util-sort.go:35-190) to <out-dir>/chunk_NNN.unik.
// (The reference's last chunk skips the taxid sort, quirk B-11; here every chunk is sorted.)
int cmd_split(Options o) {
    if (o.unique && o.repeated) die("flag -u/--unique overides -d/--repeated");
    Inputs in = load_inputs(o, false, false, false);
    std::string dir = o.out_dir.empty() ? (o.files[0] == "-" ? std::string("stdin.split") : o.files[0] + ".split") : o.out_dir;  // split.go:98-103
    if (mkdir(dir.c_str(), 0777) != 0 && errno != EEXIST) die("cannot create %s", dir.c_str());
    ukm_ctx* ctx = open_ctx(o);
    const bool tax = in.has_taxid;
    if (tax && (o.unique || o.repeated)) load_taxonomy(o, ctx);
    std::vector<uint64_t> keys;
    std::vector<uint32_t> tx;
    gather(in, tax, keys, tx);
    const size_t total = keys.size();
    const size_t chunk = o.chunk_size ? o.chunk_size : (total ? total : 1);
    const int key_bits = in.hashed ? 64 : 2 * in.k;
    const int fold = o.unique ? UKM_FOLD_UNIQUE : (o.repeated ? UKM_FOLD_REPEATED_CHUNK : UKM_FOLD_PLAIN);
    const uint32_t mode = base_mode(in, true, tax);
    size_t n_chunks = 0, n_saved = 0;
    for (size_t off = 0; off < total || (total == 0 && n_chunks == 0); off += chunk) {
        const size_t m = std::min(chunk, total - off);
        uint64_t* k = keys.data() + off;
        uint32_t* t = tax ? tx.data() + off : nullptr;
        if (m > 1) {
            if (tax) CHECK(ctx, ukm_sort_pairs(ctx, k, t, m, key_bits, UKM_HOST));
            else CHECK(ctx, ukm_sort_u64(ctx, k, m, key_bits, UKM_HOST));
        }
        char name[64];
        snprintf(name, sizeof name, "/chunk_%03zu.unik", n_chunks);  // util-sort.go:192-194
        Options oc = o;
        oc.out = dir + name;
        if (fold == UKM_FOLD_PLAIN) {
            write_result(oc, in.k, mode, tax ? m : 0, o.max_taxid, k, t, m);  // Number: util-sort.go:181 (taxid chunks only)
            n_saved += m;
        } else {
            ukm_span s;
            memset(&s, 0, sizeof s);
            s.keys = k;
            s.taxids = t;
            s.n = s.cap = m;
            s.where = UKM_HOST;
            s.sorted = 1;
            Result r(m + 2, tax);
            CHECK(ctx, ukm_fold_sorted(ctx, fold, &s, tax ? UKM_F_TAXID : 0, &r.span));
            write_result(oc, in.k, mode, 0, o.max_taxid, r.codes.data(), tax ? r.taxids.data() : nullptr, r.n());
            n_saved += r.n();
        }
        ++n_chunks;
        if (total == 0) break;
    }
    logi(o, "%zu chunk files with total %zu k-mers saved to dir: %s", n_chunks, n_saved, dir.c_str());
    ukm_destroy(ctx);
    return 0;
}

// merge (merge.go:53-335): k-way merge of sorted chunk files with the plain / -u / -d folds of the FINAL round of
// mergeChunksFile (util-sort.go:227-606).  -D: the arguments are directories holding chunk_NNN.unik files (merge.go:78-132).
int cmd_merge(Options o) {
    if (o.unique && o.repeated) die("flag -u/--unique overides -d/--repeated");
    if (o.is_dir) {
        std::vector<std::string> files;
        for (auto& d : o.files) {
            DIR* dh = opendir(d.c_str());
            if (!dh) die("cannot read directory %s", d.c_str());
            std::vector<std::string> found;
            while (dirent* e = readdir(dh)) {
                const std::string n = e->d_name;  // default pattern ^chunk_\d+\.unik$ (merge.go:344)
                if (n.size() > 11 && n.compare(0, 6, "chunk_") == 0 && n.compare(n.size() - 5, 5, ".unik") == 0 &&
                    n.find_first_not_of("0123456789", 6) == n.size() - 5)
                    found.push_back(d + "/" + n);
            }
            closedir(dh);
            std::sort(found.begin(), found.end());
            files.insert(files.end(), found.begin(), found.end());
        }
        if (files.empty()) {
            fprintf(stderr, "[WARN] no valid chunk files given\n");
            return 0;
        }
        o.files = files;
    }
    Inputs in = load_inputs(o, true, false, true);  // merge.go:168-170: input files should be sorted
    ukm_ctx* ctx = open_ctx(o);
    const bool tax = in.has_taxid;
    if (tax) load_taxonomy(o, ctx);  // merge.go:196-201
    std::vector<ukm_span> sp = spans_of(in, tax);
    size_t total = 0;
    for (auto& s : sp) total += s.n;
    const int fold = o.unique ? UKM_FOLD_UNIQUE : (o.repeated ? UKM_FOLD_REPEATED_FINAL : UKM_FOLD_PLAIN);
    Result r(total + 2, tax);
    CHECK(ctx, ukm_merge_sorted(ctx, fold, sp.data(), (int)sp.size(), tax ? UKM_F_TAXID : 0, &r.span));
    write_result(o, in.k, base_mode(in, true, tax), 0, o.max_taxid, r.codes.data(), tax ? r.taxids.data() : nullptr, r.n());  // Number unset
    ukm_destroy(ctx);
    return 0;
}

int cmd_union(Options o) {  // union.go:53-312
    if (o.files.size() == 1) { copy_single(o); return 0; }
    Inputs in = load_inputs(o, false, false, true);
    ukm_ctx* ctx = open_ctx(o);
    const bool tax = in.has_taxid;
    if (tax) load_taxonomy(o, ctx);  // union.go:139-153
    // the engine needs sorted duplicate-free streams; union.go accepts any file, so unsorted ones are sorted + deduplicated first
    for (auto& f : in.files) {
        if (f.h.is(unik::Sorted)) continue;
        const int key_bits = in.hashed ? 64 : 2 * in.k;
        if (tax && f.taxids.size() == f.codes.size()) CHECK(ctx, ukm_sort_pairs(ctx, f.codes.data(), f.taxids.data(), f.codes.size(), key_bits, UKM_HOST));
        else CHECK(ctx, ukm_sort_u64(ctx, f.codes.data(), f.codes.size(), key_bits, UKM_HOST));
        ukm_span s;
        memset(&s, 0, sizeof s);
        s.keys = f.codes.data();
        s.taxids = (tax && f.taxids.size() == f.codes.size() && !f.codes.empty()) ? f.taxids.data() : nullptr;
        s.n = s.cap = f.codes.size();
        s.where = UKM_HOST;
        Result r(f.codes.size(), s.taxids != nullptr);
        CHECK(ctx, ukm_fold_sorted(ctx, UKM_FOLD_UNIQUE, &s, s.taxids ? UKM_F_TAXID : 0, &r.span));
        f.codes.assign(r.codes.begin(), r.codes.begin() + r.n());
        if (s.taxids) f.taxids.assign(r.taxids.begin(), r.taxids.begin() + r.n());
    }
    std::vector<ukm_span> sp = spans_of(in, tax);
    size_t total = 0;
    for (auto& s : sp) total += s.n;
    Result r(total, tax);
    CHECK(ctx, ukm_union(ctx, sp.data(), (int)sp.size(), tax ? UKM_F_TAXID : 0, &r.span));
    uint32_t mode = base_mode(in, o.sorted, tax);  // union.go:221-235
    if (!o.sorted && o.compact && !in.hashed) mode |= unik::Compact;
    const uint64_t number = (o.sorted || tax) ? r.n() : 0;  // union.go:245
    write_result(o, in.k, mode, number, o.max_taxid, r.codes.data(), tax ? r.taxids.data() : nullptr, r.n());
    ukm_destroy(ctx);
    return 0;
}

int cmd_inter(Options o) {  // inter.go:54-357
    if (o.files.size() == 1) { copy_single(o); return 0; }
    Inputs in = load_inputs(o, true, false, !o.mix_taxid);
    bool any_tax = false;
    for (auto& f : in.files) any_tax |= !o.ignore_taxid && f.h.has_taxid_info();
    const bool mix = o.mix_taxid && any_tax;  // inter.go:148-155: hasMixTaxid
    ukm_ctx* ctx = open_ctx(o);
    const bool tax = in.has_taxid && !mix;
    if (tax || mix) load_taxonomy(o, ctx);
    std::vector<ukm_span> sp = spans_of(in, tax || mix);
    Result r(sp[0].n, tax || mix);
    const unsigned flags = (tax ? UKM_F_TAXID : 0) | (mix ? UKM_F_MIX_TAXID : 0);
    CHECK(ctx, ukm_inter(ctx, sp.data(), (int)sp.size(), flags, &r.span));
    if (r.n() == 0) logi(o, "no intersection found");
    write_result(o, in.k, base_mode(in, true, tax || mix), r.n(), o.max_taxid, r.codes.data(), (tax || mix) ? r.taxids.data() : nullptr,
                 r.n());  // inter.go:324-340
    ukm_destroy(ctx);
    return 0;
}

int cmd_diff(Options o) {  // diff.go:58-606
    Inputs in = load_inputs(o, false, true, o.compare_taxid);
    ukm_ctx* ctx = open_ctx(o);
    const bool tax = in.has_taxid;
    if (o.compare_taxid && tax) load_taxonomy(o, ctx);  // diff.go:124-129
    else if (o.compare_taxid) fprintf(stderr, "[WARN] no taxid information found in the first file, flag -t/--compare-taxid ignored\n");
    // the sender skips any subject whose path equals files[0] (diff.go:461-478)
    std::vector<ukm_span> all = spans_of(in, tax);
    std::vector<ukm_span> sp{all[0]};
    for (size_t i = 1; i < all.size(); ++i)
        if (o.files[i] != o.files[0]) sp.push_back(all[i]);
    Result r(sp[0].n, tax);
    const unsigned flags = (tax ? UKM_F_TAXID : 0) | ((o.compare_taxid && tax) ? UKM_F_COMPARE_TAXID : 0);
    if (sp.size() == 1) {
        // no subject at all: no worker ever stores a map, m0 stays nil => empty output (quirk B-5, diff.go:485-523,570-572)
        r.span.n = 0;
    } else {
        CHECK(ctx, ukm_diff(ctx, sp.data(), (int)sp.size(), flags, &r.span));
    }
    uint32_t mode = base_mode(in, o.sorted, tax);  // diff.go:546-560
    if (!o.sorted && o.compact && !in.hashed) mode |= unik::Compact;
    const uint64_t number = o.sorted ? r.n() : 0;  // diff.go:566-568
    write_result(o, in.k, mode, number, o.max_taxid, r.codes.data(), tax ? r.taxids.data() : nullptr, r.n());
    ukm_destroy(ctx);
    return 0;
}

int cmd_common(Options o) {  // common.go:59-361
    const size_t nfiles = o.files.size();
    if (nfiles > 65535) die("at most 65535 files supported");  // common.go:75-77
    uint16_t threshold;  // common.go:93-105
    if (o.number > 0) {
        if ((size_t)o.number > nfiles) die("value of -n/--number (%d) should be <= number of files (%zu)", o.number, nfiles);
        threshold = (uint16_t)o.number;
    } else {
        if (o.proportion <= 0 || o.proportion > 1) die("value of -p/--proportion should be in (0, 1]");
        threshold = (uint16_t)((double)nfiles * o.proportion);
    }
    if (nfiles == 1) { copy_single(o); return 0; }
    Inputs in = load_inputs(o, true, false, !o.mix_taxid);
    bool any_tax = false;
    for (auto& f : in.files) any_tax |= !o.ignore_taxid && f.h.has_taxid_info();
    const bool mix = o.mix_taxid && any_tax;
    ukm_ctx* ctx = open_ctx(o);
    const bool tax = in.has_taxid;  // B-8: with --mix-taxid and no taxids in file 0 the output taxids are all 0
    if (tax) load_taxonomy(o, ctx);
    std::vector<ukm_span> sp = spans_of(in, tax);
    size_t total = 0;
    for (auto& s : sp) total += s.n;
    Result r(total, tax || mix);
    CHECK(ctx, ukm_common(ctx, sp.data(), (int)sp.size(), tax ? UKM_F_TAXID : 0, threshold, &r.span));
    if (mix && !tax) std::fill(r.taxids.begin(), r.taxids.end(), 0u);
    if (r.n() == 0) logi(o, "no shared k-mers found");
    write_result(o, in.k, base_mode(in, true, tax || mix), r.n(), o.max_taxid, r.codes.data(), (tax || mix) ? r.taxids.data() : nullptr,
                 r.n());  // common.go:313-337
    ukm_destroy(ctx);
    return 0;
}

int cmd_view(Options o) {  // view.go: text rows, `-t` adds the taxid, `-N` prints the code instead of the k-mer
    for (auto& path : o.files) {
        unik::File f = unik::read_file(path, false);
        const bool tx = o.show_taxid && f.taxids.size() == f.codes.size();
        std::string kmer((size_t)f.h.k, 'A');
        for (size_t i = 0; i < f.codes.size(); ++i) {
            if (f.h.is(unik::Hashed) || o.show_code) {
                printf("%llu", (unsigned long long)f.codes[i]);
            } else {
                uint64_t c = f.codes[i];
                for (int j = f.h.k - 1; j >= 0; --j) { kmer[j] = "ACGT"[c & 3]; c >>= 2; }
                fputs(kmer.c_str(), stdout);
            }
            if (tx) printf("\t%u", f.taxids[i]);
            putchar('\n');
        }
    }
    return 0;
}

int cmd_info(Options o) {  // info.go: header fields
    printf("file\tk\tcanonical\thashed\tscaled\tinclude-taxid\tglobal-taxid\tsorted\tcompact\tnumber\tcounted\n");
    for (auto& path : o.files) {
        unik::File f = unik::read_file(path, false);
        const unik::Header& h = f.h;
        printf("%s\t%d\t%d\t%d\t%d\t%d\t%u\t%d\t%d\t%llu\t%zu\n", path.c_str(), h.k, h.is(unik::Canonical), h.is(unik::Hashed),
               h.is(unik::Scaled), h.is(unik::IncludeTaxID), h.global_taxid, h.is(unik::Sorted), h.is(unik::Compact),
               (unsigned long long)h.number, f.codes.size());
    }
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 2 || !strcmp(argv[1], "-h") || !strcmp(argv[1], "--help")) {
        fprintf(stderr,
                "unikmer-b200: k-mer set operations on B200 (libukm)\n\n"
                "usage: unikmer-b200 <count|sort|union|inter|diff|common|split|merge|view|info> [flags] [files]\n"
                "flags follow unikmer: -o, -C, -c, -i, -I, --max-taxid, --data-dir, --verbose, --device;\n"
                "  count: -k -K -H -s --circular -D -W -t   sort: -u -d   union: -s   inter: -m\n"
                "  diff: -s -t   common: -n -p -m   split: -O -m -u -d   merge: -D -u -d   view: -t -N\n\n"
                "NOTE: the .unik v5 byte layout used here is restated from memory of shenwei356/unik v5.0.1 (its source is not\n"
                "part of the reference tree): files are self-consistent, interchange with files written by the real unikmer\n"
                "is UNVERIFIED (description length field, reserved header block, taxid-length byte; see host/unik.hpp).\n");
        return argc < 2 ? 1 : 0;
    }
    const std::string cmd = argv[1];
    try {
        Options o = parse(argc, argv, 2);
        if (cmd == "count") return cmd_count(o);
        if (cmd == "sort") return cmd_sort(o);
        if (cmd == "union") return cmd_union(o);
        if (cmd == "inter") return cmd_inter(o);
        if (cmd == "diff") return cmd_diff(o);
        if (cmd == "common") return cmd_common(o);
        if (cmd == "split") return cmd_split(o);
        if (cmd == "merge") return cmd_merge(o);
        if (cmd == "view") return cmd_view(o);
        if (cmd == "info") return cmd_info(o);
        die("unknown command: %s", cmd.c_str());
    } catch (const std::exception& e) {
        die("%s", e.what());
    }
}
