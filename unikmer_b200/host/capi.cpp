// capi.cpp -- C entry points of the host-side .unik codec (libukm_host.so) for the Python layer and tests.
#include <stdlib.h>
#include <string.h>

#include "unik.hpp"

static thread_local std::string g_err;

extern "C" {

struct ukmh_header {
    int k;
    uint32_t flag;
    uint64_t number;
    uint32_t global_taxid;
    uint32_t taxid_bytes;
    uint32_t scale;
    uint64_t max_hash;
    char description[1032];
};

const char* ukmh_last_error(void) { return g_err.c_str(); }
void ukmh_free(void* p) { free(p); }

static unik::Header to_header(const ukmh_header* h) {
    unik::Header u;
    u.k = h->k;
    u.flag = h->flag;
    u.number = h->number;
    u.global_taxid = h->global_taxid;
    u.taxid_bytes = (uint8_t)h->taxid_bytes;
    u.scale = h->scale;
    u.max_hash = h->max_hash;
    u.description = h->description;
    return u;
}
static void from_header(const unik::Header& u, ukmh_header* h) {
    memset(h, 0, sizeof *h);
    h->k = u.k;
    h->flag = u.flag;
    h->number = u.number;
    h->global_taxid = u.global_taxid;
    h->taxid_bytes = u.taxid_bytes;
    h->scale = u.scale;
    h->max_hash = u.max_hash;
    strncpy(h->description, u.description.c_str(), sizeof(h->description) - 1);
}

// encode to a malloc'd buffer (uncompressed stream)
int ukmh_unik_encode(const ukmh_header* h, const uint64_t* codes, const uint32_t* taxids, size_t n, uint8_t** out, size_t* len) {
    try {
        std::vector<uint8_t> b = unik::encode(to_header(h), codes, taxids, n);
        *out = (uint8_t*)malloc(b.size() ? b.size() : 1);
        memcpy(*out, b.data(), b.size());
        *len = b.size();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

static int give(const unik::File& f, ukmh_header* h, uint64_t** codes, uint32_t** taxids, size_t* n, size_t* n_tax) {
    from_header(f.h, h);
    *n = f.codes.size();
    *n_tax = f.taxids.size();
    *codes = (uint64_t*)malloc((*n ? *n : 1) * sizeof(uint64_t));
    *taxids = (uint32_t*)malloc((*n_tax ? *n_tax : 1) * sizeof(uint32_t));
    memcpy(*codes, f.codes.data(), *n * sizeof(uint64_t));
    memcpy(*taxids, f.taxids.data(), *n_tax * sizeof(uint32_t));
    return 0;
}

int ukmh_unik_decode(const uint8_t* buf, size_t len, int ignore_taxid, ukmh_header* h, uint64_t** codes, uint32_t** taxids, size_t* n,
                     size_t* n_tax) {
    try {
        return give(unik::decode(buf, len, ignore_taxid != 0), h, codes, taxids, n, n_tax);
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

int ukmh_unik_read_file(const char* path, int ignore_taxid, ukmh_header* h, uint64_t** codes, uint32_t** taxids, size_t* n, size_t* n_tax) {
    try {
        return give(unik::read_file(path, ignore_taxid != 0), h, codes, taxids, n, n_tax);
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

int ukmh_unik_write_file(const char* path, const ukmh_header* h, const uint64_t* codes, const uint32_t* taxids, size_t n, int compress,
                         int level) {
    try {
        unik::write_file(path, to_header(h), codes, taxids, n, compress != 0, level);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
}
