/*
 * oracle.c -- CPU restatement of the unikmer hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing in the product (unikmer_b200/, libukm.so) may include, link or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, as the checker / the CPU baseline.
 *
 * The reference (shenwei356/unikmer @ d608829, pure Go) cannot be built here (no Go
 * toolchain, un-vendored modules), so this file restates the
 * reference's algorithms in plain C, function by function, citing the Go lines it
 * follows (paths relative to /root/reference/unikmer/cmd/).  Third-party arithmetic
 * (shenwei356/kmers v0.1.0, will-rowe/nthash v0.4.0, shenwei356/bio v0.13.3
 * sketches/taxdump, twotwotwo/sorts) is restated from the published algorithms
 * (SURVEY.md Appendix A).
 *
 * Pinning: tests/test_oracle_kat.py checks this file against every known answer the
 * reference holds for the path (README.md / analysis/distance/README.md: K1..K9 of
 * SURVEY.md section 4).  Pieces with no vector anywhere in the reference
 * (taxdump.LCA edge cases, IUPAC handling inside ntHash) are "parity unpinned"
 * and say so at their definition.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_OK 0
#define ORC_E_ILLEGAL_BASE (-1)
#define ORC_E_ARG (-2)
#define ORC_E_NOMEM (-3)
#define ORC_E_PANIC (-4) /* the reference would panic / index out of range here */

/* ------------------------------------------------------------------------- */
/* A.1  shenwei356/kmers v0.1.0: 2-bit code, revcomp, canonical  [pinned K1-K7] */
/* ------------------------------------------------------------------------- */

/* base -> 2 bits.  A0 C1 G2 T3 (U=T), degenerate IUPAC -> alphabetically first
 * base (M,V,H,R,D,W,N->A; S,B,Y->C; K->G) [RECALL, unpinned for non-ACGT], any
 * other byte is kmers.ErrIllegalBase (surfaced at count.go:363-366). */
static int8_t g_base2bit[256];
static int g_tables_ready = 0;

static void init_tables(void) {
    if (g_tables_ready) return;
    memset(g_base2bit, -1, sizeof g_base2bit);
    const char* a0 = "AaMmVvHhRrDdWwNn";
    const char* c1 = "CcSsBbYy";
    const char* g2 = "GgKk";
    const char* t3 = "TtUu";
    for (const char* p = a0; *p; ++p) g_base2bit[(uint8_t)*p] = 0;
    for (const char* p = c1; *p; ++p) g_base2bit[(uint8_t)*p] = 1;
    for (const char* p = g2; *p; ++p) g_base2bit[(uint8_t)*p] = 2;
    for (const char* p = t3; *p; ++p) g_base2bit[(uint8_t)*p] = 3;
    g_tables_ready = 1;
}

/* kmers.Encode: first base in the most significant used bit pair. */
int orc_encode(const uint8_t* s, int k, uint64_t* code) {
    init_tables();
    if (k < 1 || k > 32) return ORC_E_ARG;
    uint64_t c = 0;
    for (int i = 0; i < k; ++i) {
        int8_t b = g_base2bit[s[i]];
        if (b < 0) return ORC_E_ILLEGAL_BASE;
        c = (c << 2) | (uint64_t)b;
    }
    *code = c;
    return ORC_OK;
}

/* kmers.RevComp: complement every base (XOR 3), reverse the k 2-bit groups. */
uint64_t orc_revcomp(uint64_t code, int k) {
    uint64_t r = 0;
    for (int i = 0; i < k; ++i) {
        r = (r << 2) | ((code & 3u) ^ 3u);
        code >>= 2;
    }
    return r;
}

/* kmers.Canonical = min(code, revcomp(code)). */
uint64_t orc_canonical(uint64_t code, int k) {
    uint64_t r = orc_revcomp(code, k);
    return r < code ? r : code;
}

/* kmers.Decode (view.go:173): inverse with alphabet ACGT. */
void orc_decode(uint64_t code, int k, char* out) {
    for (int i = k - 1; i >= 0; --i) {
        out[i] = "ACGT"[code & 3u];
        code >>= 2;
    }
}

/* A.3  bio/sketches NewKmerIterator/NextKmer (call sites count.go:321,363).
 * One iterator per record; emits len-k+1 codes (len codes when circular: the
 * iterator runs over seq + seq[0:k-1]).  Rolling update:
 *   code = ((prev & mask) << 2) + b ;  rc = ((b^3) << 2(k-1)) + (prevRC >> 2)
 * Returns the number of codes written, or a negative error.  len < k is
 * sketches.ErrShortSeq: the record is skipped by the caller (count.go:324-328);
 * here it yields 0 codes. */
int64_t orc_kmer_iter(const uint8_t* seq, int64_t len, int k, int canonical, int circular,
                      uint64_t* out) {
    init_tables();
    if (k < 1 || k > 32) return ORC_E_ARG;
    if (len < k) return 0;
    int64_t total = circular ? len + k - 1 : len;
    uint64_t mask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
    uint64_t fw = 0, rc = 0;
    int64_t n = 0;
    for (int64_t i = 0; i < total; ++i) {
        uint8_t ch = seq[i < len ? i : i - len];
        int8_t b = g_base2bit[ch];
        if (b < 0) return ORC_E_ILLEGAL_BASE;
        fw = ((fw << 2) | (uint64_t)b) & mask;
        rc = (rc >> 2) | ((uint64_t)(b ^ 3) << (2 * (k - 1)));
        if (i >= k - 1) out[n++] = (canonical && rc < fw) ? rc : fw;
    }
    return n;
}

/* ------------------------------------------------------------------------- */
/* A.2  will-rowe/nthash v0.4.0 (ntHash v1)                       [pinned K8, K9] */
/* ------------------------------------------------------------------------- */
static const uint64_t NT_A = 0x3c8bfbb395c60474ull, NT_C = 0x3193c18562a02b4cull,
                      NT_G = 0x20323ed082572324ull, NT_T = 0x295549f54be24456ull;

static inline uint64_t rol64(uint64_t v, unsigned s) { s &= 63; return s ? (v << s) | (v >> (64 - s)) : v; }
static inline uint64_t ror64(uint64_t v, unsigned s) { s &= 63; return s ? (v >> s) | (v << (64 - s)) : v; }

/* seed of a base / of its complement.  acgt(u) like ACGT(U); everything else
 * (N, IUPAC) contributes 0 and the k-mer is NOT skipped [RECALL, unpinned]. */
static inline uint64_t nt_seed(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return NT_A;
        case 'C': case 'c': return NT_C;
        case 'G': case 'g': return NT_G;
        case 'T': case 't': case 'U': case 'u': return NT_T;
        default: return 0;
    }
}
static inline uint64_t nt_seed_comp(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return NT_T;
        case 'C': case 'c': return NT_G;
        case 'G': case 'g': return NT_C;
        case 'T': case 't': case 'U': case 'u': return NT_A;
        default: return 0;
    }
}

/* sketches.NewHashIterator/NextHash (count.go:319,361): canonical = min(fwd, rev).
 *   fwd = XOR_i rol(seed[s_i], k-1-i) ;  rev = XOR_i rol(seed[comp(s_i)], i)
 *   fwd' = rol(fwd,1) ^ rol(seed[out],k) ^ seed[in]
 *   rev' = ror(rev,1) ^ ror(seed[comp(out)],1) ^ rol(seed[comp(in)],k-1)
 * k <= 64 in hashed mode (count.go:85-87). */
int64_t orc_nthash_iter(const uint8_t* seq, int64_t len, int k, int canonical, int circular,
                        uint64_t* out) {
    if (k < 1 || k > 64) return ORC_E_ARG;
    if (len < k) return 0;
    int64_t total = circular ? len + k - 1 : len;
#define SEQ(i) (seq[(i) < len ? (i) : (i) - len])
    uint64_t fh = 0, rh = 0;
    for (int i = 0; i < k; ++i) {
        fh ^= rol64(nt_seed(SEQ(i)), (unsigned)(k - 1 - i));
        rh ^= rol64(nt_seed_comp(SEQ(i)), (unsigned)i);
    }
    int64_t n = 0;
    out[n++] = (canonical && rh < fh) ? rh : fh;
    for (int64_t i = k; i < total; ++i) {
        uint8_t cin = SEQ(i), cout = SEQ(i - k);
        fh = rol64(fh, 1) ^ rol64(nt_seed(cout), (unsigned)k) ^ nt_seed(cin);
        rh = ror64(rh, 1) ^ ror64(nt_seed_comp(cout), 1) ^ rol64(nt_seed_comp(cin), (unsigned)(k - 1));
        out[n++] = (canonical && rh < fh) ? rh : fh;
    }
#undef SEQ
    return n;
}

/* ------------------------------------------------------------------------- */
/* A.5  bio/taxdump Taxonomy.LCA                          [RECALL, parity unpinned] */
/* ------------------------------------------------------------------------- */
typedef struct {
    uint32_t* parent; /* parent[t]; 0 = unknown taxid; root has parent[t]==t */
    uint32_t* merged; /* merged[t] = new id for a merged (old) id, else 0 */
    size_t n;         /* table length = max id + 1 */
} orc_tax;

orc_tax* orc_tax_new(const uint32_t* parent, size_t n, const uint32_t* merged_from,
                     const uint32_t* merged_to, size_t n_merged) {
    size_t nn = n;
    for (size_t i = 0; i < n_merged; ++i)
        if ((size_t)merged_from[i] + 1 > nn) nn = (size_t)merged_from[i] + 1;
    orc_tax* t = (orc_tax*)calloc(1, sizeof *t);
    t->parent = (uint32_t*)calloc(nn ? nn : 1, sizeof(uint32_t));
    t->merged = (uint32_t*)calloc(nn ? nn : 1, sizeof(uint32_t));
    t->n = nn;
    memcpy(t->parent, parent, n * sizeof(uint32_t));
    for (size_t i = 0; i < n_merged; ++i) t->merged[merged_from[i]] = merged_to[i];
    return t;
}
void orc_tax_free(orc_tax* t) {
    if (!t) return;
    free(t->parent); free(t->merged); free(t);
}

/* node lookup as taxdump does while walking: a node missing from Nodes is
 * remapped through the merged table; unknown and unmerged => 0. */
static inline uint32_t tax_resolve(const orc_tax* t, uint32_t x) {
    if (x < t->n && t->parent[x]) return x;
    if (x < t->n && t->merged[x]) {
        uint32_t y = t->merged[x];
        if (y < t->n && t->parent[y]) return y;
    }
    return 0;
}

/* LCA(a,b): 0 if either is 0; a if a==b (no validity check); unknown id => 0;
 * else the lowest common ancestor in the nodes.dmp tree.  Commutative and
 * associative on valid ids, 0 absorbing => folds are order-free. */
uint32_t orc_lca(const orc_tax* t, uint32_t a, uint32_t b) {
    if (a == 0 || b == 0) return 0;
    if (a == b) return a;
    if (!t) return 0;
    uint32_t ra = tax_resolve(t, a), rb = tax_resolve(t, b);
    if (!ra || !rb) return 0;
    if (ra == rb) return ra;
    /* depth of each (walk to root), then climb in lock-step */
    size_t da = 0, db = 0;
    for (uint32_t x = ra; t->parent[x] != x; x = t->parent[x]) { if (++da > t->n) return 0; }
    for (uint32_t x = rb; t->parent[x] != x; x = t->parent[x]) { if (++db > t->n) return 0; }
    while (da > db) { ra = t->parent[ra]; --da; }
    while (db > da) { rb = t->parent[rb]; --db; }
    while (ra != rb) { ra = t->parent[ra]; rb = t->parent[rb]; }
    return ra;
}

/* ------------------------------------------------------------------------- */
/* Go-map stand-in: open-addressing hash map uint64 -> {uint32 val, uint16 cnt}  */
/* ------------------------------------------------------------------------- */
typedef struct {
    uint64_t* keys;
    uint32_t* vals;
    uint16_t* cnts;
    uint8_t* used;
    size_t cap, n;
} hmap;

static inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static int hmap_init(hmap* h, size_t expect) {
    size_t cap = 16;
    while (cap < expect * 2 + 2) cap <<= 1;
    h->keys = (uint64_t*)malloc(cap * sizeof(uint64_t));
    h->vals = (uint32_t*)malloc(cap * sizeof(uint32_t));
    h->cnts = (uint16_t*)malloc(cap * sizeof(uint16_t));
    h->used = (uint8_t*)calloc(cap, 1);
    h->cap = cap; h->n = 0;
    return (h->keys && h->vals && h->cnts && h->used) ? 0 : ORC_E_NOMEM;
}
static void hmap_free(hmap* h) { free(h->keys); free(h->vals); free(h->cnts); free(h->used); }
static int hmap_grow(hmap* h);
/* returns slot; *found says whether key was present (inserted if not and insert!=0) */
static inline size_t hmap_slot(hmap* h, uint64_t key, int insert, int* found) {
    size_t m = h->cap - 1, i = (size_t)mix64(key) & m;
    while (h->used[i]) {
        if (h->keys[i] == key) { *found = 1; return i; }
        i = (i + 1) & m;
    }
    *found = 0;
    if (insert) {
        if ((h->n + 1) * 2 > h->cap) { hmap_grow(h); return hmap_slot(h, key, insert, found); }
        h->used[i] = 1; h->keys[i] = key; h->vals[i] = 0; h->cnts[i] = 0; h->n++;
    }
    return i;
}
static int hmap_grow(hmap* h) {
    hmap o = *h;
    if (hmap_init(h, o.cap)) return ORC_E_NOMEM;
    for (size_t i = 0; i < o.cap; ++i)
        if (o.used[i]) {
            int f; size_t s = hmap_slot(h, o.keys[i], 1, &f);
            h->vals[s] = o.vals[i]; h->cnts[s] = o.cnts[i];
        }
    hmap_free(&o);
    return 0;
}
/* delete(m, key) with backward-shift so probing stays valid */
static void hmap_del(hmap* h, size_t i) {
    size_t m = h->cap - 1, j = i;
    for (;;) {
        j = (j + 1) & m;
        if (!h->used[j]) break;
        size_t home = (size_t)mix64(h->keys[j]) & m;
        if ((i <= j) ? (home <= i || home > j) : (home <= i && home > j)) {
            h->keys[i] = h->keys[j]; h->vals[i] = h->vals[j]; h->cnts[i] = h->cnts[j];
            i = j;
        }
    }
    h->used[i] = 0; h->n--;
}

/* ------------------------------------------------------------------------- */
/* A.6  sortutil.Uint64s / sorts.Quicksort(CodeTaxidSlice)                      */
/* ------------------------------------------------------------------------- */
static void lsd_radix_u64(uint64_t* a, uint64_t* tmp, size_t n, uint32_t* v, uint32_t* vtmp) {
    if (n < 2) return;
    uint64_t ored = 0, anded = ~0ull;
    for (size_t i = 0; i < n; ++i) { ored |= a[i]; anded &= a[i]; }
    uint64_t diff = ored ^ anded;
    for (int pass = 0; pass < 8; ++pass) {
        int sh = pass * 8;
        if (((diff >> sh) & 0xff) == 0) continue; /* digit constant: skip */
        size_t cnt[257] = {0};
        for (size_t i = 0; i < n; ++i) cnt[((a[i] >> sh) & 0xff) + 1]++;
        for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
        if (v) {
            for (size_t i = 0; i < n; ++i) { size_t p = cnt[(a[i] >> sh) & 0xff]++; tmp[p] = a[i]; vtmp[p] = v[i]; }
            memcpy(v, vtmp, n * sizeof(uint32_t));
        } else {
            for (size_t i = 0; i < n; ++i) tmp[cnt[(a[i] >> sh) & 0xff]++] = a[i];
        }
        memcpy(a, tmp, n * sizeof(uint64_t));
    }
}

/* sortutil.Uint64s (sort.go:463, union.go:295, diff.go:587, common.go:344,
 * count.go:581): in-place ascending sort, parallel MSD radix over up to
 * sorts.MaxProcs goroutines (util.go:91).  Restated as: one parallel MSD split on
 * the top differing byte, then an LSD radix per bucket, buckets spread over
 * `threads` OpenMP threads.  The result is the unique ascending permutation. */
int orc_sort_u64(uint64_t* a, size_t n, int threads) {
    if (n < 2) return ORC_OK;
    uint64_t* tmp = (uint64_t*)malloc(n * sizeof(uint64_t));
    if (!tmp) return ORC_E_NOMEM;
    if (threads <= 1 || n < (1u << 16)) {
        lsd_radix_u64(a, tmp, n, NULL, NULL);
        free(tmp);
        return ORC_OK;
    }
    uint64_t ored = 0, anded = ~0ull;
#pragma omp parallel for num_threads(threads) reduction(| : ored) reduction(& : anded)
    for (size_t i = 0; i < n; ++i) { ored |= a[i]; anded &= a[i]; }
    uint64_t diff = ored ^ anded;
    if (!diff) { free(tmp); return ORC_OK; }
    int top = 63 - __builtin_clzll(diff);
    int sh = top >= 7 ? top - 7 : 0;
    int T = threads;
    size_t* hist = (size_t*)calloc((size_t)T * 256, sizeof(size_t));
    size_t chunk = (n + T - 1) / T;
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        size_t lo = (size_t)t * chunk, hi = lo + chunk < n ? lo + chunk : n;
        size_t* h = hist + (size_t)t * 256;
        for (size_t i = lo; i < hi; ++i) h[(a[i] >> sh) & 0xff]++;
    }
    size_t start[257]; size_t acc = 0;
    for (int d = 0; d < 256; ++d) {
        start[d] = acc;
        for (int t = 0; t < T; ++t) { size_t c = hist[(size_t)t * 256 + d]; hist[(size_t)t * 256 + d] = acc; acc += c; }
    }
    start[256] = n;
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        size_t lo = (size_t)t * chunk, hi = lo + chunk < n ? lo + chunk : n;
        size_t* h = hist + (size_t)t * 256;
        for (size_t i = lo; i < hi; ++i) tmp[h[(a[i] >> sh) & 0xff]++] = a[i];
    }
#pragma omp parallel for num_threads(T) schedule(dynamic, 1)
    for (int d = 0; d < 256; ++d) {
        size_t m = start[d + 1] - start[d];
        if (m) lsd_radix_u64(tmp + start[d], a + start[d], m, NULL, NULL);
    }
    memcpy(a, tmp, n * sizeof(uint64_t));
    free(hist); free(tmp);
    return ORC_OK;
}

/* sorts.Quicksort(CodeTaxidSlice(mt)) (sort.go:457; kmers.go:24-46): sort records
 * by Code only; the reference is NOT stable (tie order undefined, quirk B-10).
 * Restated as a stable LSD radix so the oracle is deterministic; tests compare
 * equal-code groups as multisets. */
int orc_sort_pairs(uint64_t* keys, uint32_t* taxids, size_t n) {
    if (n < 2) return ORC_OK;
    uint64_t* tmp = (uint64_t*)malloc(n * sizeof(uint64_t));
    uint32_t* vtmp = (uint32_t*)malloc(n * sizeof(uint32_t));
    if (!tmp || !vtmp) { free(tmp); free(vtmp); return ORC_E_NOMEM; }
    lsd_radix_u64(keys, tmp, n, taxids, vtmp);
    free(tmp); free(vtmp);
    return ORC_OK;
}

/* ------------------------------------------------------------------------- */
/* Fold of a sorted slice: sort.go:482-573, util-sort.go:35-190                 */
/* ------------------------------------------------------------------------- */
enum { ORC_FOLD_PLAIN = 0, ORC_FOLD_UNIQUE = 1, ORC_FOLD_REPEATED_FINAL = 2, ORC_FOLD_REPEATED_CHUNK = 3 };

/* Emits into out_keys/out_taxids (capacity >= n+2) and returns the count.
 * taxids == NULL selects the []uint64 branches.  Quirks B-1/B-2 are kept:
 *  - with taxids, -u always writes the trailing (last,lca) even for empty input
 *    => one record (^uint64(0), 0) (sort.go:505-507);
 *  - without taxids, a first code equal to ^uint64(0) is dropped by -u
 *    (sort.go:542-549) because `last` starts at ^uint64(0). */
int64_t orc_fold(int mode, const uint64_t* keys, const uint32_t* taxids, size_t n,
                 const orc_tax* tax, uint64_t* out_keys, uint32_t* out_taxids) {
    int64_t m = 0;
    uint64_t last = ~0ull;
    if (taxids) {
        uint32_t lca = 0;
        if (mode == ORC_FOLD_UNIQUE) { /* sort.go:488-507, util-sort.go:122-144 */
            int first = 1;
            for (size_t i = 0; i < n; ++i) {
                if (keys[i] == last) { lca = orc_lca(tax, taxids[i], lca); continue; }
                if (first) first = 0;
                else { out_keys[m] = last; out_taxids[m] = lca; ++m; }
                last = keys[i]; lca = taxids[i];
            }
            out_keys[m] = last; out_taxids[m] = lca; ++m;
        } else if (mode == ORC_FOLD_REPEATED_FINAL) { /* sort.go:508-532 */
            int count = 1;
            for (size_t i = 0; i < n; ++i) {
                if (keys[i] == last) { lca = orc_lca(tax, taxids[i], lca); ++count; continue; }
                if (count > 1) { out_keys[m] = last; out_taxids[m] = lca; ++m; count = 1; }
                last = keys[i]; lca = taxids[i];
            }
            if (count > 1) { out_keys[m] = last; out_taxids[m] = lca; ++m; }
        } else if (mode == ORC_FOLD_REPEATED_CHUNK) { /* util-sort.go:145-178 */
            int count = 0;
            for (size_t i = 0; i < n; ++i) {
                if (keys[i] == last) { lca = orc_lca(tax, taxids[i], lca); ++count; continue; }
                if (count > 0) {
                    out_keys[m] = last; out_taxids[m] = lca; ++m;
                    if (count > 1) { out_keys[m] = last; out_taxids[m] = lca; ++m; }
                }
                count = 1; last = keys[i]; lca = taxids[i];
            }
            out_keys[m] = last; out_taxids[m] = lca; ++m;
            if (count > 1) { out_keys[m] = last; out_taxids[m] = lca; ++m; }
        } else { /* sort.go:533-538 */
            for (size_t i = 0; i < n; ++i) { out_keys[m] = keys[i]; out_taxids[m] = taxids[i]; ++m; }
        }
    } else {
        if (mode == ORC_FOLD_UNIQUE) { /* sort.go:541-549, util-sort.go:53-60 */
            for (size_t i = 0; i < n; ++i) {
                if (keys[i] == last) continue;
                last = keys[i]; out_keys[m++] = keys[i];
            }
        } else if (mode == ORC_FOLD_REPEATED_FINAL) { /* sort.go:550-565 */
            int count = 0;
            for (size_t i = 0; i < n; ++i) {
                if (keys[i] == last) { if (count == 1) out_keys[m++] = keys[i]; ++count; }
                else { last = keys[i]; count = 1; }
            }
        } else if (mode == ORC_FOLD_REPEATED_CHUNK) { /* util-sort.go:61-92 */
            int count = 0;
            for (size_t i = 0; i < n; ++i) {
                if (keys[i] == last) { ++count; continue; }
                if (count > 0) { out_keys[m++] = last; if (count > 1) out_keys[m++] = last; }
                count = 1; last = keys[i];
            }
            out_keys[m++] = last;
            if (count > 1) out_keys[m++] = last;
        } else { /* sort.go:566-571 */
            for (size_t i = 0; i < n; ++i) out_keys[m++] = keys[i];
        }
    }
    return m;
}

/* ------------------------------------------------------------------------- */
/* Set operations                                                              */
/* ------------------------------------------------------------------------- */
typedef struct {
    const uint64_t* keys;
    const uint32_t* taxids; /* per-k-mer taxids as ReadCodeWithTaxid returns them (global taxid
                               already broadcast by the caller); NULL => all 0 */
    size_t n;
    int sorted; /* header flag reader.IsSorted() */
} orc_file;

static inline uint32_t file_tax(const orc_file* f, size_t i) { return f->taxids ? f->taxids[i] : 0; }

/* mergeChunksFile (util-sort.go:227-606): heap k-way merge of sorted streams, then the
 * same plain/unique/repeated folds (finalRound switches REPEATED_CHUNK -> _FINAL
 * emission, util-sort.go:377-388).  Restated as merge-all + orc_fold: the heap only
 * defines the (ascending) visiting order; LCA folds are order-free.  The plain
 * variant with taxids keeps ties in heap-pop order in the reference (undefined);
 * here ties are in file order.  Returns the output count. */
int64_t orc_merge_chunks(const orc_file* files, int nfiles, int has_taxid, int mode,
                         const orc_tax* tax, uint64_t* out_keys, uint32_t* out_taxids) {
    size_t total = 0;
    for (int f = 0; f < nfiles; ++f) total += files[f].n;
    uint64_t* mk = (uint64_t*)malloc((total + 1) * sizeof(uint64_t));
    uint32_t* mv = has_taxid ? (uint32_t*)malloc((total + 1) * sizeof(uint32_t)) : NULL;
    size_t* cur = (size_t*)calloc((size_t)nfiles + 1, sizeof(size_t));
    if (!mk || (has_taxid && !mv) || !cur) { free(mk); free(mv); free(cur); return ORC_E_NOMEM; }
    /* simple tournament by linear scan (nfiles is small in tests); ties -> lowest file index */
    for (size_t o = 0; o < total; ++o) {
        int best = -1;
        for (int f = 0; f < nfiles; ++f) {
            if (cur[f] >= files[f].n) continue;
            if (best < 0 || files[f].keys[cur[f]] < files[best].keys[cur[best]]) best = f;
        }
        mk[o] = files[best].keys[cur[best]];
        if (has_taxid) mv[o] = file_tax(&files[best], cur[best]);
        cur[best]++;
    }
    int64_t m = orc_fold(mode, mk, has_taxid ? mv : NULL, total, tax, out_keys, out_taxids);
    free(mk); free(mv); free(cur);
    return m;
}

/* union (union.go:186-208 accumulate, 260-305 ordered emit = the `-s` contract).
 * Go map m / mt; with taxids mt[code] = LCA(mt[code], taxid) in file order.  Output:
 * keys ascending (sortutil.Uint64s, union.go:295), taxid = mt[code].  Returns n. */
int64_t orc_union(const orc_file* files, int nfiles, int has_taxid, const orc_tax* tax,
                  int threads, uint64_t* out_keys, uint32_t* out_taxids) {
    size_t total = 0;
    for (int f = 0; f < nfiles; ++f) total += files[f].n;
    hmap h;
    if (hmap_init(&h, files[0].n + 1024)) return ORC_E_NOMEM;
    for (int f = 0; f < nfiles; ++f) {
        const orc_file* F = &files[f];
        for (size_t i = 0; i < F->n; ++i) {
            int found;
            size_t s = hmap_slot(&h, F->keys[i], 1, &found);
            if (has_taxid) h.vals[s] = found ? orc_lca(tax, h.vals[s], file_tax(F, i)) : file_tax(F, i);
        }
    }
    int64_t m = 0;
    for (size_t i = 0; i < h.cap; ++i) if (h.used[i]) out_keys[m++] = h.keys[i];
    orc_sort_u64(out_keys, (size_t)m, threads);
    if (has_taxid)
        for (int64_t i = 0; i < m; ++i) { int f; out_taxids[i] = h.vals[hmap_slot(&h, out_keys[i], 0, &f)]; }
    hmap_free(&h);
    (void)total;
    return m;
}

/* inter (inter.go:188-203 load, 205-267 two-pointer, 269-286 compaction/early stop).
 * mc <- file 0; for every next file a three-way compare walk marks matches,
 * taxid <- LCA(q,t) (hasTaxid) or the mix rule (inter.go:229-239).  Quirk B-3: a
 * later EMPTY file returns flagBreak before filtering => output = current mc; an
 * empty FIRST file panics on mc[0] when a second file exists (=> ORC_E_PANIC).
 * Single input file = byte copy (inter.go:96-120) is the caller's business. */
int64_t orc_inter(const orc_file* files, int nfiles, int has_taxid, int mix_taxid,
                  const orc_tax* tax, uint64_t* out_keys, uint32_t* out_taxids) {
    size_t n = files[0].n;
    uint64_t* mc = (uint64_t*)malloc((n + 1) * sizeof(uint64_t));
    uint32_t* mt = (uint32_t*)malloc((n + 1) * sizeof(uint32_t));
    uint8_t* mk = (uint8_t*)malloc(n + 1);
    if (!mc || !mt || !mk) { free(mc); free(mt); free(mk); return ORC_E_NOMEM; }
    for (size_t i = 0; i < n; ++i) { mc[i] = files[0].keys[i]; mt[i] = file_tax(&files[0], i); }
    for (int f = 1; f < nfiles; ++f) {
        const orc_file* F = &files[f];
        if (n == 0) { free(mc); free(mt); free(mk); return ORC_E_PANIC; } /* mc[0] */
        if (F->n == 0) break;                                            /* flagBreak, B-3 */
        memset(mk, 0, n);
        size_t ii = 0, j = 0;
        for (;;) {
            uint64_t q = mc[ii], c = F->keys[j];
            if (q < c) { if (++ii >= n) break; }
            else if (q == c) {
                uint32_t qt = mt[ii], t = file_tax(F, j);
                if (mix_taxid) mt[ii] = qt == 0 ? t : (t == 0 ? qt : orc_lca(tax, qt, t));
                else if (has_taxid) mt[ii] = orc_lca(tax, qt, t);
                mk[ii] = 1;
                if (++ii >= n) break;
                if (++j >= F->n) break;
            } else { if (++j >= F->n) break; }
        }
        size_t m = 0;
        for (size_t i = 0; i < n; ++i) if (mk[i]) { mc[m] = mc[i]; mt[m] = mt[i]; ++m; }
        n = m;
        if (n == 0) break; /* hasInter = false */
    }
    for (size_t i = 0; i < n; ++i) { out_keys[i] = mc[i]; if (out_taxids) out_taxids[i] = mt[i]; }
    free(mc); free(mt); free(mk);
    return (int64_t)n;
}

/* diff (diff.go:136-146 load; sorted subject two-pointer 380-435; unsorted subject map
 * delete 341-367; per-file map rebuild 449-453; worker-map intersection 496-515; `-s`
 * emit 566-594).  Restated for one worker (the -j split only partitions the subject
 * files; the final result is the intersection of the worker maps = file0 minus every
 * subject).  compare_taxid (-t): a shared k-mer STAYS if qtaxid==taxid or
 * LCA(taxid,qtaxid)==qtaxid (361-364, 406-409).  The result is keyed by code (map),
 * so duplicate codes of file 0 collapse (last taxid wins, diff.go:450-452); output
 * ascending with file 0's taxid.  Quirk B-5 (a sorted EMPTY subject ends the worker
 * and, if no map was ever stored, yields an empty result) is kept for nfiles>=2 with
 * one worker: the worker breaks out of its file loop; maps[i] is whatever was stored
 * before. */
int64_t orc_diff(const orc_file* files, int nfiles, int has_taxid, int compare_taxid,
                 const orc_tax* tax, int threads, uint64_t* out_keys, uint32_t* out_taxids) {
    size_t n = files[0].n;
    if (n == 0) return 0; /* diff.go:155-201: header-only output */
    uint64_t* mc = (uint64_t*)malloc((n + 1) * sizeof(uint64_t));
    uint32_t* mt = (uint32_t*)malloc((n + 1) * sizeof(uint32_t));
    uint64_t* mc2 = (uint64_t*)malloc((n + 1) * sizeof(uint64_t));
    uint32_t* mt2 = (uint32_t*)malloc((n + 1) * sizeof(uint32_t));
    if (!mc || !mt || !mc2 || !mt2) { free(mc); free(mt); free(mc2); free(mt2); return ORC_E_NOMEM; }
    for (size_t i = 0; i < n; ++i) { mc[i] = files[0].keys[i]; mt[i] = file_tax(&files[0], i); }
    hmap h; int have_map = 0;
    int64_t result = 0;
    for (int f = 1; f < nfiles; ++f) {
        const orc_file* F = &files[f];
        if (!F->sorted) {
            if (!have_map) { /* diff.go:342-348 */
                if (hmap_init(&h, n + 16)) { result = ORC_E_NOMEM; goto done; }
                for (size_t i = 0; i < n; ++i) { int fo; size_t s = hmap_slot(&h, mc[i], 1, &fo); h.vals[s] = mt[i]; }
                have_map = 1;
            }
            for (size_t j = 0; j < F->n; ++j) { /* diff.go:350-367 */
                int fo; size_t s = hmap_slot(&h, F->keys[j], 0, &fo);
                if (!fo) continue;
                uint32_t qt = h.vals[s], t = file_tax(F, j);
                if (compare_taxid && (qt == t || orc_lca(tax, t, qt) == qt)) continue;
                hmap_del(&h, s);
            }
            if (h.n == 0) { result = 0; goto done; } /* hasDiff=false */
            /* NOTE: the reference does not rebuild mc1 from m1 here; a later SORTED subject
             * walks the stale mc1 (diff.go:380-435) and then overwrites the map (449-453),
             * resurrecting k-mers the unsorted subject removed.  Kept as is. */
        } else {
            if (F->n == 0) break; /* diff.go:387-392: `break` leaves the worker loop (B-5) */
            size_t ii = 0, j = 0, m = 0;
            for (;;) { /* diff.go:395-431 */
                uint64_t q = mc[ii], c = F->keys[j];
                if (q < c) { mc2[m] = mc[ii]; mt2[m] = mt[ii]; ++m; if (++ii >= n) break; }
                else if (q == c) {
                    uint32_t qt = mt[ii], t = file_tax(F, j);
                    if (compare_taxid && (qt == t || orc_lca(tax, t, qt) == qt)) { mc2[m] = mc[ii]; mt2[m] = mt[ii]; ++m; }
                    if (++ii >= n) break;
                    if (++j >= F->n) break;
                } else { if (++j >= F->n) break; }
            }
            for (; ii < n; ++ii) { mc2[m] = mc[ii]; mt2[m] = mt[ii]; ++m; } /* diff.go:432 */
            { uint64_t* tk = mc; mc = mc2; mc2 = tk; uint32_t* tv = mt; mt = mt2; mt2 = tv; }
            n = m;
            if (n == 0) { if (have_map) hmap_free(&h); have_map = 0; result = 0; goto done0; }
            if (have_map) hmap_free(&h); /* diff.go:449-453 rebuild */
            if (hmap_init(&h, n + 16)) { result = ORC_E_NOMEM; goto done0; }
            for (size_t i = 0; i < n; ++i) { int fo; size_t s = hmap_slot(&h, mc[i], 1, &fo); h.vals[s] = mt[i]; }
            have_map = 1;
        }
    }
    if (!have_map) { result = 0; goto done0; } /* m0 == nil: empty output (B-5; also nfiles==1) */
    {
        int64_t m = 0;
        for (size_t i = 0; i < h.cap; ++i) if (h.used[i]) out_keys[m++] = h.keys[i];
        orc_sort_u64(out_keys, (size_t)m, threads); /* diff.go:587 */
        if (out_taxids)
            for (int64_t i = 0; i < m; ++i) { int fo; out_taxids[i] = h.vals[hmap_slot(&h, out_keys[i], 0, &fo)]; }
        result = m;
    }
done:
    if (have_map) hmap_free(&h);
done0:
    free(mc); free(mt); free(mc2); free(mt2);
    (void)has_taxid;
    return result;
}

/* common (common.go:220-250 first file, 252-283 other files, 329-354 select+sort+emit).
 * counts map[uint64]uint16: file 0 sets 1 (dedup), later files ++ per occurrence,
 * wrapping at 65536 (B-7); mt[code]: file 0 overwrites (last taxid wins), later
 * files LCA-fold.  Keep count >= threshold; ascending. */
int64_t orc_common(const orc_file* files, int nfiles, int has_taxid, uint16_t threshold,
                   const orc_tax* tax, int threads, uint64_t* out_keys, uint32_t* out_taxids) {
    hmap h;
    if (hmap_init(&h, files[0].n + 1024)) return ORC_E_NOMEM;
    for (size_t i = 0; i < files[0].n; ++i) {
        int fo; size_t s = hmap_slot(&h, files[0].keys[i], 1, &fo);
        if (has_taxid) h.vals[s] = file_tax(&files[0], i);
        h.cnts[s] = 1;
    }
    for (int f = 1; f < nfiles; ++f) {
        const orc_file* F = &files[f];
        for (size_t i = 0; i < F->n; ++i) {
            int fo; size_t s = hmap_slot(&h, F->keys[i], 1, &fo);
            if (has_taxid) h.vals[s] = fo ? orc_lca(tax, h.vals[s], file_tax(F, i)) : file_tax(F, i);
            h.cnts[s] = (uint16_t)(h.cnts[s] + 1);
        }
    }
    int64_t m = 0;
    for (size_t i = 0; i < h.cap; ++i) if (h.used[i] && h.cnts[i] >= threshold) out_keys[m++] = h.keys[i];
    orc_sort_u64(out_keys, (size_t)m, threads);
    if (out_taxids)
        for (int64_t i = 0; i < m; ++i) { int fo; size_t s = hmap_slot(&h, out_keys[i], 0, &fo); out_taxids[i] = has_taxid ? h.vals[s] : 0; }
    hmap_free(&h);
    return m;
}

/* count (count.go:314-322 iterator choice, 355-437 inner loop, 373 scaled filter,
 * 434-436 dedup map, 531-595 sort+emit).  `bases` is the concatenation of all
 * records (newlines already stripped, as bio/seqio/fastx hands them over), record r
 * = bases[rec_off[r] : rec_off[r+1]].  Records shorter than k are skipped
 * (count.go:324-328).  Returns the number of distinct codes, ascending in out_keys
 * (capacity >= total k-mers), or a negative error (illegal base, count.go:363-366). */
int64_t orc_count(const uint8_t* bases, const uint64_t* rec_off, size_t n_rec, int k,
                  int canonical, int hashed, int circular, int scaled, uint64_t max_hash,
                  int threads, uint64_t* out_keys, size_t out_cap) {
    size_t maxlen = 0, total = 0;
    for (size_t r = 0; r < n_rec; ++r) {
        size_t L = (size_t)(rec_off[r + 1] - rec_off[r]);
        if (L > maxlen) maxlen = L;
        if (L >= (size_t)k) total += circular ? L : L - (size_t)k + 1;
    }
    uint64_t* buf = (uint64_t*)malloc((maxlen + 64) * sizeof(uint64_t));
    if (!buf) return ORC_E_NOMEM;
    hmap h;
    if (hmap_init(&h, total / 2 + 1024)) { free(buf); return ORC_E_NOMEM; }
    for (size_t r = 0; r < n_rec; ++r) {
        const uint8_t* s = bases + rec_off[r];
        int64_t L = (int64_t)(rec_off[r + 1] - rec_off[r]);
        int64_t m = hashed ? orc_nthash_iter(s, L, k, canonical, circular, buf)
                           : orc_kmer_iter(s, L, k, canonical, circular, buf);
        if (m < 0) { hmap_free(&h); free(buf); return m; }
        for (int64_t i = 0; i < m; ++i) {
            if (scaled && buf[i] > max_hash) continue; /* count.go:373 */
            int fo; hmap_slot(&h, buf[i], 1, &fo);      /* count.go:434-436 */
        }
    }
    int64_t n = 0;
    for (size_t i = 0; i < h.cap; ++i) if (h.used[i]) { if ((size_t)n >= out_cap) { n = ORC_E_ARG; break; } out_keys[n++] = h.keys[i]; }
    if (n > 0) orc_sort_u64(out_keys, (size_t)n, threads); /* count.go:581 */
    hmap_free(&h); free(buf);
    return n;
}

/* count -W w (count.go:100-114 flag handling, 316-317 sketches.NewMinimizerSketch, 358-359
 * NextMinimizer, then the same scaled filter / dedup map / sort as orc_count).  bio/sketches is
 * not in the reference tree; what the reference pins is the answer: on E. coli MG1655 with
 * -k 31 -K -H -W 15 the file holds 549 963 k-mers (analysis/distance/README.md:8,15,39), which is
 * exactly the number of distinct values of  min(h[i .. i+w-1])  over every FULL window of w
 * consecutive canonical ntHash values of a record (reproduced; partial windows at the record
 * ends give 549 965 / 549 967).  So: per record, the sliding-window minimum of the hash stream,
 * windows never span records, records with fewer than w k-mers contribute nothing.  The code
 * emitted is the hash itself (-W switches -H on, count.go:105-109). */
int64_t orc_count_minimizer(const uint8_t* bases, const uint64_t* rec_off, size_t n_rec, int k, int w,
                            int canonical, int circular, int scaled, uint64_t max_hash,
                            int threads, uint64_t* out_keys, size_t out_cap) {
    if (w < 1) return ORC_E_ARG;
    size_t maxlen = 0, total = 0;
    for (size_t r = 0; r < n_rec; ++r) {
        size_t L = (size_t)(rec_off[r + 1] - rec_off[r]);
        if (L > maxlen) maxlen = L;
        if (L >= (size_t)k) total += circular ? L : L - (size_t)k + 1;
    }
    uint64_t* buf = (uint64_t*)malloc((maxlen + 64) * sizeof(uint64_t));
    int64_t* dq = (int64_t*)malloc((maxlen + 64) * sizeof(int64_t)); /* monotone deque of positions */
    if (!buf || !dq) { free(buf); free(dq); return ORC_E_NOMEM; }
    hmap h;
    if (hmap_init(&h, total / 4 + 1024)) { free(buf); free(dq); return ORC_E_NOMEM; }
    for (size_t r = 0; r < n_rec; ++r) {
        const uint8_t* s = bases + rec_off[r];
        int64_t L = (int64_t)(rec_off[r + 1] - rec_off[r]);
        int64_t m = orc_nthash_iter(s, L, k, canonical, circular, buf);
        if (m < 0) { hmap_free(&h); free(buf); free(dq); return m; }
        int64_t head = 0, tail = 0;
        for (int64_t i = 0; i < m; ++i) {
            while (tail > head && buf[dq[tail - 1]] >= buf[i]) --tail;
            dq[tail++] = i;
            if (dq[head] <= i - w) ++head;
            if (i >= w - 1) {
                uint64_t mn = buf[dq[head]];
                if (scaled && mn > max_hash) continue;
                int fo; hmap_slot(&h, mn, 1, &fo);
            }
        }
    }
    int64_t n = 0;
    for (size_t i = 0; i < h.cap; ++i) if (h.used[i]) { if ((size_t)n >= out_cap) { n = ORC_E_ARG; break; } out_keys[n++] = h.keys[i]; }
    if (n > 0) orc_sort_u64(out_keys, (size_t)n, threads);
    hmap_free(&h); free(buf); free(dq);
    return n;
}

/* ------------------------------------------------------------------------- */
/* Synthetic-input generators of SURVEY.md section 8(d) (CPU side, counter-based) */
/* ------------------------------------------------------------------------- */
uint64_t orc_sm64(uint64_t x) { return mix64(x); }

/* U(j;N,S) = j*W + (sm64(S+j) mod W), W = floor(2^62/N) */
void orc_universe(uint64_t j0, size_t count, uint64_t N, uint64_t S, uint64_t* out) {
    uint64_t W = (1ull << 62) / N;
    for (size_t i = 0; i < count; ++i) { uint64_t j = j0 + i; out[i] = j * W + (mix64(S + j) % W); }
}

/* file f of the C3/C5 generators: U_j for every j in [j0, j0+count) with bit f of
 * sm64(T+j) set.  Returns the number written. */
size_t orc_member_file(uint64_t j0, size_t count, uint64_t N, uint64_t S, uint64_t T, int f, uint64_t* out) {
    uint64_t W = (1ull << 62) / N;
    size_t m = 0;
    for (size_t i = 0; i < count; ++i) {
        uint64_t j = j0 + i;
        if ((mix64(T + j) >> f) & 1u) out[m++] = j * W + (mix64(S + j) % W);
    }
    return m;
}

/* Digests of inter / diff / union over the files 0..nf-1 of the C3 generator, for the universe indices [j0, j0+count):
 * the universe keys U_j are strictly increasing, so the three results are the U_j whose membership byte (bits 0..nf-1 of
 * sm64(T+j)) is all ones / is exactly bit 0 / is non-zero.  out[9] = {count, sum, xor} of inter, diff, union (sum mod
 * 2^64).  A size-independent check of FULL results at BASELINE sizes (bench.py, tests): the set operations themselves
 * are checked against orc_inter / orc_diff / orc_union on windows. */
void orc_c3_digest(uint64_t j0, uint64_t count, uint64_t N, uint64_t S, uint64_t T, int nf, uint64_t* out) {
    const uint64_t W = (1ull << 62) / N;
    const uint64_t all = nf >= 64 ? ~0ull : ((1ull << nf) - 1);
    uint64_t ci = 0, si = 0, xi = 0, cd = 0, sd = 0, xd = 0, cu = 0, su = 0, xu = 0;
#pragma omp parallel for reduction(+ : ci, si, cd, sd, cu, su) reduction(^ : xi, xd, xu) schedule(static)
    for (uint64_t i = 0; i < count; ++i) {
        const uint64_t j = j0 + i;
        const uint64_t m = mix64(T + j) & all;
        if (!m) continue;
        const uint64_t key = j * W + (mix64(S + j) % W);
        cu++; su += key; xu ^= key;
        if (m == all) { ci++; si += key; xi ^= key; }
        if (m == 1) { cd++; sd += key; xd ^= key; }
    }
    out[0] = ci; out[1] = si; out[2] = xi;
    out[3] = cd; out[4] = sd; out[5] = xd;
    out[6] = cu; out[7] = su; out[8] = xu;
}

/* C2 keys: sm64(S+i) >> 2 */
void orc_random_keys(uint64_t i0, size_t count, uint64_t S, uint64_t* out) {
    for (size_t i = 0; i < count; ++i) out[i] = mix64(S + i0 + i) >> 2;
}

/* C4 bases: base i of record r = "ACGT"[(sm64(S + (r<<32) + i/32) >> (2*(i%32))) & 3] */
void orc_synth_bases(uint64_t r, uint64_t i0, size_t count, uint64_t S, uint8_t* out) {
    for (size_t t = 0; t < count; ++t) {
        uint64_t i = i0 + t;
        out[t] = (uint8_t)"ACGT"[(mix64(S + (r << 32) + (i >> 5)) >> (2 * (i & 31))) & 3u];
    }
}
