/*
 * ukm.h -- C ABI of libukm.so, the B200 (sm_100a) k-mer set-operations engine.
 *
 * The reference (shenwei356/unikmer, pure Go, CGO_ENABLED=0) has no FFI/plugin
 * boundary of its own; this header DEFINES the drop-in boundary as the batch calls
 * a cgo shim makes in place of the reference's per-k-mer inner loops.  Each entry
 * point names the reference call site(s) it replaces (paths relative to
 * unikmer/cmd/ of the reference; see SURVEY.md 8(b) and INTEGRATION.md for the
 * cgo stub).  Everything above the boundary (flags, file lists, header checks,
 * .unik (de)serialisation, gzip, logging) stays in host code.
 *
 * Conventions
 *  - plain C types only; no CUDA/torch types in signatures (streams travel as void*).
 *  - every call returns 0 (UKM_OK) or a negative ukm_status; the message is in
 *    ukm_last_error(ctx).  The Go shim maps non-zero to checkError (util-cli.go:39-44).
 *  - a ukm_ctx owns one GPU, one CUDA stream and a stream-ordered device arena.
 *    It is single-owner: one call at a time (like one unikmer command).  One process
 *    per GPU; multi-GPU runs shard by key range above this ABI (ukm_partition_sorted
 *    + one NCCL all-to-all, see unikmer_b200/dist.py).
 *  - buffers are caller-owned and are never retained past the call (cgo rule).
 *    `where` says where a span's memory lives.  UKM_DEVICE spans chain operations
 *    without PCIe traffic.
 *  - set operations require every input to be sorted ascending and duplicate-free
 *    (what `unikmer count -s` / `sort -u` write, and what inter.go:139, diff.go:115,
 *    common.go:166 demand via the header flag).  With UKM_F_VALIDATE a violation is
 *    reported as UKM_E_NOT_SORTED_UNIQUE (the reference's duplicate-dependent quirks,
 *    SURVEY.md Appendix B-4, B-7, are not reproduced); without it the flag is trusted,
 *    exactly like the reference trusts the file header.
 */
#ifndef UKM_H
#define UKM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ukm_ctx ukm_ctx;

typedef enum ukm_status {
    UKM_OK = 0,
    UKM_E_ARG = -1,               /* bad argument */
    UKM_E_CUDA = -2,              /* CUDA runtime error (message has the details) */
    UKM_E_NOMEM = -3,             /* device or host allocation failed */
    UKM_E_CAPACITY = -4,          /* out->cap too small; out->n holds the required size */
    UKM_E_NOT_SORTED_UNIQUE = -5, /* an input that must be sorted+unique is not */
    UKM_E_ILLEGAL_BASE = -6,      /* kmers.ErrIllegalBase (count.go:363-366) */
    UKM_E_NO_TAXONOMY = -7,       /* op needs LCA but ukm_set_taxonomy was not called */
    UKM_E_PANIC = -8,             /* the reference panics on this input (inter.go:208, B-3) */
    UKM_E_INTERNAL = -9           /* kernel watchdog / internal invariant */
} ukm_status;

typedef enum ukm_where { UKM_HOST = 0, UKM_HOST_PINNED = 1, UKM_DEVICE = 2 } ukm_where;

/* sort.go:482-573 / util-sort.go:35-190 fold variants */
typedef enum ukm_fold_mode {
    UKM_FOLD_PLAIN = 0,          /* copy */
    UKM_FOLD_UNIQUE = 1,         /* sort -u: first of each run, taxid = LCA over the run */
    UKM_FOLD_REPEATED_FINAL = 2, /* sort -d: codes with multiplicity >= 2, once */
    UKM_FOLD_REPEATED_CHUNK = 3  /* dumpCodes*2File -d: every code once, repeated ones twice */
} ukm_fold_mode;

/* op flags */
#define UKM_F_TAXID 1u         /* hasTaxid: carry per-k-mer taxids, LCA-fold on collisions */
#define UKM_F_MIX_TAXID 2u     /* inter --mix-taxid rule (inter.go:229-236) */
#define UKM_F_COMPARE_TAXID 4u /* diff -t keep rule (diff.go:361-364, 406-409) */
#define UKM_F_CANONICAL 8u     /* count -K */
#define UKM_F_HASHED 16u       /* count -H: ntHash v1 instead of the 2-bit code */
#define UKM_F_CIRCULAR 32u     /* count --circular */
#define UKM_F_SCALED 64u       /* count -D: keep code <= max_hash (count.go:373) */
#define UKM_F_VALIDATE 128u    /* set ops: verify every input is sorted + duplicate-free first (one extra read);
                                  without it the header flag is trusted, as the reference does (inter.go:139) */
#define UKM_F_SHARD 256u       /* inter: the spans are key-range SLICES of files (multi-GPU shards, streamed key ranges), not
                                  whole files: the reference's whole-file quirks are off -- an empty first span gives an empty
                                  result instead of UKM_E_PANIC (inter.go:208), an empty later span empties the result instead
                                  of keeping the current set (inter.go:211-215, B-3) -- so that the concatenated per-slice
                                  results equal the whole-file result.  The caller applies the quirks once, on the file sizes. */

/* One k-mer stream: what unik.Reader.ReadCodeWithTaxid yields for a file, as arrays.
 * Mirrors []uint64 / []CodeTaxid (kmers.go:24-46) in SoA form. */
typedef struct ukm_span {
    uint64_t* keys;        /* n codes */
    uint32_t* taxids;      /* n taxids, or NULL */
    uint32_t global_taxid; /* used for every code when taxids == NULL and UKM_F_TAXID is set
                              (unik header global taxid, README.md:169-171) */
    size_t n;              /* in: element count; out: result count */
    size_t cap;            /* out spans: capacity of keys/taxids in elements */
    int where;             /* ukm_where */
    int sorted;            /* in: header flag reader.IsSorted() (diff.go subject files) */
} ukm_span;

/* per-kernel-family timing collected when ukm_stats_enable(ctx,1) */
typedef struct ukm_kernel_stat {
    char name[48];
    uint64_t launches;
    double ms;         /* sum of CUDA-event durations on the ctx stream */
    double algo_bytes; /* sum of algorithmic bytes the launches moved */
} ukm_kernel_stat;

/* ---- context ----------------------------------------------------------------- */
ukm_ctx* ukm_create(int device);
void ukm_destroy(ukm_ctx* ctx);
const char* ukm_last_error(ukm_ctx* ctx); /* valid until the next call on ctx; ctx may be NULL */
const char* ukm_version(void);
int ukm_set_stream(ukm_ctx* ctx, void* cuda_stream); /* run on the caller's cudaStream_t */
void* ukm_get_stream(ukm_ctx* ctx);
int ukm_sync(ukm_ctx* ctx);

void* ukm_alloc_pinned(size_t bytes);
void ukm_free_pinned(void* p);
void* ukm_alloc_device(ukm_ctx* ctx, size_t bytes);
int ukm_free_device(ukm_ctx* ctx, void* p);
int ukm_copy(ukm_ctx* ctx, void* dst, int dst_where, const void* src, int src_where, size_t bytes);

/* kernels launched by this context so far (bench.py's gpu_launches) */
uint64_t ukm_launch_count(ukm_ctx* ctx);
int ukm_stats_enable(ukm_ctx* ctx, int on);
int ukm_stats_reset(ukm_ctx* ctx);
int ukm_stats_get(ukm_ctx* ctx, ukm_kernel_stat* out, int cap, int* n);

/* ---- taxonomy: loadTaxonomy (util.go:119-171) -> taxondb.LCA (14 call sites) --- */
/* parent[t] for t in [0,n): 0 = unknown taxid, root has parent[t] == t.  merged_from ->
 * merged_to is merged.dmp.  Copies everything; the arrays may be freed on return. */
int ukm_set_taxonomy(ukm_ctx* ctx, const uint32_t* parent, size_t n, const uint32_t* merged_from,
                     const uint32_t* merged_to, size_t n_merged);
/* out[i] = LCA(a[i], b[i]) with the device function the folds use (test hook for a12). */
int ukm_lca_batch(ukm_ctx* ctx, const uint32_t* a, const uint32_t* b, size_t n, uint32_t* out, int where);

/* ---- sort ---------------------------------------------------------------------- */
/* sortutil.Uint64s(m): sort.go:274,337,463; union.go:274,295; diff.go:587; common.go:344;
 * count.go:581; split.go:311,383.  In place, ascending.  key_bits = number of significant
 * low bits (2k for k-mer codes, 64 for hashes; 0 means 64). */
int ukm_sort_u64(ukm_ctx* ctx, uint64_t* keys, size_t n, int key_bits, int where);
/* sorts.Quicksort(CodeTaxidSlice(mt)): sort.go:268,331,457; split.go:305.  SoA, stable
 * (the reference's tie order is undefined, kmers.go:33-46). */
int ukm_sort_pairs(ukm_ctx* ctx, uint64_t* keys, uint32_t* taxids, size_t n, int key_bits, int where);
/* same on Go's 16-byte []CodeTaxid AoS {uint64 code; uint32 taxid; 4 B pad}, host memory. */
int ukm_sort_codetaxid16(ukm_ctx* ctx, void* aos16, size_t n, int key_bits);

/* ---- fold of a sorted slice: sort.go:482-573, util-sort.go:35-190 --------------- */
int ukm_fold_sorted(ukm_ctx* ctx, int mode, const ukm_span* in, unsigned flags, ukm_span* out);

/* ---- N-way operations on sorted, duplicate-free streams --------------------------- */
/* mergeChunksFile (util-sort.go:227-606): k-way merge + fold(mode); inputs sorted, duplicates allowed. */
int ukm_merge_sorted(ukm_ctx* ctx, int mode, const ukm_span* in, int n_in, unsigned flags, ukm_span* out);
/* union.go:186-208 + ordered emit 260-305 (the `-s` contract). */
int ukm_union(ukm_ctx* ctx, const ukm_span* in, int n_in, unsigned flags, ukm_span* out);
/* inter.go:188-286, iterated in file order; flags: UKM_F_TAXID | UKM_F_MIX_TAXID. */
int ukm_inter(ukm_ctx* ctx, const ukm_span* in, int n_in, unsigned flags, ukm_span* out);
/* diff.go:136-146,341-515 + `-s` emit 566-594; in[0] sorted; subjects sorted or not
 * (in[i].sorted == 0: sorted on the device first).  flags: UKM_F_TAXID | UKM_F_COMPARE_TAXID.
 * Two deliberate deviations from the reference's behaviour, both reference bugs:
 *  - an EMPTY sorted subject is skipped ("subtract nothing"); the reference's worker leaves its loop there and
 *    drops its remaining files (diff.go:387-392, quirk B-5, depends on -j and scheduling);
 *  - a sorted subject that follows an UNSORTED one is subtracted from the current result; the reference walks
 *    it against a slice that still holds the keys the unsorted subject removed from its map (diff.go:341-367 vs
 *    380-435), so removed keys can reappear.  The result here is always file0 minus the union of the subjects. */
int ukm_diff(ukm_ctx* ctx, const ukm_span* in, int n_in, unsigned flags, ukm_span* out);
/* common.go:220-283,329-354; threshold as computed at common.go:93-105. */
int ukm_common(ukm_ctx* ctx, const ukm_span* in, int n_in, unsigned flags, uint16_t threshold, ukm_span* out);

/* ---- several operations over the same files, streamed from host memory ---------------- */
/* The reference runs `unikmer inter`, `diff`, `union` as separate commands, each reading every .unik file again
 * (inter.go:188-203, diff.go:136-146 + 380-435, union.go:186-208).  Host-resident inputs make PCIe the bound of a
 * GPU step, so this call runs ANY SUBSET of the three over the same inputs with every input byte uploaded ONCE: the
 * key space is cut into ranges (all three are key-local), the slices of range c+1 are uploaded while the operations
 * run on range c and the results of range c-1 are downloaded.  outs[k] receives the result of ops[k]; results are
 * exactly those of ukm_inter / ukm_diff / ukm_union on the whole files (the whole-file rules -- empty first file,
 * empty later file -- are applied on the file sizes).  Operations asked for together share passes over the inputs
 * (per key range when streamed, over the whole files when the inputs are DEVICE spans): with a union in the list,
 * inter and diff come out of the union's pass (its last merge level sees how many files hold every key); inter and
 * diff without a union share one filter pass (after the first subject every key of file 0 can only still belong to
 * one of the two results).  Up to eight files per pass; more files, taxids or inputs that are not duplicate-free take
 * the single-operation paths.
 * Keys only (flags: 0, UKM_F_VALIDATE, UKM_F_SHARD for device-resident key-range slices); inputs all HOST /
 * HOST_PINNED (pinned memory is what lets the copies overlap: ukm_alloc_pinned) or all DEVICE (nothing is streamed).
 * ukm_inter / ukm_diff / ukm_union take the same streamed path on their own when every input is in host memory. */
typedef enum ukm_setop { UKM_OP_INTER = 0, UKM_OP_DIFF = 1, UKM_OP_UNION = 2 } ukm_setop;
int ukm_setops_stream(ukm_ctx* ctx, const ukm_span* in, int n_in, const int* ops, int n_ops, unsigned flags,
                      ukm_span* outs);

/* ---- count: count.go:314-322 iterators, 355-437 inner loop, 531-595 sort+emit ------ */
/* bases = concatenated records (line breaks stripped, as bio/seqio/fastx yields them),
 * record r = bases[rec_off[r] .. rec_off[r+1]).  Writes the distinct codes ascending. */
int ukm_count_seq(ukm_ctx* ctx, const uint8_t* bases, const uint64_t* rec_off, size_t n_rec, int k,
                  unsigned flags, uint64_t max_hash, int where, ukm_span* out);
/* count -H -W w: sketches.NewMinimizerSketch / NextMinimizer (count.go:100-114, 316-317, 358-359), then the same
 * scaled filter, dedup and sort as ukm_count_seq.  The codes are the distinct minima of every full window of w
 * consecutive ntHash values of a record (pinned by analysis/distance/README.md:8: 549 963 on MG1655, k=31, w=15).
 * UKM_F_HASHED is implied; flags: UKM_F_CANONICAL | UKM_F_CIRCULAR | UKM_F_SCALED. */
int ukm_count_minimizer(ukm_ctx* ctx, const uint8_t* bases, const uint64_t* rec_off, size_t n_rec, int k, int w,
                        unsigned flags, uint64_t max_hash, int where, ukm_span* out);
/* the iterator alone (sketches.NextKmer / NextHash, count.go:361,363; `count --linear`):
 * every k-mer code / hash in record-then-position order, no dedup, no sort. */
int ukm_kmers_seq(ukm_ctx* ctx, const uint8_t* bases, const uint64_t* rec_off, size_t n_rec, int k,
                  unsigned flags, uint64_t max_hash, int where, ukm_span* out);

/* ---- multi-GPU key-range sharding ------------------------------------ */
/* offsets[0..n_split+1]: offsets[0]=0, offsets[i+1] = lower_bound(in->keys, splitters[i]),
 * offsets[n_split+1] = n.  offsets and splitters are HOST arrays. */
int ukm_partition_sorted(ukm_ctx* ctx, const ukm_span* in, const uint64_t* splitters, int n_split,
                         uint64_t* offsets);
/* 0 if keys are strictly increasing, UKM_E_NOT_SORTED_UNIQUE otherwise. */
int ukm_check_sorted_unique(ukm_ctx* ctx, const ukm_span* in);

/* ---- synthetic inputs of SURVEY.md 8(d) (bench / test helpers, device output) ------- */
/* C2: out[i] = sm64(seed + i0 + i) >> 2 */
int ukm_synth_random_keys(ukm_ctx* ctx, uint64_t i0, size_t count, uint64_t seed, uint64_t* d_out);
/* C3/C5: U(j;N,S) for j in [j0, j0+count) with bit f of sm64(T+j) set (f < 0: all j).
 * d_out needs room for `count` keys; *n_out receives the number written. */
int ukm_synth_member_file(ukm_ctx* ctx, uint64_t j0, size_t count, uint64_t N, uint64_t S, uint64_t T,
                          int f, uint64_t* d_out, size_t* n_out);
/* C4: base i of record r = "ACGT"[(sm64(S + (r<<32) + i/32) >> (2*(i%32))) & 3] */
int ukm_synth_bases(ukm_ctx* ctx, uint64_t r, uint64_t i0, size_t count, uint64_t S, uint8_t* d_out);

#ifdef __cplusplus
}
#endif
#endif /* UKM_H */
