// nway_core.cuh -- the arithmetic of the N-way (N <= 8) single-pass union, written so that the same
// functions run inside the CUDA kernel (nway.cu) and, compiled by g++, inside the host model that the
// CPU test-suite drives (tests/host/nway_model.cpp): tile partition by key (multi-sequence selection on
// rank(K) = sum_f lower_bound(F_f, K)), per-tile run tables, merge-path split, the plain two-way walk of the
// inner levels and the de-duplicating walk of the last level.
//
// What it replaces: the hash-set union of union.go:186-208 + the key sort of union.go:260-305 (and, as
// the merge step, mergeChunksFile's heap, util-sort.go:227-606) -- here every input is read ONCE: a tile
// of the key space is brought into shared memory from all N files and merged there in log2(N) levels.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define NW_HD __host__ __device__ __forceinline__
#else
#define NW_HD inline
#endif

// global-memory probe of the partition (a hook so the host tools can count probes and cache lines)
#ifndef NW_PROBE
#define NW_PROBE(ptr) (*(ptr))
#endif

constexpr int NW_MAX = 8;       // files per pass

// ---------------------------------------------------------------------------------------------------
// partition: cut every file at lower_bound(key); rank = total elements below the cut
// ---------------------------------------------------------------------------------------------------
struct NwBound {
    uint64_t key;
    long long rank;
    long long pos[NW_MAX];
};

struct NwFiles {
    const uint64_t* k[NW_MAX];
    long long n[NW_MAX];
    int nf;  // files in use (the others have n = 0)
};

NW_HD long long nw_lower_bound(const uint64_t* F, long long lo, long long hi, uint64_t key) {
    while (lo < hi) {
        const long long mid = lo + ((hi - lo) >> 1);
        if (F[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// mid = the cut at `key`: pos[f] = lower_bound(F_f, key) searched inside [lo.pos[f], hi.pos[f]].
// `frac` = where between lo and hi the cut is expected (files that are subsets of one key distribution are cut at
// about the same fraction of their brackets): every search starts there, gallops out (steps 16, 32, ...) until the
// answer is bracketed and bisects inside.  Compared with bisecting the whole bracket this touches a handful of
// neighbouring cache lines per file instead of one line per level (ncu: the bisection pulled ~15 KB of DRAM per
// boundary, as much as the tile itself).  The eight searches advance in lock step so that their dependent,
// long-latency loads are in flight together.
NW_HD void nw_cut_at(const NwFiles& F, uint64_t key, const NwBound& lo, const NwBound& hi, double frac, NwBound* mid) {
    long long l[NW_MAX], r[NW_MAX], step[NW_MAX];  // answer in [l, r]; step > 0: galloping right, < 0: left, 0: bisecting
    int first[NW_MAX];
    bool busy = false;
#pragma unroll
    for (int f = 0; f < NW_MAX; ++f) {
        l[f] = lo.pos[f];
        r[f] = (f < F.nf) ? hi.pos[f] : lo.pos[f];
        step[f] = 0;
        first[f] = 1;
        busy |= l[f] < r[f];
    }
    while (busy) {
        long long x[NW_MAX];
        uint64_t v[NW_MAX];
#pragma unroll
        for (int f = 0; f < NW_MAX; ++f) {
            if (l[f] < r[f]) {
                if (first[f]) x[f] = l[f] + (long long)(frac * (double)(r[f] - l[f]));
                else if (step[f] > 0) x[f] = l[f] + step[f] - 1;
                else if (step[f] < 0) x[f] = r[f] + step[f];
                else x[f] = l[f] + ((r[f] - l[f]) >> 1);
                if (x[f] < l[f]) x[f] = l[f];
                if (x[f] > r[f] - 1) x[f] = r[f] - 1;
                v[f] = NW_PROBE(F.k[f] + x[f]);
            } else {
                x[f] = 0;
                v[f] = 0;
            }
        }
        busy = false;
#pragma unroll
        for (int f = 0; f < NW_MAX; ++f) {
            if (l[f] < r[f]) {
                const bool below = v[f] < key;  // the answer is right of x
                if (below) l[f] = x[f] + 1;
                else r[f] = x[f];
                if (first[f]) {
                    first[f] = 0;
                    step[f] = below ? 16 : -16;
                } else if (step[f] > 0) {
                    step[f] = below ? step[f] * 2 : 0;  // overshot: the bracket is closed, bisect
                } else if (step[f] < 0) {
                    step[f] = below ? 0 : step[f] * 2;
                }
                busy |= l[f] < r[f];
            }
        }
    }
    mid->key = key;
    mid->rank = 0;
#pragma unroll
    for (int f = 0; f < NW_MAX; ++f) {
        mid->pos[f] = l[f];
        mid->rank += l[f];
    }
}

// Find a cut whose rank is within `tol` of R, given a bracket lo.rank <= R <= hi.rank (lo.key <= hi.key;
// hi may be the end-of-everything boundary: pos = n).  Multi-sequence selection with pivots taken from the
// DATA, so it does not care how the keys are spread over the key space: the pivot is the element at the
// wanted rank fraction of the file with the widest position bracket (even rounds; one to three rounds on real
// inputs) or that bracket's median (odd rounds; bounds the worst case at ~2 * 8 * log2(n) rounds).  Every
// round searches only inside the per-file position brackets, which shrink with the key bracket.  If no cut
// lands within tol (needs a key that occurs more than tol times, i.e. inputs that are not duplicate-free,
// or tol < 8) the upper end of the final bracket is returned; nway_check validates the tile sizes.
NW_HD void nw_refine(const NwFiles& F, long long R, long long tol, NwBound lo, NwBound hi, NwBound* out) {
    if (R - lo.rank <= tol) { *out = lo; return; }
    if (hi.rank - R <= tol) { *out = hi; return; }
    for (int round = 0; round < 1024; ++round) {
        int fs = 0;
        long long w = -1;
#pragma unroll
        for (int f = 0; f < NW_MAX; ++f) {
            const long long wf = hi.pos[f] - lo.pos[f];
            if (wf > w) { w = wf; fs = f; }
        }
        if (w < 2) break;
        long long d;
        if (round & 1) {
            d = w >> 1;
        } else {
            const double fr = (double)(R - lo.rank) / (double)(hi.rank - lo.rank);
            d = (long long)(fr * (double)w);
        }
        if (d < 1) d = 1;
        if (d > w - 1) d = w - 1;
        const double frac = (double)d / (double)w;
        const uint64_t* Fs = F.k[0];
        long long lofs = lo.pos[0];
#pragma unroll
        for (int f = 1; f < NW_MAX; ++f)
            if (f == fs) { Fs = F.k[f]; lofs = lo.pos[f]; }
        NwBound mid;
        nw_cut_at(F, NW_PROBE(Fs + lofs + d), lo, hi, frac, &mid);  // lo.key < mid.key < hi.key: strictly inside the bracket
        const long long err = mid.rank - R;
        if (err >= -tol && err <= tol) { *out = mid; return; }
        if (err < 0) lo = mid;
        else hi = mid;
    }
    *out = hi;
}

// the bracket that holds everything: [cut below every key, cut above every key]
NW_HD void nw_global_bracket(const NwFiles& F, NwBound* lo, NwBound* hi) {
    uint64_t kmin = ~0ull, kmax = 0;
    long long total = 0;
    for (int f = 0; f < NW_MAX; ++f) {
        lo->pos[f] = 0;
        hi->pos[f] = (f < F.nf) ? F.n[f] : 0;
        if (f < F.nf && F.n[f] > 0) {
            const uint64_t a = F.k[f][0], b = F.k[f][F.n[f] - 1];
            if (a < kmin) kmin = a;
            if (b > kmax) kmax = b;
            total += F.n[f];
        }
    }
    if (kmin > kmax) kmin = kmax = 0;  // nothing at all
    lo->key = kmin;  // lower_bound(kmin) = 0 in every file
    lo->rank = 0;
    hi->key = kmax;  // stands for "above kmax": candidate cuts are lo.key < K < kmax
    hi->rank = total;
}

// ---------------------------------------------------------------------------------------------------
// per-tile geometry: where the N segments sit in the slot, and the two-way merges of every level
// ---------------------------------------------------------------------------------------------------
struct NwPair {
    int srcA, lenA, srcB, lenB;  // runs in the source buffer of the level
    int dst;                     // start of the merged run in the destination buffer (dense)
};

template <int NWAY>
struct NwGeom {
    static constexpr int LEVELS = NWAY == 8 ? 3 : NWAY == 4 ? 2 : 1;
    int n[NWAY];    // segment lengths
    int off[NWAY];  // first element of segment f in the slot
    int tot;
    NwPair pair[NWAY - 1];  // level 1 pairs first (NWAY/2), then level 2, ...; the last entry is the final merge
    int tb[NWAY - 1 + 3];   // thread bases: for level l, tb[tbo(l) + m] .. ; one extra end entry per level
};

// index of the first pair / first thread-base entry of level l (1-based)
template <int NWAY>
NW_HD int nw_pair0(int l) {
    int p = 0, m = NWAY / 2;
    for (int i = 1; i < l; ++i) { p += m; m >>= 1; }
    return p;
}
template <int NWAY>
NW_HD int nw_tb0(int l) { return nw_pair0<NWAY>(l) + (l - 1); }

// Fill pair[] and tb[] from n[] / off[].  Level 1 reads the slot (segments at off[]), every later level
// reads the dense output of the level before.
template <int NWAY, int VT>
NW_HD void nw_build_tables(NwGeom<NWAY>* g) {
    constexpr int LEVELS = NwGeom<NWAY>::LEVELS;
    int start[NWAY], len[NWAY];
    int tot = 0;
    for (int f = 0; f < NWAY; ++f) { start[f] = g->off[f]; len[f] = g->n[f]; tot += g->n[f]; }
    g->tot = tot;
    int runs = NWAY;
#pragma unroll
    for (int l = 1; l <= LEVELS; ++l) {
        const int p0 = nw_pair0<NWAY>(l), t0 = nw_tb0<NWAY>(l);
        int dst = 0, tbase = 0;
#pragma unroll
        for (int m = 0; m < NWAY / 2; ++m) {
            if (m >= runs / 2) break;
            NwPair& pr = g->pair[p0 + m];
            pr.srcA = start[2 * m]; pr.lenA = len[2 * m];
            pr.srcB = start[2 * m + 1]; pr.lenB = len[2 * m + 1];
            pr.dst = dst;
            g->tb[t0 + m] = tbase;
            const int s = pr.lenA + pr.lenB;
            start[m] = dst; len[m] = s;  // run m of the next level (m <= 2m: not yet consumed entries stay intact)
            dst += s;
            tbase += (s + VT - 1) / VT;
        }
        g->tb[t0 + runs / 2] = tbase;
        runs >>= 1;
    }
}

// ---------------------------------------------------------------------------------------------------
// merge path + walks
// ---------------------------------------------------------------------------------------------------
// number of A elements among the first `diag` merged elements; ties: A first
NW_HD int nw_merge_path(const uint64_t* A, int na, const uint64_t* B, int nb, int diag) {
    int lo = diag > nb ? diag - nb : 0;
    int hi = diag < na ? diag : na;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (A[mid] <= B[diag - 1 - mid]) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// the same split, found from the proportional guess diag * na / (na + nb): gallop out (steps 4, 8, ...) until it is
// bracketed, then bisect.  Runs that interleave evenly (subsets of one key distribution) need 6-8 probes
// instead of log2(n) + 1; every probe is two shared-memory loads, the scarce resource of the merge levels.
NW_HD int nw_merge_path_g(const uint64_t* A, int na, const uint64_t* B, int nb, int diag) {
    int lo = diag > nb ? diag - nb : 0;
    int hi = diag < na ? diag : na;
    if (lo >= hi) return lo;
    // P(x) := A[x] > B[diag-1-x] is monotone false..true on [lo, hi); the answer is the first true (or hi)
    int g = (int)(((long long)diag * na) / (na + nb));
    if (g < lo) g = lo;
    if (g > hi - 1) g = hi - 1;
    int step = 4;
    if (A[g] > B[diag - 1 - g]) {
        hi = g;
        while (hi > lo) {
            const int x = hi - step < lo ? lo : hi - step;
            if (A[x] > B[diag - 1 - x]) { hi = x; step <<= 1; }
            else { lo = x + 1; break; }
        }
    } else {
        lo = g + 1;
        while (lo < hi) {
            const int x = lo + step - 1 > hi - 1 ? hi - 1 : lo + step - 1;
            if (!(A[x] > B[diag - 1 - x])) { lo = x + 1; step <<= 1; }
            else { hi = x; break; }
        }
    }
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (A[mid] > B[diag - 1 - mid]) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}

// ---------------------------------------------------------------------------------------------------
// inter / diff riding along the last level (nway.cu): run lengths from emit masks
// ---------------------------------------------------------------------------------------------------
// A thread's last-level range holds `steps` merged keys; bit i of `emitmask` = key i starts a run of equal keys (a run
// has one key per file that holds the k-mer: the inputs are duplicate-free).  Runs are at most nf <= 8 keys long, so one
// that starts in this range ends in it or in the first 8 positions of the next thread's range.
// What a thread publishes for its predecessor: the run heads among its first 8 positions; positions past the end of
// its range count as heads (the tile ends there, so does any run).
NW_HD unsigned nw_lead_heads(unsigned emitmask, int steps) { return (emitmask | (~0u << steps)) & 0xffu; }

// Candidates of the range: bit i of *run_nf = a run of exactly nf keys starts at i (the key is in every file),
// bit i of *run_one = a run of one key starts at i (the key is in one file only).  next_heads = nw_lead_heads of
// the next thread (0xff behind the last thread).  Returns the union of the two masks.
NW_HD unsigned nw_run_candidates(unsigned emitmask, int steps, unsigned next_heads, int nf, unsigned* run_nf, unsigned* run_one) {
    const unsigned range = steps >= 32 ? ~0u : ((1u << steps) - 1u);
    // the heads of this range followed by those of the next thread's first positions: a head at i with the next head
    // at i + 1 is a run of one, with the next head at i + nf a run of nf (no loop over the heads)
    const unsigned ext = (emitmask & range) | (next_heads << steps);
    unsigned rn = ext & (ext >> nf);
    for (int k = 1; k < nf; ++k) rn &= ~(ext >> k);
    const unsigned ro = ext & (ext >> 1);
    *run_nf = rn;
    *run_one = ro;
    return (rn | ro) & range;
}

// which pair of a level does thread `tid` work on (-1: none), and its index inside the pair
template <int NPAIRS>
NW_HD int nw_find_pair(const int* tb, int tid, int* j) {
    if (tid >= tb[NPAIRS]) return -1;
    int m = 0;
#pragma unroll
    for (int i = 1; i < NPAIRS; ++i) m += (tid >= tb[i]) ? 1 : 0;
    *j = tid - tb[m];
    return m;
}

// plain two-way merge of `steps` (<= VT) elements starting at split (a, b): dst[0..steps)
// (A[na] / B[nb] may be read but are never used: every buffer carries padding behind its last run)
template <int VT>
NW_HD void nw_walk_plain(const uint64_t* A, int na, const uint64_t* B, int nb, int a, int b, int steps, uint64_t* dst) {
    const uint64_t* pa = A + a;
    const uint64_t* pb = B + b;
    const uint64_t* const ea = A + na;
    const uint64_t* const eb = B + nb;
    uint64_t ka = *pa, kb = *pb;
    if (steps == VT) {  // every thread but the last of a pair
#pragma unroll
        for (int it = 0; it < VT; ++it) {
            const bool takeA = (pa < ea) && (pb >= eb || ka <= kb);
            dst[it] = takeA ? ka : kb;
            if (takeA) ka = *++pa;
            else kb = *++pb;
        }
    } else {
        for (int it = 0; it < steps; ++it) {
            const bool takeA = (pa < ea) && (pb >= eb || ka <= kb);
            dst[it] = takeA ? ka : kb;
            if (takeA) ka = *++pa;
            else kb = *++pb;
        }
    }
}

// last level: merge into registers, flag the first element of every run of equal keys.
// Returns the emit mask (bit it = outk[it] is a new distinct key).
template <int VT>
NW_HD unsigned nw_walk_unique(const uint64_t* A, int na, const uint64_t* B, int nb, int a, int b, int steps, uint64_t* outk) {
    bool has_prev = (a > 0) || (b > 0);
    uint64_t prev = 0;
    if (a > 0) prev = A[a - 1];
    if (b > 0 && B[b - 1] > prev) prev = B[b - 1];
    const uint64_t* pa = A + a;
    const uint64_t* pb = B + b;
    const uint64_t* const ea = A + na;
    const uint64_t* const eb = B + nb;
    uint64_t ka = *pa, kb = *pb;
    unsigned mask = 0;
    if (steps == VT) {
#pragma unroll
        for (int it = 0; it < VT; ++it) {
            const bool takeA = (pa < ea) && (pb >= eb || ka <= kb);
            const uint64_t k = takeA ? ka : kb;
            outk[it] = k;
            mask |= ((!has_prev || k != prev) ? 1u : 0u) << it;
            prev = k;
            has_prev = true;
            if (takeA) ka = *++pa;
            else kb = *++pb;
        }
    } else {
#pragma unroll
        for (int it = 0; it < VT; ++it) {
            const bool takeA = (pa < ea) && (pb >= eb || ka <= kb);
            const uint64_t k = takeA ? ka : kb;
            outk[it] = k;
            const bool emit = (it < steps) && (!has_prev || k != prev);
            mask |= (emit ? 1u : 0u) << it;
            prev = k;
            has_prev = true;
            if (it < steps) {
                if (takeA) ka = *++pa;
                else kb = *++pb;
            }
        }
    }
    return mask;
}
