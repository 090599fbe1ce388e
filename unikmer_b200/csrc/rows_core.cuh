// rows_core.cuh -- the arithmetic of the row-based N-way union tile (nunion.cu), written so that the same functions run
// inside the CUDA kernel and, compiled by g++, inside the host model of the CPU test-suite (tests/host/rows_model.cpp).
//
// What it replaces: the hash-set union of union.go:186-208 + the key sort of union.go:260-305 -- a tile of the key space
// is brought into shared memory from all N files and merged there in log2(N) levels of two-way merges.  Compared with
// the per-thread sequential merge walks of nway_core.cuh (one data-dependent shared-memory load per key and level, 2.1x
// bank conflicts, a dependent chain of 13 loads per thread) a level here is, per thread ("row" of RW_E = 16 merged keys):
//
//   1. merge-path split of the row's diagonal (binary search, two loads per probe);
//   2. GATHER the row's <= 16 inputs -- a run of A ascending, then a run of B descending -- into registers with 16
//      independent loads.  Rows are stored with one pad slot per 16 keys and the B run of a pair is stored REVERSED at
//      an address congruent to its A run, so that the 16 inputs of a row sit on 16 consecutive shared-memory words
//      (mod 16): lane l reads its inputs rotated by l and every round of the gather hits 16 different banks;
//   3. sort them with a BITONIC MERGE network in registers (32 compare-exchanges, no memory traffic, full ILP): a run
//      ascending followed by a run descending is bitonic under any rotation, so the rotated order needs no fix-up;
//   4. write the 16 sorted keys as row j of the output run (stride 17: conflict-free), the pad slot repeats the last
//      key -- duplicates are harmless to a union and keep every run sorted.
//
// The last level keeps the sorted keys in registers and flags the first key of every run of equal keys.
#pragma once
#include <stdint.h>

#include "nway_core.cuh"

constexpr int RW_E = 16;            // merged keys per row (= per thread and level)
constexpr int RW_ROWS_PER_WARP = 31;  // lane 31 only computes the end split of the warp's last row

// a two-way merge of a level.  Element i of A is src[a_off + i], element i of B is src[b_off + b_dir * i] (b_dir = +1
// only on level 1, whose inputs are the files' segments as TMA delivered them), row j of the result starts at
// dst[d_off + d_dir * 17 * j].
struct RwPair {
    int a_off, a_len;
    int b_off, b_dir, b_len;
    int d_off, d_dir;
    int row0;  // first row of this pair among the rows of the level
};

template <int NWAY>
struct RwGeom {
    static constexpr int LEVELS = NWAY == 8 ? 3 : NWAY == 4 ? 2 : 1;
    int n[NWAY];    // segment lengths (slot)
    int off[NWAY];  // first element of segment f in the slot
    int tot;        // input keys of the tile
    RwPair pair[NWAY - 1];  // level 1 pairs first (NWAY/2), then level 2, ...; the last entry is the final merge
    int rows[LEVELS];       // rows of every level
};

template <int NWAY>
NW_HD int rw_pair0(int l) {  // index of the first pair of level l (1-based)
    int p = 0, m = NWAY / 2;
    for (int i = 1; i < l; ++i) { p += m; m >>= 1; }
    return p;
}

NW_HD int rw_phys_len(int n) { return n + n / RW_E; }  // n keys stored as rows of 16 + one pad slot per full row

// Fill pair[] and rows[] from n[] / off[].  Level 1 reads the slot, every later level reads the rows the level before
// wrote (level 1 -> X, level 2 -> slot, ...); a run that will be the B side of its next merge is written reversed, at an
// offset that makes (b_off - a_off) = 15 (mod 16): with a + b = 0 (mod 16) at every row start the gather is conflict-free.
// Returns the largest buffer extent any level writes (elements), so the caller can check it against the capacity.
template <int NWAY>
NW_HD int rw_build_tables(RwGeom<NWAY>* g) {
    constexpr int LEVELS = RwGeom<NWAY>::LEVELS;
    int r_off[NWAY], r_dir[NWAY], r_len[NWAY];
    int tot = 0;
    for (int f = 0; f < NWAY; ++f) {
        r_off[f] = g->off[f];
        r_dir[f] = 1;
        r_len[f] = g->n[f];
        tot += g->n[f];
    }
    g->tot = tot;
    int runs = NWAY, extent = 0;
#pragma unroll
    for (int l = 1; l <= LEVELS; ++l) {
        const int p0 = rw_pair0<NWAY>(l);
        int cursor = 0, row = 0;
#pragma unroll
        for (int m = 0; m < NWAY / 2; ++m) {
            if (m >= runs / 2) break;
            RwPair& pr = g->pair[p0 + m];
            pr.a_off = r_off[2 * m];
            pr.a_len = r_len[2 * m];
            pr.b_off = r_off[2 * m + 1];
            pr.b_dir = r_dir[2 * m + 1];
            pr.b_len = r_len[2 * m + 1];
            pr.row0 = row;
            const int nk = pr.a_len + pr.b_len;
            const int plen = rw_phys_len(nk);
            row += (nk + RW_E - 1) / RW_E;
            if (m & 1) {
                // reversed: element 0 at the high end; congruent to the A run it will be merged with (the run before it)
                int off0 = cursor + plen - 1;
                const int want = (r_off[m - 1] + 15) & 15;  // r_off[m - 1]: already the OUTPUT run m - 1 of this level
                const int pad = (want - off0) & 15;
                off0 += pad;
                pr.d_off = off0;
                pr.d_dir = -1;
                cursor = off0 + 1;
            } else {
                pr.d_off = cursor;
                pr.d_dir = 1;
                cursor += plen;
            }
            // run m of the next level (entries 2m, 2m + 1 are consumed; m <= 2m keeps unread entries intact)
            r_off[m] = pr.d_off;
            r_dir[m] = pr.d_dir;
            r_len[m] = plen;
        }
        g->rows[l - 1] = row;
        if (l < LEVELS && cursor > extent) extent = cursor;
        runs >>= 1;
    }
    return extent;
}

// number of A elements among the first `diag` merged elements; ties: A first
NW_HD int rw_merge_path(const uint64_t* A, int na, const uint64_t* B, int bdir, int nb, int diag) {
    int lo = diag > nb ? diag - nb : 0;
    int hi = diag < na ? diag : na;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (A[mid] <= B[bdir * (diag - 1 - mid)]) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// which pair of a level does row `rho` belong to (-1: none), and its index inside the pair
template <int NPAIRS>
NW_HD int rw_find_pair(const RwPair* pr, int total_rows, int rho, int* j) {
    if (rho >= total_rows) return -1;
    int m = 0;
#pragma unroll
    for (int i = 1; i < NPAIRS; ++i) m += (rho >= pr[i].row0) ? 1 : 0;
    *j = rho - pr[m].row0;
    return m;
}

NW_HD void rw_cas(uint64_t& a, uint64_t& b) {  // a <= b afterwards
    const bool sw = b < a;
    const uint64_t lo = sw ? b : a, hi = sw ? a : b;
    a = lo;
    b = hi;
}

// sort a bitonic sequence of 16 (any rotation of: ascending run, descending run)
NW_HD void rw_bitonic16(uint64_t* s) {
#pragma unroll
    for (int d = 8; d >= 1; d >>= 1) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if ((i & d) == 0) rw_cas(s[i], s[i | d]);
    }
}

// Gather the inputs of one row into s[0..16) (rotated by `rot`, 0..15) and sort them.  The row takes A[a, a + na) and
// B[b, b + nb), na + nb <= 16; missing inputs (the last row of a pair) are +inf and sort to the end.
// In circular order e = 0..15: A ascending, then the +inf fill, then B descending (e = 15 is B[b]).
NW_HD void rw_gather_sort(const uint64_t* A, const uint64_t* B, int bdir, int a, int na, int b, int nb, int rot, uint64_t* s) {
    const uint64_t* pa = A + a;               // element e of the A part: pa[e]
    const uint64_t* pb = B + bdir * (b + 15);  // element e of the B part: pb[-bdir * e]
    const int b_first = 16 - nb;              // first e of the B part
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const int e = (r + rot) & 15;
        uint64_t v = ~0ull;
        if (e < na) v = pa[e];
        else if (e >= b_first) v = pb[-bdir * e];
        s[r] = v;
    }
    rw_bitonic16(s);
}

// write the sorted row j (cnt valid keys) of a run: dst0 = address of the run's element 0, ddir = +-1
NW_HD void rw_write_row(uint64_t* dst0, int ddir, int j, int cnt, const uint64_t* s) {
    uint64_t* p = dst0 + ddir * (RW_E + 1) * j;
#pragma unroll
    for (int i = 0; i < 16; ++i)
        if (i < cnt) p[ddir * i] = s[i];
    if (cnt == 16) p[ddir * 16] = s[15];  // the pad slot repeats the last key
}
