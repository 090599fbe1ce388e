// ride_model.cpp -- host model of `inter` / `diff` riding along the last level of the N-way union kernel
// (unikmer_b200/csrc/nway.cu).
//
// Test infrastructure (CPU suite, no GPU).  The kernel never sees which file a merged key came from; it reads the
// number of files holding a key off the length of its run of equal keys, per thread and with bit arithmetic on the
// thread's emit mask and the first eight positions of the next thread (nw_lead_heads / nw_run_candidates in
// nway_core.cuh -- the SAME functions the kernel calls).  This model lays a tile's merged sequence out over the
// kernel's thread ranges (VT keys per thread, a partial last thread, idle threads behind it), runs those functions
// for every thread, applies the second phase (a run of one counts for `diff` only if the key is file 0's) and
// compares the result with set algebra on the files.  The barrier choreography and the atomics are the GPU tests'.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <random>
#include <vector>

#include "../../unikmer_b200/csrc/nway_core.cuh"

static int g_fail = 0;

// one tile: the files' keys are the whole tile (tiles are cut at key boundaries, so runs never cross them)
static bool run_tile(const std::vector<std::vector<uint64_t>>& files, int NT, int VT, const char* what) {
    const int nf = (int)files.size();
    std::vector<uint64_t> seq;
    for (const auto& f : files) seq.insert(seq.end(), f.begin(), f.end());
    std::sort(seq.begin(), seq.end());
    const int tot = (int)seq.size();
    if (tot > NT * VT) return true;  // does not fit one tile of this shape
    // ---- phase 0: what nw_walk_unique leaves in every thread (emit mask over its range) + the published heads ----
    std::vector<unsigned> emit(NT), lead(NT);
    std::vector<int> steps(NT);
    for (int tid = 0; tid < NT; ++tid) {
        int diag = tid * VT, st = tot - diag;
        if (st > VT) st = VT;
        if (st < 0) st = 0;
        unsigned m = 0;
        for (int it = 0; it < st; ++it) {
            const int i = diag + it;
            if (i == 0 || seq[i] != seq[i - 1]) m |= 1u << it;
        }
        emit[tid] = m;
        steps[tid] = st;
        lead[tid] = nw_lead_heads(m, st);
    }
    // ---- phase 1 (between the scan's barriers) + phase 2 (after them) ----
    std::vector<uint64_t> got_i, got_d;
    for (int tid = 0; tid < NT; ++tid) {
        const unsigned next_heads = tid + 1 < NT ? lead[tid + 1] : 0xffu;
        unsigned run_nf, run_one;
        unsigned m = nw_run_candidates(emit[tid], steps[tid], next_heads, nf, &run_nf, &run_one);
        while (m) {
            const int b0 = __builtin_ffs((int)m) - 1;
            m &= m - 1;
            const bool one = (run_one >> b0) & 1u;
            const uint64_t x = seq[tid * VT + b0];
            const bool in0 = std::binary_search(files[0].begin(), files[0].end(), x);  // nw_mark_f0: lower_bound in file 0's segment
            if (!in0) continue;
            (one ? got_d : got_i).push_back(x);
        }
    }
    // ---- expected: set algebra ----
    std::vector<uint64_t> want_i = files[0], want_d = files[0], tmp;
    for (int f = 1; f < nf; ++f) {
        tmp.clear();
        std::set_intersection(want_i.begin(), want_i.end(), files[f].begin(), files[f].end(), std::back_inserter(tmp));
        want_i.swap(tmp);
        tmp.clear();
        std::set_difference(want_d.begin(), want_d.end(), files[f].begin(), files[f].end(), std::back_inserter(tmp));
        want_d.swap(tmp);
    }
    std::sort(got_i.begin(), got_i.end());
    std::sort(got_d.begin(), got_d.end());
    if (got_i != want_i || got_d != want_d) {
        fprintf(stderr, "FAIL %s: nf=%d NT=%d VT=%d tot=%d  inter %zu/%zu  diff %zu/%zu\n", what, nf, NT, VT, tot, got_i.size(), want_i.size(),
                got_d.size(), want_d.size());
        ++g_fail;
        return false;
    }
    return true;
}

int main() {
    std::mt19937_64 rng(20261017);
    long long tiles = 0;
    const int shapes[][2] = {{256, 13}, {128, 17}, {256, 9}, {512, 13}, {8, 9}, {4, 13}};
    for (const auto& sh : shapes) {
        const int NT = sh[0], VT = sh[1];
        for (int nf = 2; nf <= 8; ++nf) {
            for (int trial = 0; trial < 60; ++trial) {
                // a universe small enough that every run length 1..nf occurs, sized so the tile is nearly full, short or tiny
                const int cap = NT * VT;
                const int fill = trial % 3 == 0 ? cap : trial % 3 == 1 ? cap / 2 + (int)(rng() % (cap / 2)) : 1 + (int)(rng() % 40);
                const int kind = trial % 6;
                std::vector<std::vector<uint64_t>> files(nf);
                int total = 0;
                uint64_t key = rng() % 1000;
                while (total < fill) {
                    key += 1 + rng() % 5;
                    unsigned member;
                    if (kind == 4) member = (1u << nf) - 1u;                      // identical files
                    else if (kind == 5) member = 1u << (rng() % nf);              // disjoint files
                    else if (kind == 3) member = (rng() % 4 == 0) ? (1u << nf) - 1u : (unsigned)(rng() & ((1u << nf) - 1u));
                    else member = (unsigned)(rng() & rng() & ((1u << nf) - 1u)) | ((rng() % 3 == 0) ? 1u : 0u);
                    if (member == 0) continue;
                    const int c = __builtin_popcount(member);
                    if (total + c > fill && total > 0) break;
                    for (int f = 0; f < nf; ++f)
                        if (member >> f & 1u) files[f].push_back(key);
                    total += c;
                }
                if (trial % 10 == 7) files[0].clear();       // an empty file 0: nothing in either result
                if (trial % 10 == 8) files[nf - 1].clear();  // an empty last file: the intersection is empty
                run_tile(files, NT, VT, "random");
                ++tiles;
            }
        }
    }
    // runs of nf keys placed at every offset against the thread boundaries (the run continues into the next thread)
    for (int VT : {9, 13, 17}) {
        for (int nf = 2; nf <= 8; ++nf) {
            for (int start = 0; start < 3 * VT; ++start) {
                for (int tail = 0; tail < 3; ++tail) {
                    std::vector<std::vector<uint64_t>> files(nf);
                    uint64_t key = 10;
                    for (int i = 0; i < start; ++i) files[i % 2 ? 0 : nf - 1].push_back(key += 2);  // singletons before
                    key += 2;
                    for (int f = 0; f < nf; ++f) files[f].push_back(key);                           // the full run
                    for (int i = 0; i < tail; ++i) files[0].push_back(key += 2);                    // singletons of file 0 behind
                    run_tile(files, 8, VT, "offsets");
                    ++tiles;
                }
            }
        }
    }
    if (g_fail) {
        fprintf(stderr, "%d tile(s) failed\n", g_fail);
        return 1;
    }
    printf("ride-along model ok: %lld tiles\n", tiles);
    return 0;
}
