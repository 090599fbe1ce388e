// nway_model.cpp -- host model of the N-way union tile kernel (unikmer_b200/csrc/nway.cu).
//
// Test infrastructure (CPU suite, no GPU): compiles the SAME arithmetic the kernel uses
// (unikmer_b200/csrc/nway_core.cuh: partition by multi-sequence selection, per-tile merge tables,
// merge-path split, plain and de-duplicating walks) with g++ and replays the kernel's per-thread
// schedule sequentially -- loader geometry, level 1 slot -> X, level 2 X -> slot, last level +
// unique flags, scan, staging -- then compares with std::set_union.  What it cannot cover is the
// asynchronous choreography (mbarriers, TMA, named barriers); the GPU parity tests do that.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <random>
#include <vector>

#include "../../unikmer_b200/csrc/nway_core.cuh"

template <int NWAY, int NT, int VT>
struct Shape {
    static constexpr int CAP = (NT - NWAY / 2) * VT;
    static constexpr int SLOT_E = (CAP + 2 * NWAY + 8 + 1) & ~1;
    static constexpr int X_E = (CAP + 8 + 1) & ~1;
    static constexpr int TILE = (CAP * 16 / 17) & ~31;
    static constexpr int LEVELS = NwGeom<NWAY>::LEVELS;
};

static int g_fail = 0;
#define CHECK(c, ...)                                  \
    do {                                               \
        if (!(c)) {                                    \
            fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); \
            fprintf(stderr, __VA_ARGS__);              \
            fprintf(stderr, "\n");                     \
            ++g_fail;                                  \
            return false;                              \
        }                                              \
    } while (0)

// returns false on a mismatch; *fell_back when the partition refuses the input
template <int NWAY, int NT, int VT>
bool run_union(const std::vector<std::vector<uint64_t>>& files, std::vector<uint64_t>* out, bool* fell_back, long long* max_tile) {
    using SH = Shape<NWAY, NT, VT>;
    static_assert(SH::TILE >= 32, "tile too small");
    const int nf = (int)files.size();
    *fell_back = false;
    out->clear();
    NwFiles F;
    long long total = 0;
    for (int f = 0; f < NW_MAX; ++f) {
        F.k[f] = f < nf ? files[f].data() : nullptr;
        F.n[f] = f < nf ? (long long)files[f].size() : 0;
        total += F.n[f];
    }
    F.nf = nf;
    if (total == 0) return true;
    const long long tile = SH::TILE, tol = SH::TILE / 32;
    const int num_tiles = (int)((total + tile - 1) / tile);
    // ---- three-level partition (nway_partition_kernel: strides 64, 8, 1) ----
    std::vector<NwBound> bounds(num_tiles + 1);
    {
        NwBound glo, ghi;
        nw_global_bracket(F, &glo, &ghi);
        bounds[num_tiles] = ghi;
        const int strides[3] = {64, 8, 1};
        int parent = 0;
        for (int L = 0; L < 3; ++L) {
            const int st = strides[L];
            for (long long t = 0; t < num_tiles; t += st) {
                if (parent == 0) {
                    if (t == 0) bounds[t] = glo;
                    else nw_refine(F, t * tile, tol, glo, ghi, &bounds[t]);
                } else {
                    if (t % parent == 0) continue;
                    const long long pl = t / parent * parent;
                    const long long ph = pl + parent < num_tiles ? pl + parent : num_tiles;
                    nw_refine(F, t * tile, tol, bounds[pl], bounds[ph], &bounds[t]);
                }
            }
            parent = st;
        }
    }
    std::vector<long long> part((size_t)(num_tiles + 1) * NW_MAX);
    for (int t = 0; t <= num_tiles; ++t)
        for (int f = 0; f < NW_MAX; ++f) part[(size_t)t * NW_MAX + f] = bounds[t].pos[f];
    // ---- nway_check_kernel ----
    *max_tile = 0;
    for (int t = 0; t < num_tiles; ++t) {
        long long sum = 0;
        bool bad = false;
        for (int f = 0; f < NW_MAX; ++f) {
            const long long d = part[(size_t)(t + 1) * NW_MAX + f] - part[(size_t)t * NW_MAX + f];
            if (d < 0) bad = true;
            sum += d;
        }
        if (sum > SH::CAP) bad = true;
        if (sum > *max_tile) *max_tile = sum;
        if (bad) {
            *fell_back = true;
            return true;
        }
    }
    // first row must be all zeros (nothing below the first cut)
    for (int f = 0; f < NW_MAX; ++f) CHECK(part[f] == 0, "first boundary of file %d is %lld", f, part[f]);

    // ---- tiles ----
    std::vector<uint64_t> slot(SH::SLOT_E, 0xDEADBEEFDEADBEEFull), X(SH::X_E, 0xDEADBEEFDEADBEEFull);
    for (int t = 0; t < num_tiles; ++t) {
        // poison: anything the merge uses must have been written for THIS tile
        std::fill(slot.begin(), slot.end(), 0xDEADBEEFDEADBEEFull);
        std::fill(X.begin(), X.end(), 0xDEADBEEFDEADBEEFull);
        NwGeom<NWAY> g;
        // loader: lane f = file f
        int base = 0;
        for (int f = 0; f < NWAY; ++f) {
            const long long lo = part[(size_t)t * NW_MAX + f], hi = part[(size_t)(t + 1) * NW_MAX + f];
            const int n = (int)(hi - lo);
            const int h = n > 0 ? (int)(((uintptr_t)(F.k[f] + lo) & 15u) >> 3) : 0;
            const int padded = (h + n + 1) & ~1;
            CHECK((base & 1) == 0, "odd base");
            for (int i = 0; i < n; ++i) slot[base + h + i] = F.k[f][lo + i];
            g.n[f] = n;
            g.off[f] = base + h;
            base += padded;
        }
        CHECK(base <= SH::SLOT_E - 8, "segments overflow the slot: %d", base);
        nw_build_tables<NWAY, VT>(&g);
        CHECK(g.tot <= SH::CAP, "tile too large");
        // inner levels
        const uint64_t* src = slot.data();
        uint64_t* dst = X.data();
        for (int l = 1; l < SH::LEVELS; ++l) {
            const int npairs = NWAY >> l;
            const int p0 = nw_pair0<NWAY>(l), t0 = nw_tb0<NWAY>(l);
            CHECK(g.tb[t0 + npairs] <= NT, "level %d needs %d threads", l, g.tb[t0 + npairs]);
            for (int tid = 0; tid < NT; ++tid) {
                int j = 0, m;
                if (npairs == 4) m = nw_find_pair<4>(g.tb + t0, tid, &j);
                else m = nw_find_pair<2>(g.tb + t0, tid, &j);
                if (m < 0) continue;
                const NwPair pr = g.pair[p0 + m];
                const int diag = j * VT;
                int steps = pr.lenA + pr.lenB - diag;
                CHECK(steps > 0, "thread without work inside a pair");
                if (steps > VT) steps = VT;
                const uint64_t* A = src + pr.srcA;
                const uint64_t* B = src + pr.srcB;
                const int a = nw_merge_path_g(A, pr.lenA, B, pr.lenB, diag);
                nw_walk_plain<VT>(A, pr.lenA, B, pr.lenB, a, diag - a, steps, dst + pr.dst + diag);
            }
            const uint64_t* tmp = src;
            src = dst;
            dst = const_cast<uint64_t*>(tmp);
        }
        // last level
        const NwPair pr = g.pair[NWAY - 2];
        const int tot = pr.lenA + pr.lenB;
        CHECK(tot == g.tot, "totals differ");
        std::vector<uint64_t> staged;
        for (int tid = 0; tid < NT; ++tid) {
            int diag = tid * VT;
            int steps = tot - diag;
            if (steps > VT) steps = VT;
            if (diag > tot) diag = tot;
            const uint64_t* A = src + pr.srcA;
            const uint64_t* B = src + pr.srcB;
            const int a = nw_merge_path_g(A, pr.lenA, B, pr.lenB, diag);
            uint64_t outk[VT];
            const unsigned mask = nw_walk_unique<VT>(A, pr.lenA, B, pr.lenB, a, diag - a, steps, outk);
            for (int it = 0; it < VT; ++it)
                if (mask & (1u << it)) staged.push_back(outk[it]);
        }
        for (uint64_t v : staged) CHECK(v != 0xDEADBEEFDEADBEEFull || true, "poison");
        out->insert(out->end(), staged.begin(), staged.end());
    }
    return true;
}

static std::vector<uint64_t> expected_union(const std::vector<std::vector<uint64_t>>& files) {
    std::vector<uint64_t> all;
    for (auto& f : files) all.insert(all.end(), f.begin(), f.end());
    std::sort(all.begin(), all.end());
    all.erase(std::unique(all.begin(), all.end()), all.end());
    return all;
}

template <int NWAY, int NT, int VT>
bool check_case(const char* name, const std::vector<std::vector<uint64_t>>& files, bool expect_tiled = true) {
    std::vector<uint64_t> got;
    bool fb = false;
    long long max_tile = 0;
    if (!run_union<NWAY, NT, VT>(files, &got, &fb, &max_tile)) return false;
    if (fb) {
        CHECK(!expect_tiled, "%s <%d,%d,%d>: partition refused a duplicate-free input (max tile %lld)", name, NWAY, NT, VT, max_tile);
        return true;
    }
    const std::vector<uint64_t> exp = expected_union(files);
    CHECK(got.size() == exp.size(), "%s <%d,%d,%d>: %zu keys, expected %zu", name, NWAY, NT, VT, got.size(), exp.size());
    for (size_t i = 0; i < exp.size(); ++i) CHECK(got[i] == exp[i], "%s <%d,%d,%d>: key %zu differs", name, NWAY, NT, VT, i);
    return true;
}

static std::mt19937_64 rng(12345);

static std::vector<uint64_t> sorted_unique(std::vector<uint64_t> v) {
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
    return v;
}
static std::vector<uint64_t> uniform_file(size_t n, uint64_t lo, uint64_t hi) {
    std::vector<uint64_t> v(n);
    std::uniform_int_distribution<uint64_t> d(lo, hi);
    for (auto& x : v) x = d(rng);
    return sorted_unique(v);
}
// unaligned start: drop the first element of a copy so data() + 1 style misalignment shows up
static std::vector<std::vector<uint64_t>> subset_files(const std::vector<uint64_t>& universe, int nf, double p) {
    std::vector<std::vector<uint64_t>> files(nf);
    std::uniform_real_distribution<double> d(0, 1);
    for (uint64_t x : universe)
        for (int f = 0; f < nf; ++f)
            if (d(rng) < p) files[f].push_back(x);
    return files;
}

template <int NWAY, int NT, int VT>
void battery(int nf) {
    // 1. C3-like: subsets of one universe
    check_case<NWAY, NT, VT>("subsets", subset_files(uniform_file(60000, 0, (1ull << 62) - 1), nf, 0.5));
    // 2. full 64-bit range incl. extremes
    {
        auto files = subset_files(uniform_file(20000, 0, ~0ull), nf, 0.7);
        files[0].insert(files[0].begin(), 0);
        files[0] = sorted_unique(files[0]);
        files[nf - 1].push_back(~0ull);
        files[nf - 1] = sorted_unique(files[nf - 1]);
        if (nf > 1) { files[1].push_back(~0ull); files[1] = sorted_unique(files[1]); }
        check_case<NWAY, NT, VT>("extremes", files);
    }
    // 3. identical files
    {
        auto u = uniform_file(15000, 0, 1ull << 40);
        std::vector<std::vector<uint64_t>> files(nf, u);
        check_case<NWAY, NT, VT>("identical", files);
    }
    // 4. disjoint key ranges, very different sizes, some empty
    {
        std::vector<std::vector<uint64_t>> files(nf);
        for (int f = 0; f < nf; ++f) {
            const size_t n = (f % 3 == 2) ? 0 : (size_t)(100 << (f % 7));
            files[f] = uniform_file(n, (uint64_t)f << 50, ((uint64_t)f << 50) + (1ull << 30));
        }
        check_case<NWAY, NT, VT>("disjoint", files);
    }
    // 5. clustered: dense runs of consecutive integers + sparse background
    {
        std::vector<std::vector<uint64_t>> files(nf);
        for (int f = 0; f < nf; ++f) {
            std::vector<uint64_t> v = uniform_file(3000, 0, ~0ull >> 1);
            const uint64_t c = 1000000007ull * (f % 3 + 1);
            for (uint64_t i = 0; i < 5000; ++i) v.push_back(c + i * (f % 2 + 1));
            files[f] = sorted_unique(v);
        }
        check_case<NWAY, NT, VT>("clustered", files);
    }
    // 6. tiny inputs
    {
        std::vector<std::vector<uint64_t>> files(nf);
        files[0] = {5};
        if (nf > 1) files[1] = {5, 7};
        if (nf > 2) files[nf - 1] = {1, 5, 9};
        check_case<NWAY, NT, VT>("tiny", files);
        std::vector<std::vector<uint64_t>> one(nf);
        one[nf - 1] = {42};
        check_case<NWAY, NT, VT>("single", one);
    }
    // 7. one big file + small ones; odd sizes for alignment variety
    {
        std::vector<std::vector<uint64_t>> files(nf);
        files[0] = uniform_file(50001, 0, 1ull << 62);
        for (int f = 1; f < nf; ++f) files[f] = uniform_file(97 + 13 * f, 0, 1ull << 62);
        check_case<NWAY, NT, VT>("skewed", files);
    }
    // 8. geometric key distribution (density varies over 40 binades)
    {
        std::vector<std::vector<uint64_t>> files(nf);
        std::uniform_real_distribution<double> d(0, 40);
        for (int f = 0; f < nf; ++f) {
            std::vector<uint64_t> v(8000);
            for (auto& x : v) x = (uint64_t)exp2(d(rng) + 20.0);
            files[f] = sorted_unique(v);
        }
        check_case<NWAY, NT, VT>("geometric", files);
    }
}

int main() {
    // small shapes: many tiles, every thread-to-pair corner; then the shipped shapes
    for (int nf = 5; nf <= 8; ++nf) battery<8, 64, 5>(nf);
    for (int nf = 3; nf <= 4; ++nf) battery<4, 64, 5>(nf);
    battery<2, 64, 5>(2);
    battery<8, 256, 9>(8);
    battery<8, 128, 17>(7);
    battery<8, 256, 13>(8);
    battery<8, 512, 13>(8);
    battery<4, 512, 13>(4);
    battery<4, 256, 9>(4);
    battery<4, 128, 17>(3);
    battery<2, 256, 9>(2);
    battery<2, 128, 17>(2);
    // not duplicate-free: one key repeated far beyond a tile must be refused, never mis-merged
    {
        std::vector<std::vector<uint64_t>> files(8);
        for (int f = 0; f < 8; ++f) {
            files[f] = uniform_file(2000, 0, 1ull << 40);
            files[f].insert(files[f].end(), 3000, 1ull << 41);
        }
        check_case<8, 256, 9>("duplicates", files, false);
    }
    if (g_fail) {
        fprintf(stderr, "%d failures\n", g_fail);
        return 1;
    }
    printf("nway model ok\n");
    return 0;
}
