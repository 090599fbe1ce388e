// rows_model.cpp -- host model of the row-based N-way union tile kernel (unikmer_b200/csrc/nunion.cu).
//
// Test infrastructure (CPU suite, no GPU): compiles the SAME arithmetic the kernel uses (unikmer_b200/csrc/rows_core.cuh:
// per-tile pair tables with reversed / congruent runs, merge-path split in padded rows, rotated gather, bitonic merge
// network, row write with pad slot; unikmer_b200/csrc/nway_core.cuh: the partition) with g++ and replays the kernel's
// per-thread schedule sequentially -- 8 warps x 31 rows per level, lane 31 computing only the end split -- then compares
// with std::set_union.  It also checks the claims the design rests on: every gather round of a half-warp touches 16
// different banks on the levels that read rows (level >= 2), and no level overflows its buffer.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <random>
#include <vector>

#include "../../unikmer_b200/csrc/rows_core.cuh"

static int g_fail = 0;
#define CHECK(c, ...)                                            \
    do {                                                         \
        if (!(c)) {                                              \
            fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); \
            fprintf(stderr, __VA_ARGS__);                        \
            fprintf(stderr, "\n");                               \
            ++g_fail;                                            \
            return false;                                        \
        }                                                        \
    } while (0)

constexpr int NWARPS = 8;
constexpr uint64_t POISON = 0xDEADBEEFDEADBEEFull;

template <int NWAY>
struct Shape {
    static constexpr int ROWS = NWARPS * RW_ROWS_PER_WARP;          // rows a level may have
    static constexpr int CAP_OUT = ROWS * RW_E;                      // keys the last level may see
    // input keys of a tile: the last level sees them after log2(NWAY) - 1 rounds of padding (x 17/16 each) and every
    // pair may end in a partial row
    static constexpr int LEVELS = RwGeom<NWAY>::LEVELS;
    static constexpr int CAP = NWAY == 8 ? 3456 : NWAY == 4 ? 3680 : 3936;
    static constexpr int TILE = (CAP * 16 / 17) & ~31;  // boundaries are exact to +-TILE/32
    static constexpr int BUF_E = CAP_OUT + 64;
};

static long long g_conflict_rounds = 0, g_rounds = 0;

template <int NWAY>
bool run_union(const std::vector<std::vector<uint64_t>>& files, std::vector<uint64_t>* out, bool* fell_back) {
    using SH = Shape<NWAY>;
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
    for (int t = 0; t < num_tiles; ++t) {
        long long sum = 0;
        for (int f = 0; f < NW_MAX; ++f) sum += bounds[t + 1].pos[f] - bounds[t].pos[f];
        if (sum > SH::CAP) {
            *fell_back = true;
            return true;
        }
    }
    std::vector<uint64_t> slot(SH::BUF_E), X(SH::BUF_E);
    for (int t = 0; t < num_tiles; ++t) {
        std::fill(slot.begin(), slot.end(), POISON);
        std::fill(X.begin(), X.end(), POISON);
        RwGeom<NWAY> g;
        int base = 0;
        for (int f = 0; f < NWAY; ++f) {
            const long long lo = bounds[t].pos[f], hi = bounds[t + 1].pos[f];
            const int n = (int)(hi - lo);
            const int h = n > 0 ? (int)(((uintptr_t)(F.k[f] + lo) & 15u) >> 3) : 0;
            const int padded = (h + n + 1) & ~1;
            for (int i = 0; i < n; ++i) slot[base + h + i] = F.k[f][lo + i];
            g.n[f] = n;
            g.off[f] = base + h;
            base += padded;
        }
        CHECK(base <= SH::BUF_E, "segments overflow the slot");
        const int extent = rw_build_tables<NWAY>(&g);
        CHECK(extent <= SH::BUF_E, "a level writes %d elements, buffers hold %d", extent, SH::BUF_E);
        const uint64_t* src = slot.data();
        uint64_t* dst = X.data();
        std::vector<uint64_t> staged;
        for (int l = 1; l <= SH::LEVELS; ++l) {
            const bool last = l == SH::LEVELS;
            const int npairs = NWAY >> l;
            const RwPair* prs = g.pair + rw_pair0<NWAY>(l);
            const int rows = g.rows[l - 1];
            CHECK(rows <= SH::ROWS, "level %d has %d rows, the CTA handles %d", l, rows, SH::ROWS);
            for (int w = 0; w < NWARPS; ++w) {
                // every lane: the start split of its row (lane 31: of the row after the warp's last)
                int m_[32], j_[32], a_[32];
                for (int lane = 0; lane < 32; ++lane) {
                    const int rho = w * RW_ROWS_PER_WARP + lane;
                    int j = 0;
                    const int m = npairs == 4 ? rw_find_pair<4>(prs, rows, rho, &j) : npairs == 2 ? rw_find_pair<2>(prs, rows, rho, &j)
                                                                                                  : rw_find_pair<1>(prs, rows, rho, &j);
                    m_[lane] = m;
                    j_[lane] = j;
                    a_[lane] = 0;
                    if (m >= 0) {
                        const RwPair& pr = prs[m];
                        a_[lane] = rw_merge_path(src + pr.a_off, pr.a_len, src + pr.b_off, pr.b_dir, pr.b_len, j * RW_E);
                    }
                }
                int bank_use[2][16][16];
                memset(bank_use, 0, sizeof bank_use);
                for (int lane = 0; lane < RW_ROWS_PER_WARP; ++lane) {
                    const int m = m_[lane];
                    if (m < 0) continue;
                    const RwPair& pr = prs[m];
                    const int j = j_[lane], a = a_[lane], b = j * RW_E - a;
                    const int nk = pr.a_len + pr.b_len;
                    int cnt = nk - j * RW_E;
                    CHECK(cnt > 0, "row without keys");
                    if (cnt > RW_E) cnt = RW_E;
                    const int a2 = (m_[lane + 1] == m) ? a_[lane + 1] : pr.a_len;
                    const int na = a2 - a, nb = cnt - na;
                    CHECK(na >= 0 && nb >= 0 && b + nb <= pr.b_len && a + na <= pr.a_len, "bad split");
                    const int rot = (lane - (pr.a_off + a)) & 15;
                    // bank check (levels that read rows): round r of this lane reads word (pos mod 16)
                    for (int r = 0; r < 16; ++r) {
                        const int e = (r + rot) & 15;
                        int pos = -1;
                        if (e < na) pos = pr.a_off + a + e;
                        else if (e >= 16 - nb) pos = pr.b_off + pr.b_dir * (b + 15 - e);
                        if (pos >= 0) {
                            CHECK(pos < SH::BUF_E, "gather out of the buffer");
                            if (l > 1) bank_use[lane >> 4][r][pos & 15]++;
                        }
                    }
                    uint64_t s[16];
                    rw_gather_sort(src + pr.a_off, src + pr.b_off, pr.b_dir, a, na, b, nb, rot, s);
                    for (int i = 0; i < cnt; ++i) CHECK(s[i] != POISON || false, "poison in a row");
                    for (int i = 1; i < cnt; ++i) CHECK(s[i - 1] <= s[i], "row not sorted");
                    if (!last) {
                        rw_write_row(dst + pr.d_off, pr.d_dir, j, cnt, s);
                    } else {
                        // first key of every run of equal keys; the key before the row = max(A[a - 1], B[b - 1])
                        bool has_prev = a > 0 || b > 0;
                        uint64_t prev = 0;
                        if (a > 0) prev = src[pr.a_off + a - 1];
                        if (b > 0) {
                            const uint64_t pb = src[pr.b_off + pr.b_dir * (b - 1)];
                            if (pb > prev) prev = pb;
                        }
                        for (int i = 0; i < cnt; ++i) {
                            if (!has_prev || s[i] != prev) staged.push_back(s[i]);
                            prev = s[i];
                            has_prev = true;
                        }
                    }
                }
                for (int h = 0; h < 2; ++h)
                    for (int r = 0; r < 16; ++r) {
                        bool conflict = false;
                        for (int bk = 0; bk < 16; ++bk) conflict |= bank_use[h][r][bk] > 1;
                        g_rounds++;
                        g_conflict_rounds += conflict;
                    }
            }
            const uint64_t* tmp = src;
            src = dst;
            dst = const_cast<uint64_t*>(tmp);
        }
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

template <int NWAY>
bool check_case(const char* name, const std::vector<std::vector<uint64_t>>& files, bool expect_tiled = true) {
    std::vector<uint64_t> got;
    bool fb = false;
    if (!run_union<NWAY>(files, &got, &fb)) return false;
    if (fb) {
        CHECK(!expect_tiled, "%s <%d>: partition refused a duplicate-free input", name, NWAY);
        return true;
    }
    const std::vector<uint64_t> exp = expected_union(files);
    CHECK(got.size() == exp.size(), "%s <%d>: %zu keys, expected %zu", name, NWAY, got.size(), exp.size());
    for (size_t i = 0; i < exp.size(); ++i) CHECK(got[i] == exp[i], "%s <%d>: key %zu differs", name, NWAY, i);
    return true;
}

static std::mt19937_64 rng(4321);
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
static std::vector<std::vector<uint64_t>> subset_files(const std::vector<uint64_t>& universe, int nf, double p) {
    std::vector<std::vector<uint64_t>> files(nf);
    std::uniform_real_distribution<double> d(0, 1);
    for (uint64_t x : universe)
        for (int f = 0; f < nf; ++f)
            if (d(rng) < p) files[f].push_back(x);
    return files;
}

template <int NWAY>
void battery(int nf) {
    check_case<NWAY>("subsets", subset_files(uniform_file(60000, 0, (1ull << 62) - 1), nf, 0.5));
    check_case<NWAY>("sparse subsets", subset_files(uniform_file(60000, 0, (1ull << 62) - 1), nf, 0.05));
    {
        auto files = subset_files(uniform_file(20000, 0, ~0ull), nf, 0.7);
        files[0].insert(files[0].begin(), 0);
        files[0] = sorted_unique(files[0]);
        files[nf - 1].push_back(~0ull);
        files[nf - 1] = sorted_unique(files[nf - 1]);
        if (nf > 1) { files[1].push_back(~0ull); files[1] = sorted_unique(files[1]); }
        check_case<NWAY>("extremes", files);
    }
    {
        auto u = uniform_file(15000, 0, 1ull << 40);
        std::vector<std::vector<uint64_t>> files(nf, u);
        check_case<NWAY>("identical", files);
    }
    {
        std::vector<std::vector<uint64_t>> files(nf);
        for (int f = 0; f < nf; ++f) {
            const size_t n = (f % 3 == 2) ? 0 : (size_t)(100 << (f % 7));
            files[f] = uniform_file(n, (uint64_t)f << 50, ((uint64_t)f << 50) + (1ull << 30));
        }
        check_case<NWAY>("disjoint", files);
    }
    {
        std::vector<std::vector<uint64_t>> files(nf);
        for (int f = 0; f < nf; ++f) {
            std::vector<uint64_t> v = uniform_file(3000, 0, ~0ull >> 1);
            const uint64_t c = 1000000007ull * (f % 3 + 1);
            for (uint64_t i = 0; i < 20000; ++i) v.push_back(c + i * (f % 2 + 1));
            files[f] = sorted_unique(v);
        }
        check_case<NWAY>("clustered", files);
    }
    {
        std::vector<std::vector<uint64_t>> files(nf);
        files[0] = uniform_file(50000, 0, 1ull << 62);
        for (int f = 1; f < nf; ++f) files[f] = uniform_file(37 + 11 * f, 0, 1ull << 62);
        check_case<NWAY>("skewed", files);
    }
    {
        std::vector<std::vector<uint64_t>> files(nf);
        for (int f = 0; f < nf; ++f) files[f] = {5};
        files[0] = {1, 5, 9};
        check_case<NWAY>("tiny", files);
    }
    // every total from 1 to a few rows, to walk through all partial-row shapes
    for (int n = 1; n <= 70; ++n) {
        auto u = uniform_file((size_t)n, 0, 1ull << 30);
        check_case<NWAY>("small totals", subset_files(u, nf, 0.6));
    }
    // misaligned starts
    {
        auto files = subset_files(uniform_file(30000, 0, 1ull << 50), nf, 0.5);
        std::vector<std::vector<uint64_t>> shifted(nf);
        std::vector<std::vector<uint64_t>> keep(nf);
        for (int f = 0; f < nf; ++f) {
            keep[f].assign(files[f].begin() + (f % 2), files[f].end());
            shifted[f] = keep[f];
        }
        check_case<NWAY>("odd starts", shifted);
    }
}

int main() {
    battery<8>(8);
    battery<8>(5);
    battery<8>(7);
    battery<4>(4);
    battery<4>(3);
    battery<2>(2);
    if (g_fail) {
        fprintf(stderr, "%d failures\n", g_fail);
        return 1;
    }
    printf("rows model ok: %lld gather rounds on row levels, %lld with a bank conflict\n", g_rounds, g_conflict_rounds);
    if (g_conflict_rounds != 0) {
        fprintf(stderr, "the gather of a row level is not conflict-free\n");
        return 1;
    }
    return 0;
}
