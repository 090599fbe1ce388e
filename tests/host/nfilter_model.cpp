// nfilter_model.cpp -- host model of the per-lane searches of the single-pass inter / diff filter
// (unikmer_b200/csrc/nfilter_core.cuh, used by nfilter.cu).
//
// Test infrastructure (CPU suite, no GPU).  The kernel answers "is this key of file 0 in the tile's segment of file f"
// (inter.go:228-257 / diff.go:395-431, per key) with searches whose probe sequence is fixed per warp: steps
// 2^(lg-1) .. 1 from the first element, every probe address clamped to the last element, lg the same for all lanes
// and only >= what a lane's own length needs.  This model runs the SAME functions on a byte array standing in for
// shared memory and checks, for every length 0 .. 2200, every admissible lg and keys on and between the elements:
// nf_find / nf_find2 against std::binary_search, nf_rank against std::lower_bound, nf_lg against its definition, and that
// no probe ever reads outside [first element, last element] of the segment it was given.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <random>
#include <vector>

static uint32_t g_lo = 0, g_hi = 0;
static long long g_oob = 0, g_probes = 0;
#define NF_HOST_CHECK(a)                        \
    do {                                        \
        ++g_probes;                             \
        if ((a) < g_lo || (a) > g_hi || ((a) & 7u)) ++g_oob; \
    } while (0)
#include "../../unikmer_b200/csrc/nfilter_core.cuh"

const unsigned char* nf_host_smem = nullptr;

static int g_fail = 0;
#define CHECK(c, ...)                                            \
    do {                                                         \
        if (!(c)) {                                              \
            if (g_fail < 20) {                                   \
                fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); \
                fprintf(stderr, __VA_ARGS__);                    \
                fprintf(stderr, "\n");                           \
            }                                                    \
            ++g_fail;                                            \
        }                                                        \
    } while (0)

int main() {
    std::mt19937_64 rng(20261017);
    for (int n = 0; n < 5000; ++n) {
        int lg = 0;
        while ((1 << lg) < (n > 1 ? n : 1)) ++lg;
        CHECK(nf_lg(n) == lg, "nf_lg(%d) = %d, want %d", n, nf_lg(n), lg);
    }
    std::vector<uint64_t> smem(4096 + 64);
    nf_host_smem = reinterpret_cast<const unsigned char*>(smem.data());
    long long cases = 0;
    std::vector<int> lengths;
    for (int n = 0; n <= 300; ++n) lengths.push_back(n);
    for (int n : {511, 512, 513, 1000, 1023, 1024, 1025, 2047, 2048, 2049, 2200}) lengths.push_back(n);
    for (int n : lengths) {
        for (int rep = 0; rep < (n <= 300 ? 3 : 1); ++rep) {
            const int base = 8 + (int)(rng() % 16);  // the segment starts somewhere inside the slot
            // sorted, duplicate-free, with gaps of 1 (neighbouring keys) and wider ones; extremes included now and then
            std::vector<uint64_t> seg(n);
            uint64_t key = rep == 1 ? 0 : rng() % 1000;
            for (int i = 0; i < n; ++i) {
                seg[i] = key;
                key += 1 + (rng() % 3 == 0 ? 0 : rng() % 1000);
            }
            if (rep == 2 && n > 0) seg[n - 1] = ~0ull;
            for (auto& v : smem) v = rng();  // whatever surrounds the segment must not matter
            for (int i = 0; i < n; ++i) smem[base + i] = seg[i];
            const uint32_t seg_a = (uint32_t)base * 8u;
            g_lo = seg_a;
            g_hi = seg_a + (uint32_t)(n > 0 ? n - 1 : 0) * 8u;
            std::vector<uint64_t> probes = {0ull, ~0ull, rng()};
            for (int i = 0; i < n; i += (n > 300 ? 7 : 1)) {
                probes.push_back(seg[i]);
                probes.push_back(seg[i] + 1);
                if (seg[i] > 0) probes.push_back(seg[i] - 1);
            }
            const int lg0 = nf_lg(n);
            for (int lg = lg0; lg <= std::max(lg0 + 2, 12); ++lg) {
                for (size_t q = 0; q < probes.size(); ++q) {
                    const uint64_t x = probes[q];
                    const bool want = std::binary_search(seg.begin(), seg.end(), x);
                    const int want_rank = (int)(std::lower_bound(seg.begin(), seg.end(), x) - seg.begin());
                    CHECK(nf_find(seg_a, n, x, lg) == want, "nf_find n=%d lg=%d x=%llu", n, lg, (unsigned long long)x);
                    CHECK(nf_rank(seg_a, n, x, lg) == want_rank, "nf_rank n=%d lg=%d x=%llu: %d want %d", n, lg, (unsigned long long)x,
                          nf_rank(seg_a, n, x, lg), want_rank);
                    const uint64_t y = probes[(q * 7 + 3) % probes.size()];
                    bool f0, f1;
                    nf_find2(seg_a, n, x, y, lg, &f0, &f1);
                    CHECK(f0 == want && f1 == std::binary_search(seg.begin(), seg.end(), y), "nf_find2 n=%d lg=%d", n, lg);
                    ++cases;
                }
            }
        }
    }
    CHECK(g_oob == 0, "%lld of %lld probes outside their segment", g_oob, g_probes);
    if (g_fail) {
        fprintf(stderr, "%d check(s) failed\n", g_fail);
        return 1;
    }
    printf("nfilter model ok: %lld searches, %lld probes, none outside its segment\n", cases, g_probes);
    return 0;
}
