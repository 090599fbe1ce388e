// nunion.cu -- single-pass N-way (N <= 8) union of sorted duplicate-free k-mer streams, row-based merge levels.
//
// Replaces the hash-set union of union.go:186-208 and its key sort (union.go:260-305) with ONE pass over the inputs: the
// key space is cut into tiles of ~TILE elements summed over all files (multi-sequence selection, nway.cu), every tile
// is brought into shared memory with one 1-D TMA bulk copy per file and merged there in log2(N) levels of two-way merges.
// HBM traffic is the algorithmic minimum (every input byte read once, every output byte written once).
//
// What is new against nway.cu (whose levels are per-thread sequential merge walks: one dependent, bank-conflicted
// shared-memory load per key and level) is the level itself -- rows_core.cuh: a thread owns a ROW of 16 merged keys; it
// finds the row's merge-path split, gathers the row's inputs (a run of A ascending + a run of B descending) with 16
// independent, bank-conflict-free loads (rows are stored with one pad slot per 16 keys, the B run of every pair
// reversed at a congruent address, each lane reads rotated by its lane index), sorts them with a bitonic merge network
// in registers and writes the row back conflict-free.  The last level keeps the row in registers, flags the first key of
// every run of equal keys (the same k-mer in several files, and the pad slots, which repeat a key), and the distinct
// keys leave through a block scan, in-place staging and a deferred, coalesced copy-out.
//
// Kernel shape (persistent, warp-specialised, launched cooperatively; tiles round-robin):
//   warp 0 loader : lane f owns file f -- tile geometry, unaligned head / tail by plain loads, body by TMA
//                   (cp.async.bulk -> mbarrier complete_tx); lane 0 lays out the pair tables of the tile.
//   warp 1 prefix : output offsets of the grid iteration (gathers the G counts of tiles [i*G, (i+1)*G)).
//   warps 2..9    : 31 rows per warp and level (lane 31 only computes the end split of the warp's last row, so no split
//                   ever crosses a warp through shared memory): level 1 slot -> X, level 2 X -> slot, last level ->
//                   registers; the copy-out of the tile before runs between level 1 and level 2.
#include <stdlib.h>

#include "common.cuh"
#include "rows_core.cuh"

int ukm_nway_partition(ukm_ctx* ctx, const NwFiles& F, long long total, int tile, int cap, ukm_tmp& tmp, NwBound** d_bounds_out,
                       uint64_t** d_status_out, int* num_tiles_out, bool* bad);

namespace {

constexpr int NU_WARPS = 8;
constexpr int NU_NT = NU_WARPS * 32;  // consumer threads
constexpr int NU_AUX = 64;            // loader + prefix warps
constexpr int NU_MAX_GRID = 512;      // prefix warp keeps NU_MAX_GRID / 32 counts per lane
constexpr int NU_SLOTS = 2;

template <int NWAY>
struct NuShape {
    static constexpr int ROWS = NU_WARPS * RW_ROWS_PER_WARP;  // rows a level may have
    // input keys a tile may hold: the last level sees them after LEVELS - 1 rounds of padding (x 17/16 each), and every
    // pair of every level may end in a partial row (tests/host/rows_model.cpp walks these bounds)
    static constexpr int CAP = NWAY == 8 ? 3456 : NWAY == 4 ? 3680 : 3936;
    static constexpr int TILE = (CAP * 16 / 17) & ~31;        // nominal tile; boundaries are exact to +-TILE/32
    static constexpr int BUF_E = ROWS * RW_E + 64;             // elements per buffer (slot / X): 32256 bytes, a multiple of 128
    static constexpr int LEVELS = RwGeom<NWAY>::LEVELS;
};

struct NuArgs {
    NwFiles F;
    const NwBound* bounds;  // bounds[t].pos[f] = first element of tile t in file f
    uint64_t* outK;
    uint64_t* status;  // one count word per tile (flag << 62 | count)
    unsigned long long* total_out;
    int num_tiles;
    int copy_pos;  // where the copy-out of the tile before runs: 1 = after level 1, 2 = after level 2, 3 = after the scan of this tile
    int* err;
};

__device__ __forceinline__ uint64_t nu_lds(uint32_t a) {
    uint64_t v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void nu_sts(uint32_t a, uint64_t v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }

__device__ __forceinline__ int nu_slice_h(const uint64_t* g, long long start) {
    return (int)((reinterpret_cast<uintptr_t>(g + start) & 15u) >> 3);
}

template <int IMM>
__device__ __forceinline__ void nu_sts_imm(uint32_t a, uint64_t v) {  // st.shared [a + IMM]: no address arithmetic per store
    asm volatile("st.shared.u64 [%0+%2], %1;" ::"r"(a), "l"(v), "n"(IMM) : "memory");
}

// rw_merge_path on shared addresses: A = byte address of A[0], B of B[0], bdir8 = +-8.  The split of a diagonal is almost
// always close to diag * na / (na + nb) (the runs interleave evenly): gallop out from that guess until the answer is
// bracketed (steps 4, 8, ..), then bisect the bracket -- 6-8 probes instead of log2(n) + 1, each two shared-memory loads.
__device__ __forceinline__ int nu_merge_path(uint32_t A, int na, uint32_t B, int bdir8, int nb, int diag) {
    int lo = diag > nb ? diag - nb : 0;
    int hi = diag < na ? diag : na;
    if (lo >= hi) return lo;
    // P(x) := A[x] > B[diag - 1 - x] is monotone false..true on [lo, hi); the answer is the first true (or hi)
    auto pred = [&](int x) { return nu_lds(A + (unsigned)x * 8u) > nu_lds(B + (unsigned)(bdir8 * (diag - 1 - x))); };
    int g = (int)(((long long)diag * na) / (na + nb));
    g = g < lo ? lo : (g > hi - 1 ? hi - 1 : g);
    int step = 4;
    if (pred(g)) {
        hi = g;
        while (hi > lo) {
            const int x = hi - step < lo ? lo : hi - step;
            if (pred(x)) { hi = x; step <<= 1; }
            else { lo = x + 1; break; }
        }
    } else {
        lo = g + 1;
        while (lo < hi) {
            const int x = lo + step - 1 > hi - 1 ? hi - 1 : lo + step - 1;
            if (!pred(x)) { lo = x + 1; step <<= 1; }
            else { hi = x; break; }
        }
    }
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (pred(mid)) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}

// one level of one tile, one warp: 31 rows.  LAST: the row stays in s[], *emit = bit i set if s[i] is a new distinct key.
template <int NWAY, int L>
__device__ __forceinline__ void nu_level(const RwGeom<NWAY>& g, uint32_t src, uint32_t dst, int w, unsigned lane, uint64_t* s,
                                         unsigned* emit) {
    constexpr int LEVELS = RwGeom<NWAY>::LEVELS;
    constexpr bool LAST = L == LEVELS;
    constexpr bool BFWD = L == 1;  // level 1 reads the files' segments as TMA delivered them: B ascending in memory too
    constexpr int NP = NWAY >> L;
    const RwPair* prs = g.pair + rw_pair0<NWAY>(L);
    const int rows = g.rows[L - 1];
    const int rho = w * RW_ROWS_PER_WARP + (int)lane;
    int j = 0;
    const int m = rw_find_pair<NP>(prs, rows, rho, &j);
    int a = 0;
    RwPair pr;
    pr.a_off = pr.a_len = pr.b_off = pr.b_len = pr.d_off = pr.row0 = 0;
    pr.b_dir = pr.d_dir = 1;
    uint32_t A = src, B = src;
    constexpr int bdir8 = BFWD ? 8 : -8;
    if (m >= 0) {
        pr = prs[m];
        A = src + (unsigned)pr.a_off * 8u;
        B = src + (unsigned)pr.b_off * 8u;
        a = nu_merge_path(A, pr.a_len, B, bdir8, pr.b_len, j * RW_E);
    }
    const int m_next = __shfl_down_sync(0xffffffffu, m, 1);
    const int a_next = __shfl_down_sync(0xffffffffu, a, 1);
    *emit = 0;
    if (lane < RW_ROWS_PER_WARP && m >= 0) {
        const int nk = pr.a_len + pr.b_len;
        int cnt = nk - j * RW_E;
        cnt = cnt > RW_E ? RW_E : cnt;
        const int a2 = (m_next == m) ? a_next : pr.a_len;
        const int na = a2 - a, nb = cnt - na, b = j * RW_E - a;
        const unsigned rot8 = (unsigned)(((int)lane - (pr.a_off + a)) & 15) * 8u;
        // gather: element e (circular order) of the row -- A ascending, +inf fill, B descending (e = 15 is B[b]); lane l
        // starts at e = rot, so that on the levels that read rows every round of a half-warp touches 16 different banks
        const uint32_t pa = A + (unsigned)a * 8u;
        const uint32_t pb = B + (unsigned)(bdir8 * (b + 15));  // element e of the B part: pb - bdir8 * e
        const unsigned na8 = (unsigned)na * 8u;
        if (cnt == RW_E) {
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const unsigned e8 = (rot8 + 8u * r) & 120u;
                const bool inA = e8 < na8;
                const uint32_t addr = BFWD ? (inA ? pa + e8 : pb - e8) : ((inA ? pa : pb) + e8);
                s[r] = nu_lds(addr);
            }
        } else {
            const unsigned bf8 = (unsigned)(16 - nb) * 8u;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const unsigned e8 = (rot8 + 8u * r) & 120u;
                const bool inA = e8 < na8;
                const uint32_t addr = BFWD ? (inA ? pa + e8 : pb - e8) : ((inA ? pa : pb) + e8);
                uint64_t v = nu_lds(addr);  // always inside the buffers (rows_model.cpp), meaningless between the two runs
                if (!inA && e8 < bf8) v = ~0ull;
                s[r] = v;
            }
        }
        rw_bitonic16(s);
        if (!LAST) {
            const uint32_t prow = dst + (unsigned)pr.d_off * 8u + (unsigned)(pr.d_dir * 8 * (RW_E + 1) * j);
            if (cnt == RW_E) {
                if (pr.d_dir > 0) {
#define NU_ST(i) nu_sts_imm<8 * (i)>(prow, s[i]);
                    NU_ST(0) NU_ST(1) NU_ST(2) NU_ST(3) NU_ST(4) NU_ST(5) NU_ST(6) NU_ST(7)
                    NU_ST(8) NU_ST(9) NU_ST(10) NU_ST(11) NU_ST(12) NU_ST(13) NU_ST(14) NU_ST(15)
#undef NU_ST
                    nu_sts_imm<8 * 16>(prow, s[15]);  // the pad slot repeats the last key
                } else {
#define NU_ST(i) nu_sts_imm<-8 * (i)>(prow, s[i]);
                    NU_ST(0) NU_ST(1) NU_ST(2) NU_ST(3) NU_ST(4) NU_ST(5) NU_ST(6) NU_ST(7)
                    NU_ST(8) NU_ST(9) NU_ST(10) NU_ST(11) NU_ST(12) NU_ST(13) NU_ST(14) NU_ST(15)
#undef NU_ST
                    nu_sts_imm<-8 * 16>(prow, s[15]);
                }
            } else {
                const int ddir8 = pr.d_dir * 8;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (i < cnt) nu_sts(prow + (unsigned)(ddir8 * i), s[i]);
            }
        } else {
            // first key of every run of equal keys; the key before the row is max(A[a - 1], B[b - 1])
            bool has_prev = a > 0 || b > 0;
            uint64_t prev = 0;
            if (a > 0) prev = nu_lds(A + (unsigned)(a - 1) * 8u);
            if (b > 0) {
                const uint64_t vb = nu_lds(B + (unsigned)(bdir8 * (b - 1)));
                prev = vb > prev ? vb : prev;
            }
            unsigned mask = (!has_prev || s[0] != prev) ? 1u : 0u;
#pragma unroll
            for (int i = 1; i < 16; ++i) mask |= (s[i] != s[i - 1] ? 1u : 0u) << i;
            *emit = cnt == RW_E ? mask : (mask & ((1u << cnt) - 1u));
        }
    }
}

template <int NWAY>
__global__ void __launch_bounds__(NU_NT + NU_AUX, 2) nunion_kernel(const NuArgs p) {
    using SH = NuShape<NWAY>;
    constexpr int LEVELS = SH::LEVELS;
    constexpr int SLOTS = NU_SLOTS;
    extern __shared__ __align__(128) unsigned char nu_smem[];
    uint64_t* s_slots = reinterpret_cast<uint64_t*>(nu_smem);  // SLOTS * BUF_E
    uint64_t* s_x = s_slots + (size_t)SLOTS * SH::BUF_E;       // BUF_E
    __shared__ __align__(8) uint64_t full_bar[SLOTS], empty_bar[SLOTS], pre_bar[SLOTS];
    __shared__ unsigned long long s_cnt[SLOTS], s_pre[SLOTS];
    __shared__ RwGeom<NWAY> s_geom[SLOTS];
    __shared__ const uint64_t* s_fk[NW_MAX];
    __shared__ unsigned s_scan[NU_WARPS + 2];

    const int G = gridDim.x;
    const int n_my = (p.num_tiles - (int)blockIdx.x + G - 1) / G;  // tiles of this CTA (>= 1: the grid never exceeds the tiles)
    if (threadIdx.x == 0) {
        for (int s = 0; s < SLOTS; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], NU_WARPS);
            mbar_init(&pre_bar[s], 1);
        }
#pragma unroll
        for (int f = 0; f < NW_MAX; ++f) s_fk[f] = p.F.k[f];
        mbar_fence_init();
    }
    __syncthreads();
    const unsigned lane = lane_id();

    if (threadIdx.x < 32) {
        // ================= loader warp: lane f owns file f =================
        const bool mine = lane < NWAY;
        const uint64_t* fk = mine ? s_fk[lane] : nullptr;
        // software pipeline: cut positions two tiles ahead, the unaligned head / tail elements one tile ahead, so that no
        // global-memory latency sits between a freed slot and the TMA issue of its next tile
        long long lo_a = 0, nn_a = 0, lo_b = 0, nn_b = 0;
        uint64_t hv_a = 0, tv_a = 0;
        auto fetch_bounds = [&](int li, long long* lo, long long* nn) {
            *lo = 0;
            *nn = 0;
            if (mine && li < n_my) {
                const int t = (int)blockIdx.x + li * G;
                *lo = p.bounds[t].pos[lane];
                *nn = p.bounds[t + 1].pos[lane] - *lo;
            }
        };
        auto fetch_edges = [&](long long lo, long long nn, uint64_t* hv, uint64_t* tv) {
            *hv = 0;
            *tv = 0;
            if (mine && nn > 0 && nn <= SH::CAP) {
                *hv = fk[lo];
                *tv = fk[lo + nn - 1];
            }
        };
        fetch_bounds(0, &lo_a, &nn_a);
        fetch_edges(lo_a, nn_a, &hv_a, &tv_a);
        fetch_bounds(1, &lo_b, &nn_b);
        for (int li = 0; li < n_my; ++li) {
            const int s = li % SLOTS, u = li / SLOTS;
            uint64_t hv_b, tv_b;
            long long lo_c, nn_c;
            fetch_edges(lo_b, nn_b, &hv_b, &tv_b);  // in flight until the next iteration uses them
            fetch_bounds(li + 2, &lo_c, &nn_c);
            const long long lo = lo_a;
            int n = (nn_a < 0 || nn_a > SH::CAP) ? -1 : (int)nn_a;
            const uint64_t hv = hv_a, tv = tv_a;
            lo_a = lo_b; nn_a = nn_b; hv_a = hv_b; tv_a = tv_b;
            lo_b = lo_c; nn_b = nn_c;
            if (u > 0 && !mbar_wait(&empty_bar[s], (unsigned)(u - 1) & 1u)) {
                if (lane == 0) atomicExch(p.err, (int)UKM_E_INTERNAL);
            }
            int sum = n;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
            const bool bad = __any_sync(0xffffffffu, n < 0) || sum > SH::CAP;  // cannot happen after the partition's check
            if (bad) {
                if (lane == 0) atomicExch(p.err, (int)UKM_E_INTERNAL);
                n = 0;
            }
            const int h = (n > 0) ? nu_slice_h(fk, lo) : 0;
            const int padded = (h + n + 1) & ~1;
            int incl = padded;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, d);
                if ((int)lane >= d) incl += v;
            }
            const int base = incl - padded;  // even
            uint64_t* slot = s_slots + (size_t)s * SH::BUF_E;
            int head = 0, body = 0;
            if (n > 0) {
                head = h ? 1 : 0;
                body = (n - head) & ~1;
                if (head) slot[base + h] = hv;
                if (head + body < n) slot[base + h + n - 1] = tv;
            }
            unsigned bytes = (unsigned)body * 8u;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, d);
            if (lane < NWAY) {
                s_geom[s].n[lane] = n;
                s_geom[s].off[lane] = base + h;
            }
            __syncwarp();
            if (lane == 0) {
                rw_build_tables<NWAY>(&s_geom[s]);
                mbar_expect_tx(&full_bar[s], bytes);  // arrive (release: publishes the plain stores and the tables) + tx count
            }
            __syncwarp();
            if (body) tma_load_1d(slot + base + h + head, fk + lo + head, (unsigned)body * 8u, &full_bar[s]);
        }
        return;
    }
    if (threadIdx.x < 64) {
        // ================= prefix warp (same scheme as nway_kernel / setop_pipe_kernel) =================
        constexpr int MAXM = NU_MAX_GRID / 32;
        unsigned long long P = 0;  // outputs of all earlier grid iterations (identical on every CTA)
        for (int bi = 0; bi < n_my; ++bi) {
            const int s = bi % SLOTS;
            const int tile0 = bi * G;
            const int n_iter = (p.num_tiles - tile0) < G ? (p.num_tiles - tile0) : G;
            unsigned long long val[MAXM];
            unsigned have = 0;
            unsigned spins = 0;
#pragma unroll
            for (int m = 0; m < MAXM; ++m) {
                val[m] = 0;
                if ((int)lane + 32 * m >= n_iter) have |= 1u << m;
            }
            while (true) {
#pragma unroll
                for (int m = 0; m < MAXM; ++m) {
                    if (!(have & (1u << m))) {
                        const uint64_t wd = ld_relaxed_u64(&p.status[tile0 + (int)lane + 32 * m]);
                        if (wd >> 62) {
                            val[m] = UKM_LB_VALUE(wd);
                            have |= 1u << m;
                        }
                    }
                }
                if (__all_sync(0xffffffffu, have == ((1u << MAXM) - 1))) break;
                __nanosleep(100);
                if (++spins > UKM_WATCHDOG_SPINS) {
                    if (lane == 0) atomicExch(p.err, (int)UKM_E_INTERNAL);
                    break;
                }
            }
            unsigned long long before = 0, all = 0;
#pragma unroll
            for (int m = 0; m < MAXM; ++m) {
                all += val[m];
                if ((int)lane + 32 * m < (int)blockIdx.x) before += val[m];
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                before += __shfl_xor_sync(0xffffffffu, before, d);
                all += __shfl_xor_sync(0xffffffffu, all, d);
            }
            if (lane == 0) {
                s_pre[s] = P + before;
                if (tile0 + n_iter == p.num_tiles && (int)blockIdx.x == n_iter - 1) *p.total_out = P + all;
                mbar_arrive(&pre_bar[s]);
            }
            P += all;
            __syncwarp();
        }
        return;
    }

    // ================= consumers =================
    const int tid = (int)threadIdx.x - NU_AUX;
    const int w = tid >> 5;
    const uint32_t x_a = smem_u32(s_x);
    // copy the staged distinct keys of tile ip out (its offset has arrived by now, or arrives while we wait here)
    auto copy_out = [&](int ip) {
        const int sp = ip % SLOTS, up = ip / SLOTS;
        const uint64_t* prev = s_slots + (size_t)sp * SH::BUF_E;
        if (!mbar_wait(&pre_bar[sp], (unsigned)up & 1u)) {
            if (tid == 0) atomicExch(p.err, (int)UKM_E_INTERNAL);
        }
        const unsigned long long prefix = s_pre[sp];
        const unsigned n_prev = (unsigned)s_cnt[sp];
        uint64_t* out = p.outK + prefix;
        for (unsigned q = tid; q < n_prev; q += NU_NT) out[q] = prev[q];
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[sp]);  // release: this warp's reads of the slot are done
    };
    for (int i = 0; i < n_my; ++i) {
        const int s = i % SLOTS, u = i / SLOTS;
        uint64_t* slot = s_slots + (size_t)s * SH::BUF_E;
        const uint32_t slot_a = smem_u32(slot);
        if (!mbar_wait(&full_bar[s], (unsigned)u & 1u)) {
            if (tid == 0) atomicExch(p.err, (int)UKM_E_INTERNAL);
        }
        const RwGeom<NWAY>& g = s_geom[s];
        uint64_t row[16];
        unsigned emit = 0;
        // The staged keys of tile i - 1 sit in the other slot; they can leave any time after the first barrier of this
        // tile (every consumer has staged by then).  Later = more slack for the offset of tile i - 1 to arrive (every CTA
        // of the grid iteration must have counted its tile), earlier = the slot is free sooner for the TMA of tile i + 1.
        bool pending = i > 0;
        if constexpr (LEVELS == 1) {
            named_bar_sync(1, NU_NT);
            if (pending && p.copy_pos <= 2) { copy_out(i - 1); pending = false; }
            nu_level<NWAY, 1>(g, slot_a, x_a, w, lane, row, &emit);
        } else {
            nu_level<NWAY, 1>(g, slot_a, x_a, w, lane, row, &emit);
            named_bar_sync(1, NU_NT);  // level 1 complete
            if (pending && p.copy_pos <= 1) { copy_out(i - 1); pending = false; }
            if constexpr (LEVELS == 2) {
                nu_level<NWAY, 2>(g, x_a, slot_a, w, lane, row, &emit);
                if (pending && p.copy_pos <= 2) { copy_out(i - 1); pending = false; }
            } else {
                nu_level<NWAY, 2>(g, x_a, slot_a, w, lane, row, &emit);
                named_bar_sync(1, NU_NT);
                if (pending && p.copy_pos <= 2) { copy_out(i - 1); pending = false; }
                nu_level<NWAY, 3>(g, slot_a, x_a, w, lane, row, &emit);
            }
        }
        unsigned tile_total;
        unsigned o = group_excl_scan_u32<NU_NT>((unsigned)__popc(emit), (unsigned)tid, s_scan, &tile_total, 1);
        // every consumer is past its reads of the slot and of X (two barriers inside the scan): stage in place
        if (tid == 0) {
            s_cnt[s] = tile_total;
            st_relaxed_u64(&p.status[(int)blockIdx.x + i * G], UKM_LB_PARTIAL | (uint64_t)tile_total);
        }
        if (pending) copy_out(i - 1);
#pragma unroll
        for (int it = 0; it < 16; ++it) {
            if (emit & (1u << it)) slot[o++] = row[it];
        }
    }
    named_bar_sync(1, NU_NT);
    copy_out(n_my - 1);
}

template <int NWAY>
int launch_nunion(ukm_ctx* ctx, NuArgs a, long long total, ukm_tmp& tmp, bool* fell_back) {
    using SH = NuShape<NWAY>;
    constexpr size_t smem = (size_t)(NU_SLOTS + 1) * SH::BUF_E * 8;
    auto kern = nunion_kernel<NWAY>;
    int ctas_per_sm = 0;
    UKM_TRY(ukm_kernel_config(ctx, kern, smem, NU_NT + NU_AUX, &ctas_per_sm));
    NwBound* d_bounds = nullptr;
    uint64_t* d_status = nullptr;
    int num_tiles = 0;
    UKM_TRY(ukm_nway_partition(ctx, a.F, total, SH::TILE, SH::CAP, tmp, &d_bounds, &d_status, &num_tiles, fell_back));
    if (*fell_back) return UKM_OK;  // some key occurs far too often for a tile: inputs are not duplicate-free
    a.bounds = d_bounds;
    a.status = d_status;
    a.total_out = reinterpret_cast<unsigned long long*>(d_status + num_tiles);
    a.num_tiles = num_tiles;
    int grid = ctas_per_sm * ctx->sm_count;
    if (grid > NU_MAX_GRID) grid = NU_MAX_GRID;
    if (grid > num_tiles) grid = num_tiles;
    UKM_TRY(ukm_launch_coop(ctx, kern, grid, NU_NT + NU_AUX, smem, a));  // the offset hand-off needs the whole grid resident
    UKM_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch, a.total_out, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    UKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    tmp.free_now(d_bounds);
    tmp.free_now(d_status);
    return UKM_OK;
}

}  // namespace

// Opt-in (UKM_NUNION=1).  Measured on B200, C3 (8 x 5e8 keys): 44.7 ms against 33.2 ms for the sequential-walk kernel of
// nway.cu.  The row levels do what they were built for -- shared-memory wavefronts per key halve (5.4e9 against 1.1e10 per
// union at equal tile sizes, the gather of the row levels is conflict-free) -- but the bitonic merge network costs as many
// instructions as the walk it replaces (32 compare-exchanges x 6 instructions on 64-bit keys per 16 keys and level:
// ~36 instructions per key and level, 3.9e10 warp instructions per union with the spin loops), and with two slots the
// offset hand-off between the CTAs of a grid iteration has less than one tile of slack: half of the issued instructions
// are polls.  A three-level merge of 64-bit keys is issue-bound on this chip either way.
bool ukm_nunion_enabled() {
    const char* e = getenv("UKM_NUNION");
    return e && e[0] == '1';
}

// Union of nf (2..8) sorted duplicate-free device arrays into outK (capacity >= sum of the lengths).
// *fell_back = true (and nothing written) when the inputs cannot be tiled -- the caller then uses the two-way tree.
int ukm_nunion(ukm_ctx* ctx, const uint64_t* const* keys, const size_t* n, int nf, uint64_t* outK, size_t* n_out, bool* fell_back) {
    *fell_back = false;
    *n_out = 0;
    if (nf < 2 || nf > NW_MAX) return ukm_fail(ctx, UKM_E_ARG, "nunion: 2..8 inputs");
    NuArgs a;
    long long total = 0;
    for (int f = 0; f < NW_MAX; ++f) {
        a.F.k[f] = f < nf ? keys[f] : nullptr;
        a.F.n[f] = f < nf ? (long long)n[f] : 0;
        total += a.F.n[f];
    }
    a.F.nf = nf;
    if (total == 0) return UKM_OK;
    a.outK = outK;
    a.err = ctx->d_err;
    {
        const char* e = getenv("UKM_NUNION_COPYPOS");  // A/B runs
        const int v = e ? atoi(e) : 2;
        a.copy_pos = (v >= 1 && v <= 3) ? v : 2;
    }
    ukm_tmp tmp(ctx);
    {
        ukm_stat_scope st(ctx, "setop_union_nway", (double)total * 8.0);  // every input key read once (+ the output, added below)
        int r;
        if (nf <= 2) r = launch_nunion<2>(ctx, a, total, tmp, fell_back);
        else if (nf <= 4) r = launch_nunion<4>(ctx, a, total, tmp, fell_back);
        else r = launch_nunion<8>(ctx, a, total, tmp, fell_back);
        UKM_TRY(r);
    }
    if (*fell_back) {
        if (ctx->stats_on && !ctx->pending.empty()) ctx->pending.back().bytes = 0;
        return UKM_OK;
    }
    *n_out = (size_t)ctx->h_scratch[0];
    if (ctx->stats_on && !ctx->pending.empty()) ctx->pending.back().bytes += (double)*n_out * 8.0;
    return UKM_OK;
}
