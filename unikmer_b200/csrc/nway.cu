// nway.cu -- single-pass N-way (N <= 8) union of sorted duplicate-free k-mer streams.
//
// Replaces the hash-set union of union.go:186-208 and its key sort (union.go:260-305) -- and the levels of the
// two-way merge tree this library used before -- with ONE pass over the inputs: the key space is cut into tiles
// of ~TILE elements summed over all files (partition kernels below: multi-sequence selection on
// rank(K) = sum_f lower_bound(F_f, K) with pivots taken from the data, three levels), every tile is brought into
// shared memory with one 1-D TMA bulk copy per file and merged there in log2(N) levels of two-way merge-path
// walks; the last level drops equal neighbours (the same k-mer in several files) and the distinct keys leave
// through the same deferred, coalesced copy-out as the two-way pipeline (setops.cu).  HBM traffic is the
// algorithmic minimum: every input byte read once, every output byte written once (a tree of two-way passes
// moves the data log2(N) times).
//
// Kernel shape (persistent, warp-specialised; grid = resident CTAs, tiles round-robin):
//   warp 0 loader : lane f owns file f -- tile geometry, unaligned head/tail by plain loads, body by TMA
//                   (cp.async.bulk -> mbarrier complete_tx); lane 0 lays out the merge tables of the tile.
//   warp 1 prefix : output offsets of the grid iteration (gathers the G counts of tiles [i*G, (i+1)*G)).
//   warps 2+ consumers: level 1 slot -> X, level 2 X -> slot, last level -> registers (+ unique flags),
//                   scan, publish count, copy tile i-DEFER out, stage this tile's keys in place.
#include <stdlib.h>

#include "common.cuh"
#include "nway_core.cuh"

int ukm_nway_partition(ukm_ctx* ctx, const NwFiles& F, long long total, int tile, int cap, ukm_tmp& tmp, NwBound** d_bounds_out,
                       uint64_t** d_status_out, int* num_tiles_out, bool* bad);

namespace {

// ---------------------------------------------------------------------------------------------------
// partition kernels: one thread per boundary, three levels (every 64th tile from the global bracket, every
// 8th from those, every tile from those) so that the bracket of a search is always small
// ---------------------------------------------------------------------------------------------------
struct NwPartArgs {
    NwFiles F;
    long long total;
    long long tile;  // nominal elements per tile
    long long tol;   // accepted rank error of a boundary
    int num_tiles;
};

// bounds[t] = the cut in front of tile t (t = num_tiles: the end of every file).  One launch per level:
// boundaries at multiples of `stride` that the level above (multiples of `parent`; 0 = none) has not produced.
__global__ void nway_partition_kernel(const NwPartArgs a, NwBound* __restrict__ bounds, int stride, int parent) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long t = i * stride;
    if (t > a.num_tiles && !(parent == 0 && i == 0)) return;
    if (parent == 0) {
        NwBound lo, hi, out;
        nw_global_bracket(a.F, &lo, &hi);
        if (i == 0) bounds[a.num_tiles] = hi;  // the end boundary
        if (t >= a.num_tiles) return;
        if (t == 0) out = lo;
        else nw_refine(a.F, t * a.tile, a.tol, lo, hi, &out);
        bounds[t] = out;
        return;
    }
    if (t >= a.num_tiles || t % parent == 0) return;
    const long long pl = t / parent * parent;
    const long long ph = pl + parent < a.num_tiles ? pl + parent : a.num_tiles;
    NwBound out;
    nw_refine(a.F, t * a.tile, a.tol, bounds[pl], bounds[ph], &out);
    bounds[t] = out;
}

// every tile must fit the shared-memory slot (it does for duplicate-free inputs); info[0] = bad flag, info[1] = largest tile
__global__ void nway_check_kernel(const NwBound* __restrict__ bounds, int num_tiles, int cap, int* __restrict__ info) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_tiles) return;
    long long sum = 0;
    bool bad = false;
#pragma unroll
    for (int f = 0; f < NW_MAX; ++f) {
        const long long d = bounds[t + 1].pos[f] - bounds[t].pos[f];
        if (d < 0) bad = true;
        sum += d;
    }
    if (sum > cap) bad = true;
    if (bad) atomicExch(&info[0], 1);
    if (sum > 0x7fffffff) sum = 0x7fffffff;
    atomicMax(&info[1], (int)sum);
}

// ---------------------------------------------------------------------------------------------------
// the tile kernel
// ---------------------------------------------------------------------------------------------------
constexpr int NWK_MAX_GRID = 512;  // prefix warp keeps NWK_MAX_GRID/32 counts per lane
constexpr int NWK_AUX = 64;        // loader + prefix warps

enum NwOp { NWOP_UNION = 0, NWOP_INTER = 1, NWOP_DIFF = 2 };

struct NwArgs {
    NwFiles F;
    const NwBound* bounds;  // bounds[t].pos[f] = first element of tile t in file f
    uint64_t* outK;
    uint64_t* status;  // one count word per tile (flag << 62 | count)
    unsigned long long* total_out;
    int num_tiles;
    int null_mode;  // measurement aids for inter / diff (results are NOT valid): UKM_NWAY_NULL=1 tiles do no work, =2 load pipeline only
    // union only, optional: inter and diff of the same files ride along (one bit per key of file 0, by position: bit set =
    // the key is in the intersection of all files / in no other file).  NULL = plain union.
    unsigned long long* mask_i;
    unsigned long long* mask_d;
    int* err;
};

// 16-byte alignment helpers for the 1-D bulk copies (same contract as slice_* in setops.cu): element
// start+i of g lands in s[h + i], h = misalignment of g+start in elements; s is 16-byte aligned.
__device__ __forceinline__ int nw_slice_h(const uint64_t* g, long long start) {
    return (int)((reinterpret_cast<uintptr_t>(g + start) & 15u) >> 3);
}

template <int NWAY, int NT, int VT>
struct NwShape {
    static constexpr int CAP = (NT - NWAY / 2) * VT;               // most elements a tile may hold
    static constexpr int SLOT_E = (CAP + 2 * NWAY + 8 + 1) & ~1;   // + per-segment alignment slack + read-past padding
    static constexpr int X_E = (CAP + 8 + 1) & ~1;
    static constexpr int TILE = (CAP * 16 / 17) & ~31;             // nominal tile; boundaries are exact to +-TILE/32
    static constexpr int LEVELS = NwGeom<NWAY>::LEVELS;
};

// ---- inter / diff of one tile: which of file 0's keys occur in the other files -----------------------------
// File 0's keys of the tile are the candidates; they are looked up (branch-free binary search in shared memory) in
// the other files' segments.  The first NWF_SEQ files run one after the other over a shrinking work list -- after
// each file only the survivors stay (inter: found, diff: not found), which is inter.go:205-286 / diff.go:380-435
// with its early exit, on a tile -- and the remaining files are probed together, one (candidate, file) pair per
// thread, clearing the candidate's flag.  Work lists are unordered (a shared counter hands out slots); the flags
// are per position in file 0's segment, so the final compaction restores key order.
// (An earlier version hashed file 0's keys and probed every key of every other file: 133 instructions per key,
// 28-34 ms per C3 operation -- DESIGN.md 4.3.)
constexpr int NWF_SEQ = 3;

__device__ __forceinline__ bool nwf_contains(const uint64_t* seg, int n, uint64_t x) {
    int pos = 0;  // lower_bound of x; the answer lies in [pos, pos + n]
    int m = n;
    while (m > 1) {
        const int half = m >> 1;
        if (seg[pos + half] < x) pos += half;
        m -= half;
    }
    if (m == 1 && seg[pos] < x) ++pos;
    return pos < n && seg[pos] == x;
}

template <int OP, int NT, int VT>
__device__ __forceinline__ unsigned nw_filter_tile(const uint64_t* slot, const NwGeom<NW_MAX>& g, uint64_t* xreg, int* s_cnt3, int tid,
                                                   int nf, bool null_mode, uint64_t* outk) {
    constexpr int CAP = NwShape<NW_MAX, NT, VT>::CAP;
    uint8_t* alive = reinterpret_cast<uint8_t*>(xreg);                                 // CAP bytes (+ padding)
    uint16_t* list0 = reinterpret_cast<uint16_t*>(xreg + (CAP + 7) / 8 + 1);            // CAP entries each
    uint16_t* list1 = list0 + ((CAP + 3) & ~3);
    const int n0 = g.n[0];
    const uint64_t* seg0 = slot + g.off[0];
    const int ept = (n0 + NT - 1) / NT;  // file-0 keys per thread in the ordered pass: thread t owns [t*ept, (t+1)*ept)
    unsigned mask = 0;
#pragma unroll
    for (int r = 0; r < VT; ++r) outk[r] = 0;
    if (null_mode) return 0;  // measurement aid: the tile machinery without any work
    const int nsub = nf - 1;
    const int nseq = nsub < NWF_SEQ ? nsub : NWF_SEQ;
    // flags start at 0; the survivors of the sequential rounds set theirs
    for (int j = tid; j < (n0 + 3) / 4; j += NT) reinterpret_cast<uint32_t*>(alive)[j] = 0;
    if (tid < 3) s_cnt3[tid] = 0;
    named_bar_sync(1, NT);
    // sequential rounds over files 1 .. nseq: list (round - 1) -> list (round)
    int cnt_in = n0;
    for (int f = 1; f <= nseq; ++f) {
        const uint16_t* lin = (f & 1) ? list1 : list0;  // round 1 reads the identity list (no array)
        uint16_t* lout = (f & 1) ? list0 : list1;
        const uint64_t* seg = slot + g.off[f];
        const int nfseg = g.n[f];
        int* c_out = &s_cnt3[f % 3];
        for (int i = tid; i < cnt_in; i += NT) {
            const int idx = f == 1 ? i : (int)lin[i];
            const uint64_t x = seg0[idx];
            const bool found = nwf_contains(seg, nfseg, x);
            if (OP == NWOP_INTER ? found : !found) lout[atomicAdd(c_out, 1)] = (uint16_t)idx;
        }
        if (tid == 0) s_cnt3[(f + 1) % 3] = 0;  // the counter of the round after next (nobody reads or adds to it now)
        named_bar_sync(1, NT);
        cnt_in = *c_out;
    }
    const uint16_t* lfin = (nseq & 1) ? list0 : list1;
    for (int i = tid; i < cnt_in; i += NT) alive[nseq == 0 ? i : (int)lfin[i]] = 1;
    named_bar_sync(1, NT);
    // the remaining files together: one (survivor, file) pair per thread, a miss (inter) / a hit (diff) clears the flag
    const int npar = nsub - nseq;
    if (npar > 0) {
        const int pairs = cnt_in * npar;
        for (int q = tid; q < pairs; q += NT) {
            const int i = q / npar, f = nseq + 1 + (q - i * npar);
            const int idx = (int)lfin[i];
            const bool found = nwf_contains(slot + g.off[f], g.n[f], seg0[idx]);
            if (OP == NWOP_INTER ? !found : found) alive[idx] = 0;
        }
        named_bar_sync(1, NT);
    }
    // ordered pass: my file-0 positions, in key order
#pragma unroll
    for (int r = 0; r < VT; ++r) {
        const int i0 = tid * ept + r;
        if (r < ept && i0 < n0 && alive[i0]) {
            outk[r] = seg0[i0];
            mask |= 1u << r;
        }
    }
    return mask;
}

template <int OP, int NWAY, int NT, int VT>
constexpr int nw_x_elems() {  // the second shared-memory region: merge buffer X (union) or flags + two work lists (inter / diff)
    using SH = NwShape<NWAY, NT, VT>;
    return OP == NWOP_UNION ? SH::X_E : (SH::CAP + 7) / 8 + 1 + 2 * (((SH::CAP + 3) & ~3) / 4) + 2;
}

// inter / diff riding along the union: set the bit of key x in the mask of the intersection (one = false) or of the
// difference (one = true) if x is one of file 0's n0 keys of the tile at f0 (global position g0 of the first)
__device__ __forceinline__ void nw_mark_f0(const uint64_t* f0, int n0, long long g0, uint64_t x, bool one, unsigned long long* mask_i,
                                           unsigned long long* mask_d) {
    int lo_ = 0, hi_ = n0;  // lower_bound of x
    while (lo_ < hi_) {
        const int mid = (lo_ + hi_) >> 1;
        if (f0[mid] < x) lo_ = mid + 1;
        else hi_ = mid;
    }
    if (lo_ < n0 && f0[lo_] == x) {
        const unsigned long long idx = (unsigned long long)(g0 + lo_);
        atomicOr((one ? mask_d : mask_i) + (idx >> 6), 1ull << (idx & 63));
    }
}

template <int OP, int NWAY, int NT, int VT, int SLOTS, int MINB, int DEFER = 1>
__global__ void __launch_bounds__(NT + NWK_AUX, MINB) nway_kernel(const NwArgs p) {
    using SH = NwShape<NWAY, NT, VT>;
    static_assert(OP == NWOP_UNION || NWAY == NW_MAX, "inter / diff tiles are laid out for 8 files");
    constexpr int NW = NT / 32;
    static_assert(DEFER >= 1 && DEFER <= SLOTS - 2, "copy-out lag in tiles; the loader runs SLOTS - DEFER - 1 tiles ahead");
    constexpr int LEVELS = SH::LEVELS;
    extern __shared__ __align__(16) unsigned char nw_smem[];
    uint64_t* s_slots = reinterpret_cast<uint64_t*>(nw_smem);  // SLOTS * SLOT_E
    uint64_t* s_x = s_slots + (size_t)SLOTS * SH::SLOT_E;      // nw_x_elems()
    __shared__ __align__(8) uint64_t full_bar[SLOTS], empty_bar[SLOTS], pre_bar[SLOTS];
    __shared__ unsigned long long s_cnt[SLOTS], s_pre[SLOTS];
    __shared__ NwGeom<NWAY> s_geom[SLOTS];
    __shared__ const uint64_t* s_fk[NW_MAX];
    __shared__ unsigned s_scan[NW + 2];
    __shared__ int s_cnt3[4];  // rotating work-list counters of the inter / diff rounds
    __shared__ long long s_g0[SLOTS];  // first element of the tile in file 0 (global position)
    __shared__ unsigned char s_lead[NT];  // leading keys of a thread's range that continue the run of the thread before
    // inter / diff riding along need file 0's segment of the tile at the last level; with three levels the slot has been
    // overwritten by then (level 2 writes it), so a copy is kept here when it fits (else: the global array, slow, rare)
    constexpr int F0CAP = (OP == NWOP_UNION && LEVELS == 3) ? 512 : 1;
    __shared__ uint64_t s_f0[F0CAP];
    constexpr int NCAND = OP == NWOP_UNION ? 128 : 1;  // candidate list of a tile (a run of nf or of 1 key)
    __shared__ uint64_t s_ckey[NCAND];
    __shared__ unsigned char s_cone[NCAND];
    __shared__ int s_ncand;

    const int G = gridDim.x;
    const int n_my = (p.num_tiles - (int)blockIdx.x + G - 1) / G;  // tiles of this CTA
    if (threadIdx.x == 0) {
        for (int s = 0; s < SLOTS; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], NT);
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
        const uint64_t* fk = lane < NWAY ? s_fk[lane] : nullptr;
        // the cut positions of a tile are fetched one tile ahead, so their (L2 / HBM) latency is not on the per-tile path
        long long lo_nx = 0, hi_nx = 0;
        if (lane < NWAY && n_my > 0) {
            lo_nx = p.bounds[blockIdx.x].pos[lane];
            hi_nx = p.bounds[blockIdx.x + 1].pos[lane];
        }
        for (int li = 0; li < n_my; ++li) {
            const int s = li % SLOTS, u = li / SLOTS;
            long long lo = lo_nx;
            int n = (int)(hi_nx - lo_nx);
            if (lane < NWAY && li + 1 < n_my) {
                const int tnx = (int)blockIdx.x + (li + 1) * G;
                lo_nx = p.bounds[tnx].pos[lane];
                hi_nx = p.bounds[tnx + 1].pos[lane];
            }
            if (u > 0 && !mbar_wait(&empty_bar[s], (unsigned)(u - 1) & 1u)) {
                if (lane == 0) atomicExch(p.err, (int)UKM_E_INTERNAL);
            }
            int sum = n;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
            const bool bad = __any_sync(0xffffffffu, n < 0) || sum > SH::CAP;  // cannot happen after nway_check
            if (bad) {
                if (lane == 0) atomicExch(p.err, (int)UKM_E_INTERNAL);
                n = 0;
            }
            const int h = (n > 0) ? nw_slice_h(fk, lo) : 0;
            const int padded = (h + n + 1) & ~1;
            int incl = padded;  // inclusive scan over the lanes
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, d);
                if ((int)lane >= d) incl += v;
            }
            const int base = incl - padded;  // even
            uint64_t* slot = s_slots + (size_t)s * SH::SLOT_E;
            // 16-byte aligned body through TMA, the (at most one element) head and tail through plain loads
            int head = 0, body = 0;
            if (n > 0) {
                head = h ? 1 : 0;
                body = (n - head) & ~1;
                if (head) slot[base + h] = fk[lo];
                if (head + body < n) slot[base + h + n - 1] = fk[lo + n - 1];
            }
            unsigned bytes = (unsigned)body * 8u;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, d);
            if (lane < NWAY) {
                s_geom[s].n[lane] = n;
                s_geom[s].off[lane] = base + h;
            }
            if (lane == 0) s_g0[s] = lo;
            __syncwarp();
            if (lane == 0) {
                if (OP == NWOP_UNION) nw_build_tables<NWAY, VT>(&s_geom[s]);
                mbar_expect_tx(&full_bar[s], bytes);  // arrive (release: publishes the plain stores and the tables) + tx count
            }
            __syncwarp();
            if (body) tma_load_1d(slot + base + h + head, fk + lo + head, (unsigned)body * 8u, &full_bar[s]);
        }
        return;
    }
    if (threadIdx.x < 64) {
        // ================= prefix warp (same scheme as setop_pipe_kernel) =================
        if (OP != NWOP_UNION && p.null_mode == 2) return;  // measurement aid: no output offsets at all
        constexpr int MAXM = NWK_MAX_GRID / 32;
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
                        const uint64_t w = ld_relaxed_u64(&p.status[tile0 + (int)lane + 32 * m]);
                        if (w >> 62) {
                            val[m] = UKM_LB_VALUE(w);
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
    const int tid = (int)threadIdx.x - NWK_AUX;
    if (OP != NWOP_UNION && p.null_mode == 2) {
        // measurement aid (UKM_NWAY_NULL=2): the load pipeline alone -- wait for a tile, hand its slot back
        for (int i = 0; i < n_my; ++i) {
            const int s = i % SLOTS, u = i / SLOTS;
            if (!mbar_wait(&full_bar[s], (unsigned)u & 1u)) {
                if (tid == 0) atomicExch(p.err, (int)UKM_E_INTERNAL);
            }
            mbar_arrive(&empty_bar[s]);
        }
        if (tid == 0 && blockIdx.x == 0) *p.total_out = 0;
        return;
    }
    for (int i = 0; i < n_my + DEFER; ++i) {
        unsigned emitmask = 0;
        uint64_t outk[VT];
        unsigned off = 0;
        int last_steps = 0;
        uint64_t* slot = nullptr;
        if (i < n_my) {
            const int s = i % SLOTS, u = i / SLOTS;
            slot = s_slots + (size_t)s * SH::SLOT_E;
            if (!mbar_wait(&full_bar[s], (unsigned)u & 1u)) {
                if (tid == 0) atomicExch(p.err, (int)UKM_E_INTERNAL);
            }
            const NwGeom<NWAY>& g = s_geom[s];
            if (OP == NWOP_UNION && tid == 0 && p.mask_i) s_ncand = 0;  // read again only after the barriers of the scan
            if constexpr (OP != NWOP_UNION) {
                emitmask = nw_filter_tile<OP, NT, VT>(slot, g, s_x, s_cnt3, tid, p.F.nf, p.null_mode != 0, outk);
            } else {
            const uint64_t* src = slot;
            uint64_t* dst = s_x;
            if (LEVELS == 3 && p.mask_i && g.n[0] <= F0CAP)
                for (int j = tid; j < g.n[0]; j += NT) s_f0[j] = slot[g.off[0] + j];
            // inner levels: plain two-way merges, every pair of runs by its own group of threads
#pragma unroll
            for (int l = 1; l < LEVELS; ++l) {
                const int npairs = NWAY >> l;
                const int p0 = nw_pair0<NWAY>(l), t0 = nw_tb0<NWAY>(l);
                int j = 0, m;
                if (npairs == 4) m = nw_find_pair<4>(g.tb + t0, tid, &j);
                else m = nw_find_pair<2>(g.tb + t0, tid, &j);
                if (m >= 0) {
                    const NwPair pr = g.pair[p0 + m];
                    const int diag = j * VT;
                    int steps = pr.lenA + pr.lenB - diag;
                    if (steps > VT) steps = VT;
                    const uint64_t* A = src + pr.srcA;
                    const uint64_t* B = src + pr.srcB;
                    const int a = nw_merge_path_g(A, pr.lenA, B, pr.lenB, diag);
                    nw_walk_plain<VT>(A, pr.lenA, B, pr.lenB, a, diag - a, steps, dst + pr.dst + diag);
                }
                named_bar_sync(1, NT);
                // level 1: slot -> X; level 2: X -> slot
                const uint64_t* t = src;
                src = dst;
                dst = const_cast<uint64_t*>(t);
            }
            // last level: merge into registers, keep the first key of every run of equal keys
            {
                const NwPair pr = g.pair[NWAY - 2];
                const int tot = pr.lenA + pr.lenB;
                int diag = tid * VT;
                int steps = tot - diag;
                if (steps > VT) steps = VT;
                if (diag > tot) diag = tot;
                const uint64_t* A = src + pr.srcA;
                const uint64_t* B = src + pr.srcB;
                const int a = nw_merge_path_g(A, pr.lenA, B, pr.lenB, diag);
                emitmask = nw_walk_unique<VT>(A, pr.lenA, B, pr.lenB, a, diag - a, steps, outk);
                last_steps = steps > 0 ? steps : 0;
                // the leading keys that were not emitted continue the run the thread before started
                // the run heads among the first 8 positions (positions past the range count as heads: the tile ends there)
                if (p.mask_i) s_lead[tid] = (unsigned char)nw_lead_heads(emitmask, last_steps);
            }
            }  // union
            // block scan of the distinct-key counts (group_excl_scan_u32, spelled out: the candidates of the riding inter / diff
            // are collected between its two barriers, looked up after the second -- no barrier of their own)
            unsigned tile_total;
            {
                const unsigned cnt_ = (unsigned)__popc(emitmask);
                const unsigned incl_ = warp_incl_scan_u32(cnt_);
                const unsigned w_ = (unsigned)tid >> 5;
                if (lane == 31) s_scan[w_] = incl_;
                named_bar_sync(1, NT);  // every s_lead is visible too
                if constexpr (OP == NWOP_UNION) {
                    if (p.mask_i) {
                        // inter / diff ride along: a run of equal keys has one key per file that holds it (the inputs are
                        // duplicate-free), so a run as long as the file count is a key of the intersection, and a run of one
                        // whose key sits in file 0 is a key of the difference.  Runs are at most nf <= 8 < VT keys long: one
                        // that starts here can only continue into the NEXT thread (s_lead).  The candidates (few: a run of
                        // nf or of 1) go to a list; after the second barrier one thread per candidate finds the key's position
                        // in file 0 -- the bit to set -- by binary search in file 0's keys of the tile.
                        const int n0 = g.n[0];
                        // file 0's keys of the tile: still in the slot (one or two levels), the copy made before level 1 (three
                        // levels), or -- the copy did not fit -- the global array
                        const uint64_t* f0 = LEVELS < 3 ? slot + g.off[0] : (n0 <= F0CAP ? s_f0 : s_fk[0] + s_g0[s]);
                        const unsigned next_heads = tid + 1 < NT ? (unsigned)s_lead[tid + 1] : 0xffu;
                        unsigned run_nf, run_one;
                        unsigned m = nw_run_candidates(emitmask, last_steps, next_heads, p.F.nf, &run_nf, &run_one);
                        while (m) {  // rarely more than one iteration
                            const int b0 = __ffs(m) - 1;
                            m &= m - 1;
                            const bool one = (run_one >> b0) & 1u;
                            uint64_t x = outk[0];
#pragma unroll
                            for (int it = 1; it < VT; ++it)
                                if (it == b0) x = outk[it];
                            const int pos = atomicAdd(&s_ncand, 1);
                            if (pos < NCAND) {
                                s_ckey[pos] = x;
                                s_cone[pos] = one;
                            } else {
                                nw_mark_f0(f0, n0, s_g0[s], x, one, p.mask_i, p.mask_d);  // more candidates than the list holds: in place
                            }
                        }
                    }
                }
                if (w_ == 0) {
                    const unsigned x_ = (lane < (unsigned)NW) ? s_scan[lane] : 0u;
                    const unsigned xi_ = warp_incl_scan_u32(x_);
                    if (lane < (unsigned)NW) s_scan[lane] = xi_ - x_;
                    if (lane == NW - 1) s_scan[NW] = xi_;
                }
                named_bar_sync(1, NT);  // the candidate list is complete
                tile_total = s_scan[NW];
                off = s_scan[w_] + incl_ - cnt_;
            }
            if constexpr (OP == NWOP_UNION) {
                if (p.mask_i) {
                    const int n0 = g.n[0];
                    const uint64_t* f0 = LEVELS < 3 ? slot + g.off[0] : (n0 <= F0CAP ? s_f0 : s_fk[0] + s_g0[s]);
                    const int nc = s_ncand < NCAND ? s_ncand : NCAND;
                    for (int j = tid; j < nc; j += NT) nw_mark_f0(f0, n0, s_g0[s], s_ckey[j], s_cone[j] != 0, p.mask_i, p.mask_d);
                    // file 0's keys were read from the slot, which the staging below overwrites
                    if (LEVELS < 3) named_bar_sync(1, NT);
                }
            }
            // every consumer is past its reads of the slot (two barriers inside the scan): it may be overwritten
            if (tid == 0) {
                s_cnt[s] = tile_total;
                st_relaxed_u64(&p.status[(int)blockIdx.x + i * G], UKM_LB_PARTIAL | (uint64_t)tile_total);
            }
        }
        // copy tile i-DEFER out while this tile's count travels
        if (i >= DEFER && i - DEFER < n_my) {
            const int ip = i - DEFER;
            const int sp = ip % SLOTS, up = ip / SLOTS;
            const uint64_t* prev = s_slots + (size_t)sp * SH::SLOT_E;
            if (!mbar_wait(&pre_bar[sp], (unsigned)up & 1u)) {
                if (tid == 0) atomicExch(p.err, (int)UKM_E_INTERNAL);
            }
            const unsigned long long prefix = s_pre[sp];
            const unsigned n_prev = (unsigned)s_cnt[sp];
            uint64_t* out = p.outK + prefix;
            for (unsigned j = tid; j < n_prev; j += NT) out[j] = prev[j];
            mbar_arrive(&empty_bar[sp]);  // release: my reads of the slot are done
        }
        // stage this tile's distinct keys in place
        if (i < n_my) {
            unsigned o = off;
#pragma unroll
            for (int it = 0; it < VT; ++it) {
                if (emitmask & (1u << it)) slot[o++] = outk[it];
            }
        }
        named_bar_sync(1, NT);  // staged tile (and s_cnt) visible to every consumer before a later copy-out
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
template <int OP, int NWAY, int NT, int VT, int SLOTS, int MINB, int DEFER = 1>
int launch_nway(ukm_ctx* ctx, NwArgs a, NwPartArgs pa, ukm_tmp& tmp, bool* fell_back) {
    using SH = NwShape<NWAY, NT, VT>;
    constexpr size_t smem = ((size_t)SLOTS * SH::SLOT_E + nw_x_elems<OP, NWAY, NT, VT>()) * 8;
    auto kern = nway_kernel<OP, NWAY, NT, VT, SLOTS, MINB, DEFER>;
    int ctas_per_sm = 0;
    UKM_TRY(ukm_kernel_config(ctx, kern, smem, NT + NWK_AUX, &ctas_per_sm));
    // ---- partition: tiles of ~TILE elements summed over all files ----
    NwBound* d_bounds = nullptr;
    uint64_t* d_status = nullptr;
    int num_tiles = 0;
    UKM_TRY(ukm_nway_partition(ctx, pa.F, pa.total, SH::TILE, SH::CAP, tmp, &d_bounds, &d_status, &num_tiles, fell_back));
    if (*fell_back) return UKM_OK;  // some key occurs far too often for a tile: inputs are not duplicate-free
    unsigned long long* d_total = reinterpret_cast<unsigned long long*>(d_status + num_tiles);
    a.bounds = d_bounds;
    a.status = d_status;
    a.total_out = d_total;
    a.num_tiles = num_tiles;
    int grid = ctas_per_sm * ctx->sm_count;
    if (grid > NWK_MAX_GRID) grid = NWK_MAX_GRID;
    if (grid > num_tiles) grid = num_tiles;
    UKM_TRY(ukm_launch_coop(ctx, kern, grid, NT + NWK_AUX, smem, a));  // co-residency guaranteed by the driver
    UKM_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch, d_total, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    UKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    tmp.free_now(d_bounds);
    tmp.free_now(d_status);
    return UKM_OK;
}

// tile shapes: consumer threads x keys per thread x ring slots.  UKM_NWAY_CFG picks one by index for A/B runs.
int nway_cfg() {
    const char* e = getenv("UKM_NWAY_CFG");
    const int v = e ? atoi(e) : 0;
    return (v >= 0 && v < 5) ? v : 0;
}

template <int OP, int NWAY>
int launch_nway_cfg(ukm_ctx* ctx, const NwArgs& a, const NwPartArgs& pa, ukm_tmp& tmp, bool* fell_back) {
    switch (nway_cfg()) {
        case 1: return launch_nway<OP, NWAY, 128, 17, 3, 3>(ctx, a, pa, tmp, fell_back);
        case 2: return launch_nway<OP, NWAY, 512, 13, 3, 1>(ctx, a, pa, tmp, fell_back);
        case 3: return launch_nway<OP, NWAY, 128, 17, 4, 2>(ctx, a, pa, tmp, fell_back);
        case 4: return launch_nway<OP, NWAY, 256, 9, 3, 3>(ctx, a, pa, tmp, fell_back);
        default: return launch_nway<OP, NWAY, 256, 13, 3, 2>(ctx, a, pa, tmp, fell_back);
    }
}

int nway_run(ukm_ctx* ctx, int op, const char* stat_name, const uint64_t* const* keys, const size_t* n, int nf, uint64_t* outK,
             size_t* n_out, bool* fell_back, unsigned long long* mask_i = nullptr, unsigned long long* mask_d = nullptr) {
    *fell_back = false;
    *n_out = 0;
    if (nf < 2 || nf > NW_MAX) return ukm_fail(ctx, UKM_E_ARG, "nway: 2..8 inputs");
    NwArgs a;
    NwPartArgs pa;
    long long total = 0;
    for (int f = 0; f < NW_MAX; ++f) {
        a.F.k[f] = f < nf ? keys[f] : nullptr;
        a.F.n[f] = f < nf ? (long long)n[f] : 0;
        total += a.F.n[f];
    }
    a.F.nf = nf;
    if (total == 0) return UKM_OK;
    pa.F = a.F;
    pa.total = total;
    a.outK = outK;
    a.err = ctx->d_err;
    a.null_mode = 0;
    a.mask_i = mask_i;
    a.mask_d = mask_d;
#ifdef UKM_MEASURE  // measurement build only (make EXTRA=-DUKM_MEASURE): the null modes produce no valid result
    {
        const char* e = getenv("UKM_NWAY_NULL");
        a.null_mode = e ? atoi(e) : 0;
    }
#endif
    ukm_tmp tmp(ctx);
    {
        ukm_stat_scope st(ctx, stat_name, (double)total * 8.0);  // every input key read once (+ the output, added below)
        int r;
        if (op == NWOP_INTER) r = launch_nway_cfg<NWOP_INTER, 8>(ctx, a, pa, tmp, fell_back);
        else if (op == NWOP_DIFF) r = launch_nway_cfg<NWOP_DIFF, 8>(ctx, a, pa, tmp, fell_back);
        else if (nf <= 2) r = launch_nway_cfg<NWOP_UNION, 2>(ctx, a, pa, tmp, fell_back);
        else if (nf <= 4) r = launch_nway_cfg<NWOP_UNION, 4>(ctx, a, pa, tmp, fell_back);
        else r = launch_nway_cfg<NWOP_UNION, 8>(ctx, a, pa, tmp, fell_back);
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

}  // namespace

// Tiles of ~tile elements summed over all files (boundaries exact to +- tile / 32, never above `cap`): the three-level
// multi-sequence selection above + the capacity check.  Allocates the boundaries and a zeroed status array of
// num_tiles + 4 words (tile counts, then the output total, then the check words).  *bad = the inputs cannot be tiled.
// Shared with the row-based union (nunion.cu).
int ukm_nway_partition(ukm_ctx* ctx, const NwFiles& F, long long total, int tile, int cap, ukm_tmp& tmp, NwBound** d_bounds_out,
                       uint64_t** d_status_out, int* num_tiles_out, bool* bad) {
    NwPartArgs pa;
    pa.F = F;
    pa.total = total;
    pa.tile = tile;
    pa.tol = tile / 32;
    const int num_tiles = (int)((total + tile - 1) / tile);
    pa.num_tiles = num_tiles;
    NwBound* d_bounds = nullptr;
    uint64_t* d_status = nullptr;
    UKM_TRY(tmp.alloc(&d_bounds, (size_t)num_tiles + 1));
    UKM_TRY(tmp.alloc(&d_status, (size_t)num_tiles + 4));
    int* d_info = reinterpret_cast<int*>(d_status + num_tiles + 1);
    UKM_CUDA(ctx, cudaMemsetAsync(d_status, 0, ((size_t)num_tiles + 4) * sizeof(uint64_t), ctx->stream));
    {
        constexpr int S0 = 64, S1 = 8;
        const int n0 = num_tiles / S0 + 1, n1 = num_tiles / S1 + 1;
        nway_partition_kernel<<<(n0 + 63) / 64, 64, 0, ctx->stream>>>(pa, d_bounds, S0, 0);
        UKM_LAUNCHED(ctx);
        nway_partition_kernel<<<(n1 + 63) / 64, 64, 0, ctx->stream>>>(pa, d_bounds, S1, S0);
        UKM_LAUNCHED(ctx);
        nway_partition_kernel<<<(num_tiles + 1 + 127) / 128, 128, 0, ctx->stream>>>(pa, d_bounds, 1, S1);
        UKM_LAUNCHED(ctx);
    }
    nway_check_kernel<<<(num_tiles + 127) / 128, 128, 0, ctx->stream>>>(d_bounds, num_tiles, cap, d_info);
    UKM_LAUNCHED(ctx);
    UKM_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch, d_info, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    UKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *bad = reinterpret_cast<int*>(ctx->h_scratch)[0] != 0;
    *d_bounds_out = d_bounds;
    *d_status_out = d_status;
    *num_tiles_out = num_tiles;
    return UKM_OK;
}

bool ukm_nway_enabled() {
    const char* e = getenv("UKM_NWAY");
    return !(e && e[0] == '0');
}

// Union of nf (2..8) sorted duplicate-free device arrays into outK (capacity >= sum of the lengths).
// *fell_back = true (and nothing written) when the inputs cannot be tiled -- the caller then uses the
// two-way tree, which takes any input.
int ukm_nway_union(ukm_ctx* ctx, const uint64_t* const* keys, const size_t* n, int nf, uint64_t* outK, size_t* n_out,
                   bool* fell_back) {
    return nway_run(ctx, NWOP_UNION, "setop_union_nway", keys, n, nf, outK, n_out, fell_back);
}

// keys[0] filtered by membership in keys[1..nf-1] (3..8 arrays): inter keeps the keys found in every other array,
// diff the keys found in none.  outK capacity >= n[0].  Same fall-back contract as the union.
int ukm_nway_filter(ukm_ctx* ctx, bool inter, const uint64_t* const* keys, const size_t* n, int nf, uint64_t* outK, size_t* n_out,
                    bool* fell_back) {
    return nway_run(ctx, inter ? NWOP_INTER : NWOP_DIFF, inter ? "setop_inter_nway" : "setop_diff_nway", keys, n, nf, outK, n_out,
                    fell_back);
}

int ukm_masks_gather(ukm_ctx* ctx, ukm_tmp& tmp, const unsigned long long* d_masks, size_t n_masks, const uint64_t* F0, uint64_t* outK,
                     size_t* n_out);

// union, inter and diff of the same nf (2..8) sorted duplicate-free device arrays from ONE pass over the inputs: the union
// kernel sees, in its last merge level, how many files hold every key (the length of the run of equal keys), which is all
// inter (run = nf) and diff (run = 1, key in file 0) need.  outK capacity >= sum of the lengths, outI / outD >= n[0].
// Same fall-back contract as ukm_nway_union.
int ukm_nway_union3(ukm_ctx* ctx, const uint64_t* const* keys, const size_t* n, int nf, uint64_t* outK, size_t* n_out, uint64_t* outI,
                    size_t* n_i, uint64_t* outD, size_t* n_d, bool* fell_back) {
    *n_i = *n_d = 0;
    ukm_tmp tmp(ctx);
    const size_t n_masks = (n[0] + 63) / 64;
    unsigned long long* d_masks = nullptr;
    UKM_TRY(tmp.alloc(&d_masks, 2 * n_masks + 2));
    UKM_CUDA(ctx, cudaMemsetAsync(d_masks, 0, (2 * n_masks + 2) * sizeof(unsigned long long), ctx->stream));
    UKM_TRY(nway_run(ctx, NWOP_UNION, "setop_inter_diff_union_nway", keys, n, nf, outK, n_out, fell_back, d_masks, d_masks + n_masks + 1));
    if (*fell_back) return UKM_OK;
    // algorithmic bytes are per OPERATION: every input key read once + every output key written once for each
    // of union, inter and diff.  nway_run recorded the union's; the launch also did the work of the other two (whose outputs
    // are added by the gather below).  The bytes actually moved are the union's alone -- the three share one read.
    if (ctx->stats_on && !ctx->pending.empty()) {
        double in_bytes = 0;
        for (int f = 0; f < nf; ++f) in_bytes += (double)n[f] * 8.0;
        ctx->pending.back().bytes += 2.0 * in_bytes;
    }
    {
        ukm_stat_scope st(ctx, "setop_mask_gather", (double)n_masks * 16.0);
        UKM_TRY(ukm_masks_gather(ctx, tmp, d_masks, n_masks, keys[0], outI, n_i));
        UKM_TRY(ukm_masks_gather(ctx, tmp, d_masks + n_masks + 1, n_masks, keys[0], outD, n_d));
    }
    if (ctx->stats_on && !ctx->pending.empty()) ctx->pending.back().bytes += (double)(*n_i + *n_d) * 16.0;
    return UKM_OK;
}
