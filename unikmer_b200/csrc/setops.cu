// setops.cu -- union / inter / diff / common / merge on sorted k-mer streams.
//
// Replaces the inner loops of union.go:186-208,260-305; inter.go:188-286;
// diff.go:136-146,341-515,566-594; common.go:220-283,329-354 and the heap merge
// mergeChunksFile (util-sort.go:227-606).  The reference's hash maps / two-pointer
// walks / binary heap become one kernel family: a merge-path partitioned, shared-memory
// tiled two-way set operation with single-pass output (decoupled look-back), applied in
// file order (inter, diff: exactly the reference's iteration) or as a balanced tree
// (union, common, merge).
//
// Data layout: SoA -- keys uint64[n], taxids uint32[n] (optional), counts uint32[n]
// (common only).  One CTA per tile of ~THREADS*VT merged elements: the two input slices
// are brought into shared memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier),
// every thread walks VT merged elements with the reference's three-way compare, outputs
// are compacted through shared memory and written as contiguous runs.
#include <stdlib.h>

#include "common.cuh"
#include "lca.cuh"

constexpr int NW_FANIN = 8;  // sets per pass of the N-way union (nway.cu)

namespace {

enum SetOp { OP_INTER = 0, OP_DIFF = 1, OP_UNION = 2, OP_MERGE = 3 };

constexpr int SO_THREADS = 256;
constexpr int SO_VT = 15;                    // merged elements per thread (loop runs VT+1 steps)
constexpr int SO_TILE = SO_THREADS * SO_VT;  // merged elements per tile (+-1 after pairing)

// number of A elements among the first `diag` merged elements; ties: A first
template <typename KA, typename KB>
__device__ __forceinline__ int merge_path(KA a, int na, KB b, int nb, int diag) {
    int lo = diag > nb ? diag - nb : 0;
    int hi = diag < na ? diag : na;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (a[mid] <= b[diag - 1 - mid]) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// Merge-path split on global memory.  Plain bisection costs ~30 dependent probes per tile boundary, each
// pulling its own DRAM sector (ncu: 2 GB per 1e9-element pass, 10% of all traffic).  The split of a diagonal
// is almost always close to diag * nA / (nA + nB), so: gallop out from that guess until the answer is
// bracketed, then bisect inside the bracket (probes land in a handful of neighbouring sectors).
__device__ __forceinline__ long long merge_path_global(const uint64_t* __restrict__ a, long long na, const uint64_t* __restrict__ b,
                                                       long long nb, long long diag) {
    long long lo = diag > nb ? diag - nb : 0;
    long long hi = diag < na ? diag : na;
    if (lo >= hi) return lo;
    // P(x) := a[x] > b[diag-1-x] is monotone false..true on [lo, hi); the answer is the first true (or hi)
    auto pred = [&](long long x) { return a[x] > b[diag - 1 - x]; };
    long long g = (long long)((double)diag * ((double)na / (double)(na + nb)));
    if (g < lo) g = lo;
    if (g > hi - 1) g = hi - 1;
    long long step = 16;
    if (pred(g)) {  // answer <= g: walk left
        hi = g;
        while (hi > lo) {
            long long x = hi - step < lo ? lo : hi - step;
            if (pred(x)) { hi = x; step <<= 1; }
            else { lo = x + 1; break; }
        }
    } else {  // answer > g: walk right
        lo = g + 1;
        while (lo < hi) {
            long long x = lo + step - 1 > hi - 1 ? hi - 1 : lo + step - 1;
            if (!pred(x)) { lo = x + 1; step <<= 1; }
            else { hi = x; break; }
        }
    }
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (pred(mid)) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}

// tile boundaries: part[2t] = A index, part[2t+1] = B index of the first element of tile t.
// For the pairing ops an equal (A,B) pair is never split: the B twin joins the A side's tile.
__global__ void setop_partition_kernel(const uint64_t* __restrict__ A, long long nA, const uint64_t* __restrict__ B, long long nB,
                                       int num_tiles, int tile_elems, int pairing, long long* __restrict__ part,
                                       int* __restrict__ err) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > num_tiles) return;
    long long total = nA + nB;
    long long diag = (long long)t * tile_elems;
    if (diag > total) diag = total;
    long long a = merge_path_global(A, nA, B, nB, diag);
    long long b = diag - a;
    if (pairing) {
        if (a > 0 && b < nB && A[a - 1] == B[b]) ++b;
        // contract check at the tile seams (inside tiles the walk checks it)
        if (a > 0 && a < nA && A[a - 1] >= A[a]) atomicExch(err, (int)UKM_E_NOT_SORTED_UNIQUE);
        if (b > 0 && b < nB && B[b - 1] >= B[b]) atomicExch(err, (int)UKM_E_NOT_SORTED_UNIQUE);
    }
    part[2 * t] = a;
    part[2 * t + 1] = b;
}

// One thread issues the loads of elements [start, start+count) of g into shared memory so that
// element start+i lands in s[h + i], h = misalignment of the first element in elements.  The
// 16-byte aligned body goes through one TMA bulk copy; the (< 16 B) head and tail through plain
// loads.  Returns the bytes the mbarrier has to expect.
template <typename T>
__device__ __forceinline__ int slice_offset(const T* g, long long start) {
    return (int)((reinterpret_cast<uintptr_t>(g + start) & 15u) / sizeof(T));
}
template <typename T>
__device__ __forceinline__ unsigned slice_body_bytes(const T* g, long long start, int count) {
    constexpr int PER16 = 16 / sizeof(T);
    int h = slice_offset(g, start);
    int head = h ? (PER16 - h) : 0;
    if (head > count) head = count;
    int body = ((count - head) / PER16) * PER16;
    return (unsigned)(body * sizeof(T));
}
template <typename T>
__device__ __forceinline__ void slice_issue(T* s, const T* g, long long start, int count, uint64_t* bar) {
    constexpr int PER16 = 16 / sizeof(T);
    int h = slice_offset(g, start);
    int head = h ? (PER16 - h) : 0;
    if (head > count) head = count;
    int body = ((count - head) / PER16) * PER16;
    for (int i = 0; i < head; ++i) s[h + i] = g[start + i];
    if (body) tma_load_1d(s + h + head, g + start + head, (unsigned)(body * sizeof(T)), bar);
    for (int i = head + body; i < count; ++i) s[h + i] = g[start + i];
}

// the same, split so a producer can do the plain part before arming the barrier
template <typename T>
__device__ __forceinline__ void slice_issue_plain(T* s, const T* g, long long start, int count) {
    constexpr int PER16 = 16 / sizeof(T);
    int h = slice_offset(g, start);
    int head = h ? (PER16 - h) : 0;
    if (head > count) head = count;
    int body = ((count - head) / PER16) * PER16;
    for (int i = 0; i < head; ++i) s[h + i] = g[start + i];
    for (int i = head + body; i < count; ++i) s[h + i] = g[start + i];
}
template <typename T>
__device__ __forceinline__ void slice_issue_bulk(T* s, const T* g, long long start, int count, uint64_t* bar) {
    constexpr int PER16 = 16 / sizeof(T);
    int h = slice_offset(g, start);
    int head = h ? (PER16 - h) : 0;
    if (head > count) head = count;
    int body = ((count - head) / PER16) * PER16;
    if (body) tma_load_1d(s + h + head, g + start + head, (unsigned)(body * sizeof(T)), bar);
}

struct SetopArgs {
    const uint64_t* A;
    const uint32_t* tA;  // per-code taxids, or NULL: every code carries gA (unik global taxid)
    const uint32_t* cA;
    long long nA;
    const uint64_t* B;
    const uint32_t* tB;
    const uint32_t* cB;
    long long nB;
    uint32_t gA, gB;
    const long long* part;
    uint64_t* outK;
    uint32_t* outT;
    uint32_t* outC;
    uint64_t* status;        // look-back words, one per tile
    uint32_t* tile_counter;  // dynamic tile ids
    unsigned long long* total_out;
    int num_tiles;
    unsigned flags;      // UKM_F_MIX_TAXID | UKM_F_COMPARE_TAXID
    int null_mode;       // measurement aid (UKM_SETOP_NULL=1, results NOT valid): the pipeline kernel skips search and walk
    uint32_t threshold;  // OP_UNION with counts: emit only (count & 0xffff) >= threshold (0 = all)
    TaxDev tax;
    int* err;
};

template <int OP, bool TAX, bool CNT>
__global__ void __launch_bounds__(SO_THREADS) setop_kernel(const SetopArgs p) {
    constexpr int T = SO_TILE;
    constexpr int NW = SO_THREADS / 32;
    // dynamic shared memory: keys of both slices (each with up to 1 slot of alignment slack, B on
    // an even slot), then taxids, then counts (4-element alignment slack each)
    extern __shared__ __align__(16) unsigned char so_smem[];
    uint64_t* s_k = reinterpret_cast<uint64_t*>(so_smem);                          // T + 8
    uint32_t* s_t = reinterpret_cast<uint32_t*>(s_k + T + 8);                      // T + 20 (TAX)
    uint32_t* s_c = s_t + (TAX ? T + 20 : 0);                                      // T + 20 (CNT)
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_part[SO_THREADS + 1];  // packed (a << 16 | b) thread starts
    __shared__ unsigned s_scan[NW + 2];
    __shared__ int s_tile;
    __shared__ unsigned long long s_prefix;

    const int tid = threadIdx.x;
    if (tid == 0) {
        s_tile = (int)atomicAdd(p.tile_counter, 1u);
        mbar_init(&s_bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int tile = s_tile;
    const long long a_lo = p.part[2 * tile], b_lo = p.part[2 * tile + 1];
    const int na = (int)(p.part[2 * tile + 2] - a_lo), nb = (int)(p.part[2 * tile + 3] - b_lo);

    // shared-memory placement of the two slices
    const int hA = slice_offset(p.A, a_lo);
    const int offB = ((hA + na + 1) & ~1) + slice_offset(p.B, b_lo);
    int hAt = 0, offBt = 0, hAc = 0, offBc = 0;
    if (TAX) {
        hAt = p.tA ? slice_offset(p.tA, a_lo) : 0;
        offBt = ((hAt + na + 3) & ~3) + (p.tB ? slice_offset(p.tB, b_lo) : 0);
    }
    if (CNT) {
        hAc = p.cA ? slice_offset(p.cA, a_lo) : 0;
        offBc = ((hAc + na + 3) & ~3) + (p.cB ? slice_offset(p.cB, b_lo) : 0);
    }
    if (tid == 0) {
        unsigned bytes = slice_body_bytes(p.A, a_lo, na) + slice_body_bytes(p.B, b_lo, nb);
        if (TAX) {
            if (p.tA) bytes += slice_body_bytes(p.tA, a_lo, na);
            if (p.tB) bytes += slice_body_bytes(p.tB, b_lo, nb);
        }
        if (CNT) {
            if (p.cA) bytes += slice_body_bytes(p.cA, a_lo, na);
            if (p.cB) bytes += slice_body_bytes(p.cB, b_lo, nb);
        }
        mbar_expect_tx(&s_bar, bytes);
        slice_issue(s_k, p.A, a_lo, na, &s_bar);
        slice_issue(s_k + offB - slice_offset(p.B, b_lo), p.B, b_lo, nb, &s_bar);
        if (TAX) {
            if (p.tA) slice_issue(s_t, p.tA, a_lo, na, &s_bar);
            if (p.tB) slice_issue(s_t + offBt - slice_offset(p.tB, b_lo), p.tB, b_lo, nb, &s_bar);
        }
        if (CNT) {
            if (p.cA) slice_issue(s_c, p.cA, a_lo, na, &s_bar);
            if (p.cB) slice_issue(s_c + offBc - slice_offset(p.cB, b_lo), p.cB, b_lo, nb, &s_bar);
        }
    }
    if (!mbar_wait(&s_bar, 0)) {
        if (tid == 0) atomicExch(p.err, (int)UKM_E_INTERNAL);
    }
    __syncthreads();  // also publishes thread 0's plain head/tail stores

    const uint64_t* sA = s_k + hA;
    const uint64_t* sB = s_k + offB;
    const uint32_t* sTA = s_t + hAt;
    const uint32_t* sTB = s_t + offBt;
    const uint32_t* sCA = s_c + hAc;
    const uint32_t* sCB = s_c + offBc;

    // per-thread start on the merge path of the tile
    const int total = na + nb;
    {
        int diag = tid * SO_VT;
        if (diag > total) diag = total;
        int a = merge_path(sA, na, sB, nb, diag);
        int b = diag - a;
        if (OP != OP_MERGE) {
            if (a > 0 && b < nb && sA[a - 1] == sB[b]) ++b;
        }
        s_part[tid] = (a << 16) | b;
        if (tid == 0) s_part[SO_THREADS] = (na << 16) | nb;
    }
    __syncthreads();
    int ai = s_part[tid] >> 16, bi = s_part[tid] & 0xffff;
    const int a1 = s_part[tid + 1] >> 16, b1 = s_part[tid + 1] & 0xffff;

#define TXA(i) (p.tA ? sTA[i] : p.gA)
#define TXB(i) (p.tB ? sTB[i] : p.gB)
    // the reference's three-way compare walk (inter.go:228-257, diff.go:395-431), VT+1 steps
    uint64_t outk[SO_VT + 1];
    uint32_t outt[TAX ? SO_VT + 1 : 1];
    uint32_t outc[CNT ? SO_VT + 1 : 1];
    unsigned emitmask = 0;
    bool bad = false;
    if (OP != OP_MERGE) {
        if (ai > 0 && ai < na && sA[ai - 1] >= sA[ai]) bad = true;
        if (bi > 0 && bi < nb && sB[bi - 1] >= sB[bi]) bad = true;
    }
    uint64_t ka = ai < a1 ? sA[ai] : 0ull, kb = bi < b1 ? sB[bi] : 0ull;
#pragma unroll
    for (int it = 0; it <= SO_VT; ++it) {
        const bool va = ai < a1, vb = bi < b1;
        const bool eq = va && vb && ka == kb;
        const bool takeA = va && (!vb || ka <= kb);
        const bool takeB = vb && !takeA;
        bool emit;
        uint64_t k = takeA ? ka : kb;
        uint32_t tx = 0, cn = 0;
        if (OP == OP_INTER) {
            emit = eq;
            if (TAX && eq) {
                uint32_t qa = TXA(ai), qb = TXB(bi);
                if (p.flags & UKM_F_MIX_TAXID) tx = qa == 0 ? qb : (qb == 0 ? qa : lca_dev(p.tax, qa, qb));
                else tx = lca_dev(p.tax, qa, qb);
            }
        } else if (OP == OP_DIFF) {
            emit = takeA && !eq;
            if (TAX && takeA) {
                tx = TXA(ai);
                if (eq && (p.flags & UKM_F_COMPARE_TAXID)) {
                    uint32_t qb = TXB(bi);  // keep: same taxid, or subject taxid below the query's
                    if (tx == qb || lca_dev(p.tax, qb, tx) == tx) emit = true;
                }
            }
        } else if (OP == OP_UNION) {
            emit = takeA || takeB;
            if (TAX && emit) {
                if (eq) tx = lca_dev(p.tax, TXA(ai), TXB(bi));
                else tx = takeA ? TXA(ai) : TXB(bi);
            }
            if (CNT && emit) {
                uint32_t ca = takeA ? (p.cA ? sCA[ai] : 1u) : 0u;
                uint32_t cb = (takeB || eq) ? (p.cB ? sCB[bi] : 1u) : 0u;
                cn = ca + cb;
                if (p.threshold && (cn & 0xffffu) < p.threshold) emit = false;
            }
        } else {  // OP_MERGE: keep everything, A first on ties
            emit = takeA || takeB;
            if (TAX && emit) tx = takeA ? TXA(ai) : TXB(bi);
        }
        outk[it] = k;
        if (TAX) outt[it] = tx;
        if (CNT) outc[it] = cn;
        if (emit) emitmask |= 1u << it;
        // advance
        const bool advB = takeB || (eq && OP != OP_MERGE);
        if (takeA) {
            ++ai;
            uint64_t nk = ai < a1 ? sA[ai] : 0ull;
            if (OP != OP_MERGE && ai < a1 && nk <= ka) bad = true;
            ka = nk;
        }
        if (advB) {
            ++bi;
            uint64_t nk = bi < b1 ? sB[bi] : 0ull;
            if (OP != OP_MERGE && bi < b1 && nk <= kb) bad = true;
            kb = nk;
        }
    }
#undef TXA
#undef TXB
    if (bad) atomicExch(p.err, (int)UKM_E_NOT_SORTED_UNIQUE);

    // compact through shared memory (the input slices are dead after the barrier inside the scan)
    const unsigned cnt = (unsigned)__popc(emitmask);
    unsigned tile_total;
    const unsigned off = block_excl_scan_u32<SO_THREADS>(cnt, s_scan, &tile_total);
    {
        unsigned o = off;
#pragma unroll
        for (int it = 0; it <= SO_VT; ++it) {
            if (emitmask & (1u << it)) {
                s_k[o] = outk[it];
                if (TAX) s_t[o] = outt[it];
                if (CNT) s_c[o] = outc[it];
                ++o;
            }
        }
    }
    unsigned long long prefix;
    if (OP == OP_MERGE) prefix = (unsigned long long)(a_lo + b_lo);
    else prefix = tile_exclusive_prefix(p.status, tile, tile_total, p.err, &s_prefix);
    if (tid == 0 && tile == p.num_tiles - 1) *p.total_out = prefix + tile_total;
    __syncthreads();
    const unsigned long long base = prefix;
    for (unsigned i = tid; i < tile_total; i += SO_THREADS) {
        p.outK[base + i] = s_k[i];
        if (TAX) p.outT[base + i] = s_t[i];
        if (CNT && p.outC) p.outC[base + i] = s_c[i];
    }
}


// ---- keys-only fast path ----------------------------------------------------------------------
// Same tiling, but the walk is stripped to the bone (no taxids/counts, no per-step validation:
// UKM_F_VALIDATE checks inputs up front instead) and outputs are staged in a SEPARATE shared
// buffer so staging needs no barrier against the input slices.  Tile ids are tickets (see below).
template <int OP, int VT>
__global__ void __launch_bounds__(SO_THREADS, (VT <= 15 ? 3 : 2)) setop_fast_kernel(const SetopArgs p) {
    constexpr int T = SO_THREADS * VT;
    constexpr int NW = SO_THREADS / 32;
    extern __shared__ __align__(16) unsigned char so_smem[];
    uint64_t* s_k = reinterpret_cast<uint64_t*>(so_smem);  // T + 8: both input slices
    uint64_t* s_o = s_k + T + 8;                           // T + 8: staged outputs
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_part[SO_THREADS + 1];
    __shared__ unsigned s_scan[NW + 2];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_tile;

    const int tid = threadIdx.x;
    // Tile ids come from a ticket counter, not from blockIdx.x: the look-back below waits for the tiles BEFORE this one,
    // and only tickets guarantee that those were handed to CTAs that already run (CUDA does not promise any dispatch
    // order of blockIdx, and another context may hold part of the GPU).
    if (tid == 0) s_tile = (int)atomicAdd(p.tile_counter, 1u);
    __syncthreads();
    const int tile = s_tile;
    const long long a_lo = p.part[2 * tile], b_lo = p.part[2 * tile + 1];
    const int na = (int)(p.part[2 * tile + 2] - a_lo), nb = (int)(p.part[2 * tile + 3] - b_lo);
    const int hA = slice_offset(p.A, a_lo);
    const int hB = slice_offset(p.B, b_lo);
    const int offB = ((hA + na + 1) & ~1) + hB;
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&s_bar, slice_body_bytes(p.A, a_lo, na) + slice_body_bytes(p.B, b_lo, nb));
        slice_issue(s_k, p.A, a_lo, na, &s_bar);
        slice_issue(s_k + offB - hB, p.B, b_lo, nb, &s_bar);
    }
    __syncthreads();  // barrier object initialised + head/tail plain stores visible
    if (!mbar_wait(&s_bar, 0)) {
        if (tid == 0) atomicExch(p.err, (int)UKM_E_INTERNAL);
    }
    const uint64_t* sA = s_k + hA;
    const uint64_t* sB = s_k + offB;

    const int total = na + nb;
    {
        int diag = tid * VT;
        if (diag > total) diag = total;
        int a = merge_path(sA, na, sB, nb, diag);
        int b = diag - a;
        if (OP != OP_MERGE) {
            if (a > 0 && b < nb && sA[a - 1] == sB[b]) ++b;
        }
        s_part[tid] = (a << 16) | b;
        if (tid == 0) s_part[SO_THREADS] = (na << 16) | nb;
    }
    __syncthreads();
    int ai = s_part[tid] >> 16, bi = s_part[tid] & 0xffff;
    const int a1 = s_part[tid + 1] >> 16, b1 = s_part[tid + 1] & 0xffff;

    uint64_t outk[VT + 1];
    unsigned emitmask = 0;
    uint64_t ka = sA[ai], kb = sB[bi];  // reads past a slice end stay inside s_k and are never used
#pragma unroll
    for (int it = 0; it <= VT; ++it) {
        const bool pa = ai < a1, pb = bi < b1;
        const bool gt = ka > kb;
        const bool takeA = pa && (!pb || !gt);
        const bool takeB = pb && !takeA;
        const bool eq = takeA && pb && (ka == kb);
        bool emit;
        if (OP == OP_INTER) emit = eq;
        else if (OP == OP_DIFF) emit = takeA && !eq;
        else emit = takeA || takeB;
        outk[it] = (OP == OP_INTER || OP == OP_DIFF) ? ka : (takeA ? ka : kb);
        emitmask |= (emit ? 1u : 0u) << it;
        if (takeA) ka = sA[++ai];
        if (takeB || (eq && OP != OP_MERGE)) kb = sB[++bi];
    }

    const unsigned cnt = (unsigned)__popc(emitmask);
    unsigned tile_total;
    const unsigned off = block_excl_scan_u32<SO_THREADS>(cnt, s_scan, &tile_total);
    {
        unsigned o = off;
#pragma unroll
        for (int it = 0; it <= VT; ++it) {
            if (emitmask & (1u << it)) s_o[o++] = outk[it];
        }
    }
    unsigned long long prefix;
    if (OP == OP_MERGE) {
        prefix = (unsigned long long)(a_lo + b_lo);
        __syncthreads();
    } else {
        prefix = tile_exclusive_prefix(p.status, tile, tile_total, p.err, &s_prefix);
    }
    if (tid == 0 && tile == p.num_tiles - 1) *p.total_out = prefix + tile_total;
    uint64_t* dst = p.outK + prefix;
    for (unsigned i = tid; i < tile_total; i += SO_THREADS) dst[i] = s_o[i];
}


// ---- persistent, warp-specialised pipeline (keys-only) ----------------------------------------------
// CTAs stay resident for the whole launch and take tiles round-robin (tile = blockIdx.x + i*gridDim.x), so
// "iteration i" of the grid covers the contiguous tile block [i*G, (i+1)*G).  Warp roles:
//   warp 0  loader : waits for a free slot of the shared-memory ring, reads the tile geometry, issues the
//                    two TMA bulk loads (full[s] mbarrier).
//   warp 1  prefix : output offsets.  Instead of a chained look-back (measured: 60% of the kernel when the
//                    count is only known at the END of a tile's work), every CTA gathers the G counts of
//                    the iteration in ONE round of independent loads: prefix = P_i + sum of counts of lower
//                    CTAs, P_{i+1} = P_i + sum of all counts.  Posts it through pre[s].
//   warps 2+ consumers: wait full[s]; merge-path search; three-way walk; scan; publish the count; copy tile
//                    i-DEFER out (its prefix has had DEFER tile-times to arrive) and free its slot; stage
//                    this tile's outputs IN PLACE in its slot.
// Global-memory latency (geometry, TMA, counts of other CTAs) never stalls the merging warps.
constexpr int PIPE_MAX_GRID = 384;  // prefix warp keeps PIPE_MAX_GRID/32 counts per lane
constexpr int PIPE_AUX = 64;        // loader + prefix warps

struct PipeGeom {
    long long base;  // a_lo + b_lo (merged rank of the tile start)
    int na, nb, hA, offB;
};

template <int OP, int NT, int VT, int SLOTS, int MINB>  // NT = consumer threads
__global__ void __launch_bounds__(NT + PIPE_AUX, MINB) setop_pipe_kernel(const SetopArgs p) {
    constexpr int T = NT * VT;
    constexpr int SLOT = T + 8;
    constexpr int NW = NT / 32;
    constexpr int DEFER = SLOTS - 2;  // copy-out lag in tiles
    extern __shared__ __align__(16) unsigned char so_smem[];
    uint64_t* s_slots = reinterpret_cast<uint64_t*>(so_smem);  // SLOTS * SLOT
    __shared__ __align__(8) uint64_t full_bar[SLOTS], empty_bar[SLOTS], pre_bar[SLOTS];
    __shared__ unsigned long long s_cnt[SLOTS], s_pre[SLOTS];
    __shared__ PipeGeom s_geom[SLOTS];
    __shared__ int s_part[NT + 1];
    __shared__ unsigned s_scan[NW + 2];

    const int G = gridDim.x;
    const int n_my = (p.num_tiles - (int)blockIdx.x + G - 1) / G;  // tiles of this CTA
    if (threadIdx.x == 0) {
        for (int s = 0; s < SLOTS; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], NT);  // every consumer thread arrives after its share of the copy-out
            mbar_init(&pre_bar[s], 1);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const unsigned lane = lane_id();

    if (threadIdx.x < 32) {
        // ================= loader warp =================
        for (int li = 0; li < n_my; ++li) {
            const int s = li % SLOTS, u = li / SLOTS;
            if (u > 0 && !mbar_wait(&empty_bar[s], (unsigned)(u - 1) & 1u)) {
                if (lane == 0) atomicExch(p.err, (int)UKM_E_INTERNAL);
            }
            if (lane == 0) {
                const int tile = (int)blockIdx.x + li * G;
                const long long a_lo = p.part[2 * tile], b_lo = p.part[2 * tile + 1];
                const int na = (int)(p.part[2 * tile + 2] - a_lo), nb = (int)(p.part[2 * tile + 3] - b_lo);
                const int hA = slice_offset(p.A, a_lo), hB = slice_offset(p.B, b_lo);
                const int offB = ((hA + na + 1) & ~1) + hB;
                uint64_t* slot = s_slots + (size_t)s * SLOT;
                PipeGeom g;
                g.base = a_lo + b_lo;
                g.na = na; g.nb = nb; g.hA = hA; g.offB = offB;
                s_geom[s] = g;
                // head/tail elements by plain stores, bodies by TMA; the arrive (release) publishes both
                const unsigned bytes = slice_body_bytes(p.A, a_lo, na) + slice_body_bytes(p.B, b_lo, nb);
                slice_issue_plain(slot, p.A, a_lo, na);
                slice_issue_plain(slot + offB - hB, p.B, b_lo, nb);
                mbar_expect_tx(&full_bar[s], bytes);
                slice_issue_bulk(slot, p.A, a_lo, na, &full_bar[s]);
                slice_issue_bulk(slot + offB - hB, p.B, b_lo, nb, &full_bar[s]);
            }
            __syncwarp();
        }
        return;
    }
    if (threadIdx.x < 64) {
        // ================= prefix warp =================
        if (OP == OP_MERGE) return;  // positions are data independent
        constexpr int MAXM = PIPE_MAX_GRID / 32;
        unsigned long long P = 0;  // outputs of all earlier iterations (identical on every CTA)
        for (int bi = 0; bi < n_my; ++bi) {
            const int s = bi % SLOTS, u = bi / SLOTS;
            const int tile0 = bi * G;
            const int n_iter = (p.num_tiles - tile0) < G ? (p.num_tiles - tile0) : G;  // tiles in this iteration
            unsigned long long val[MAXM];
            unsigned have = 0;  // bit m: val[m] arrived
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
                __nanosleep(100);  // do not steal issue slots from the merging warps while other CTAs catch up
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
            (void)u;
            __syncwarp();
        }
        return;
    }

    // ================= consumers =================
    const int tid = (int)threadIdx.x - PIPE_AUX;
    for (int i = 0; i < n_my + DEFER; ++i) {
        unsigned emitmask = 0;
        uint64_t outk[VT + 1];
        unsigned off = 0;
        uint64_t* slot = nullptr;
        if (i < n_my) {
            const int s = i % SLOTS, u = i / SLOTS;
            slot = s_slots + (size_t)s * SLOT;
            if (!mbar_wait(&full_bar[s], (unsigned)u & 1u)) {
                if (tid == 0) atomicExch(p.err, (int)UKM_E_INTERNAL);
            }
            const PipeGeom g = s_geom[s];
            const int na = g.na, nb = g.nb;
            const uint64_t* sA = slot + g.hA;
            const uint64_t* sB = slot + g.offB;
            const int total = na + nb;
            if (!p.null_mode) {
            {
                int diag = tid * VT;
                if (diag > total) diag = total;
                int a = merge_path(sA, na, sB, nb, diag);
                int b = diag - a;
                if (OP != OP_MERGE) {
                    if (a > 0 && b < nb && sA[a - 1] == sB[b]) ++b;
                }
                s_part[tid] = (a << 16) | b;
                if (tid == 0) s_part[NT] = (na << 16) | nb;
            }
            named_bar_sync(1, NT);
            // the walk keeps POINTERS into the two slices (index arithmetic made a fifth of the kernel's instructions)
            const uint64_t* qa = sA + (s_part[tid] >> 16);
            const uint64_t* qb = sB + (s_part[tid] & 0xffff);
            const uint64_t* const ea = sA + (s_part[tid + 1] >> 16);
            const uint64_t* const eb = sB + (s_part[tid + 1] & 0xffff);
            uint64_t ka = *qa, kb = *qb;
#pragma unroll
            for (int it = 0; it <= VT; ++it) {
                const bool pa = qa < ea, pb = qb < eb;
                const bool gt = ka > kb;
                const bool takeA = pa && (!pb || !gt);
                const bool takeB = pb && !takeA;
                const bool eq = takeA && pb && (ka == kb);
                bool emit;
                if (OP == OP_INTER) emit = eq;
                else if (OP == OP_DIFF) emit = takeA && !eq;
                else emit = takeA || takeB;
                outk[it] = (OP == OP_INTER || OP == OP_DIFF) ? ka : (takeA ? ka : kb);
                emitmask |= (emit ? 1u : 0u) << it;
                if (takeA) ka = *++qa;
                if (takeB || (eq && OP != OP_MERGE)) kb = *++qb;
            }
            }  // !null_mode
            unsigned tile_total;
            off = group_excl_scan_u32<NT>((unsigned)__popc(emitmask), (unsigned)tid, s_scan, &tile_total, 1);
            // every consumer is past its walk (two barriers inside the scan): the slot may be overwritten
            if (tid == 0) {
                s_cnt[s] = tile_total;
                if (OP != OP_MERGE) st_relaxed_u64(&p.status[(int)blockIdx.x + i * G], UKM_LB_PARTIAL | (uint64_t)tile_total);
            }
        }
        // copy tile i-DEFER out while this tile's count travels
        if (i >= DEFER && i - DEFER < n_my) {
            const int ip = i - DEFER;
            const int sp = ip % SLOTS, up = ip / SLOTS;
            const uint64_t* prev = s_slots + (size_t)sp * SLOT;
            unsigned long long prefix;
            if (OP == OP_MERGE) {
                prefix = (unsigned long long)s_geom[sp].base;
                if (tid == 0 && (int)blockIdx.x + ip * G == p.num_tiles - 1) *p.total_out = prefix + s_cnt[sp];
            } else {
                if (!mbar_wait(&pre_bar[sp], (unsigned)up & 1u)) {
                    if (tid == 0) atomicExch(p.err, (int)UKM_E_INTERNAL);
                }
                prefix = s_pre[sp];
            }
            const unsigned n_prev = (unsigned)s_cnt[sp];
            uint64_t* dst = p.outK + prefix;
            for (unsigned j = tid; j < n_prev; j += NT) dst[j] = prev[j];
            mbar_arrive(&empty_bar[sp]);  // release: my reads of the slot are done
        }
        // stage this tile's outputs in place
        if (i < n_my) {
            unsigned o = off;
#pragma unroll
            for (int it = 0; it <= VT; ++it) {
                if (emitmask & (1u << it)) slot[o++] = outk[it];
            }
        }
        named_bar_sync(1, NT);  // staged tile (and s_cnt) visible to every consumer before a later copy-out
    }
}

// ---- search path for skewed pairs (|B| >> |A|): inter / diff -------------------------------------
// When the running set A is much smaller than the next file B (inter.go / diff.go after a few files),
// walking all of B is wasted work: every thread looks its A elements up in B's window for the tile
// (branch-free lower_bound over global memory, top levels L1/L2 resident) and A is compacted in order.
constexpr int SS_THREADS = 256;
constexpr int SS_ITEMS = 4;
constexpr int SS_TILE = SS_THREADS * SS_ITEMS;

// bpart[t] = lower_bound(B, A[t * SS_TILE]) for t < num_tiles, bpart[num_tiles] = nB
__global__ void search_partition_kernel(const uint64_t* __restrict__ A, long long nA, const uint64_t* __restrict__ B, long long nB,
                                        int num_tiles, long long* __restrict__ bpart) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > num_tiles) return;
    if (t == num_tiles) {
        bpart[t] = nB;
        return;
    }
    const uint64_t x = A[(long long)t * SS_TILE];
    long long lo = 0, hi = nB;
    while (lo < hi) {
        long long mid = lo + ((hi - lo) >> 1);
        if (B[mid] < x) lo = mid + 1;
        else hi = mid;
    }
    bpart[t] = lo;
}

template <int OP, bool TAX, int MODE>
__global__ void __launch_bounds__(SS_THREADS, TAX ? 6 : 8) setop_search_kernel(const SetopArgs p) {
    constexpr int NW = SS_THREADS / 32;
    __shared__ uint64_t s_o[SS_TILE];
    __shared__ uint32_t s_ot[TAX ? SS_TILE : 1];
    __shared__ unsigned s_scan[NW + 2];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_tile;
    const int tid = threadIdx.x;
    if (tid == 0) s_tile = (int)atomicAdd(p.tile_counter, 1u);  // ticket order: the look-back only waits for running tiles
    __syncthreads();
    const int tile = s_tile;
    const long long base = (long long)tile * SS_TILE + (long long)tid * SS_ITEMS;
    const long long blo = p.part[tile], bhi = p.part[tile + 1];
    uint64_t a[SS_ITEMS];
    long long pos[SS_ITEMS];
#pragma unroll
    for (int j = 0; j < SS_ITEMS; ++j) {
        a[j] = (base + j < p.nA) ? p.A[base + j] : ~0ull;
        pos[j] = blo;
    }
    long long n = bhi - blo;
    if (MODE != 0 && n > 64 && n < (1ll << 30)) {
        // interpolated start + gallop + bisection: the probes of one look-up stay within a few sectors of
        // the answer instead of spreading over the whole window (bisection reads ~12 sectors per look-up,
        // i.e. all of B once |A| >= |B| / 64).  Mode 2 anchors items 1.. on item 0's position.
        // Positions are 32-bit offsets into the window (larger windows take the bisection below).
        const uint64_t* __restrict__ W = p.B + blo;
        const int wn = (int)n;
        const uint64_t bl = __ldg(W), bh = __ldg(W + wn - 1);
        const float dens = (float)(wn - 1) / (float)(bh - bl + 1);
        auto guess = [&](uint64_t x) -> int {
            if (x <= bl) return 0;
            if (x > bh) return wn;
            return (int)fminf((float)(x - bl) * dens, (float)wn);
        };
        // One probe per item and round, the items of a thread in lock step (independent loads in flight).
        // Every round picks q in [lo, hi) and W[q] rules one side out.  st > 0: galloping right from the
        // guess (steps 4, 8, ..), st < 0: galloping left, st == 0: bisecting the bracket the gallop closed.
        int lo[SS_ITEMS], hi[SS_ITEMS], st[SS_ITEMS];
        auto first = [&](int j, int lo0, int g) {  // probe the guess: it decides the gallop's direction
            lo[j] = lo0; hi[j] = wn; st[j] = 0;
            if (lo0 < wn) {
                const int q = max(lo0, min(g, wn - 1));
                const bool lt = __ldg(W + q) < a[j];
                if (lt) lo[j] = q + 1;
                else hi[j] = q;
                st[j] = lt ? 4 : -4;
            }
        };
        auto rounds = [&](int j0) {
            for (;;) {
                bool busy = false;
#pragma unroll
                for (int j = j0; j < SS_ITEMS; ++j) {
                    if (lo[j] < hi[j]) {
                        int q = st[j] > 0 ? lo[j] + st[j] - 1 : (st[j] < 0 ? hi[j] + st[j] : lo[j] + ((hi[j] - lo[j]) >> 1));
                        q = max(lo[j], min(q, hi[j] - 1));
                        const bool lt = __ldg(W + q) < a[j];
                        if (lt) lo[j] = q + 1;
                        else hi[j] = q;
                        st[j] = ((st[j] > 0 && !lt) || (st[j] < 0 && lt)) ? 0 : st[j] * 2;  // |st| < 2 * wn < 2^31
                        busy = true;
                    }
                }
                if (!busy) break;
            }
        };
        if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < SS_ITEMS; ++j) first(j, 0, guess(a[j]));
            rounds(0);
        } else {
            first(0, 0, guess(a[0]));
#pragma unroll
            for (int j = 1; j < SS_ITEMS; ++j) { lo[j] = 0; hi[j] = 0; st[j] = 0; }
            rounds(0);
            // a[j] > a[0]: the answer is at or after item 0's; the ~0 padding past nA lands on the end
#pragma unroll
            for (int j = 1; j < SS_ITEMS; ++j)
                first(j, lo[0], a[j] > bh ? wn : lo[0] + (int)fminf((float)(a[j] - a[0]) * dens, (float)wn));
            rounds(1);
        }
#pragma unroll
        for (int j = 0; j < SS_ITEMS; ++j) pos[j] = blo + lo[j];
    } else {
        // branch-free lower_bound (uniform trip count: n depends only on the window size), SS_ITEMS
        // independent chains per thread.  Invariant: the answer lies in [pos, pos + n].
        while (n > 1) {
            const long long half = n >> 1;
#pragma unroll
            for (int j = 0; j < SS_ITEMS; ++j) {
                const uint64_t v = __ldg(p.B + pos[j] + half);
                if (v < a[j]) pos[j] += half;
            }
            n -= half;
        }
        if (n == 1) {
#pragma unroll
            for (int j = 0; j < SS_ITEMS; ++j)
                if (__ldg(p.B + pos[j]) < a[j]) pos[j] += 1;
        }
    }
    unsigned mask = 0;
    uint32_t tx[TAX ? SS_ITEMS : 1] = {0};
#pragma unroll
    for (int j = 0; j < SS_ITEMS; ++j) {
        bool valid = base + j < p.nA;
        bool found = valid && pos[j] < bhi && p.B[pos[j]] == a[j];
        bool emit = valid && (OP == OP_INTER ? found : !found);
        if (TAX && valid) {
            uint32_t qa = p.tA ? p.tA[base + j] : p.gA;
            if (OP == OP_INTER) {
                if (found) {
                    uint32_t qb = p.tB ? p.tB[pos[j]] : p.gB;
                    if (p.flags & UKM_F_MIX_TAXID) tx[j] = qa == 0 ? qb : (qb == 0 ? qa : lca_dev(p.tax, qa, qb));
                    else tx[j] = lca_dev(p.tax, qa, qb);
                }
            } else {
                tx[j] = qa;
                if (found && (p.flags & UKM_F_COMPARE_TAXID)) {
                    uint32_t qb = p.tB ? p.tB[pos[j]] : p.gB;
                    if (qa == qb || lca_dev(p.tax, qb, qa) == qa) emit = true;
                }
            }
        }
        if (emit) mask |= 1u << j;
    }
    unsigned tile_total;
    const unsigned off = block_excl_scan_u32<SS_THREADS>((unsigned)__popc(mask), s_scan, &tile_total);
    {
        unsigned o = off;
#pragma unroll
        for (int j = 0; j < SS_ITEMS; ++j)
            if (mask & (1u << j)) {
                s_o[o] = a[j];
                if (TAX) s_ot[o] = tx[j];
                ++o;
            }
    }
    const unsigned long long prefix = tile_exclusive_prefix(p.status, tile, tile_total, p.err, &s_prefix);
    if (tid == 0 && tile == p.num_tiles - 1) *p.total_out = prefix + tile_total;
    for (unsigned i = tid; i < tile_total; i += SS_THREADS) {
        p.outK[prefix + i] = s_o[i];
        if (TAX) p.outT[prefix + i] = s_ot[i];
    }
}

// ---- host side: one two-way pass ----------------------------------------------------------
struct DevSet {
    uint64_t* k = nullptr;
    uint32_t* t = nullptr;  // NULL with taxids on: every code carries g
    uint32_t* c = nullptr;
    size_t n = 0;
    uint32_t g = 0;
};

constexpr size_t setop_smem(bool tax, bool cnt) {
    return (size_t)(SO_TILE + 8) * 8 + (tax ? (size_t)(SO_TILE + 20) * 4 : 0) + (cnt ? (size_t)(SO_TILE + 20) * 4 : 0);
}

template <int OP, bool TAX, bool CNT>
int launch_setop_v(ukm_ctx* ctx, const SetopArgs& a) {
    constexpr size_t smem = setop_smem(TAX, CNT);
    auto kern = setop_kernel<OP, TAX, CNT>;
    UKM_TRY(ukm_kernel_config(ctx, kern, smem, 0, nullptr));
    kern<<<a.num_tiles, SO_THREADS, smem, ctx->stream>>>(a);
    UKM_LAUNCHED(ctx);
    return UKM_OK;
}

template <int OP>
int launch_setop(ukm_ctx* ctx, bool tax, bool cnt, const SetopArgs& a) {
    if (tax && cnt) return launch_setop_v<OP, true, true>(ctx, a);
    if (tax) return launch_setop_v<OP, true, false>(ctx, a);
    if (cnt) return launch_setop_v<OP, false, true>(ctx, a);
    return launch_setop_v<OP, false, false>(ctx, a);
}

const char* op_name(int op, bool tax) {
    switch (op) {
        case OP_INTER: return tax ? "setop_inter_tax" : "setop_inter";
        case OP_DIFF: return tax ? "setop_diff_tax" : "setop_diff";
        case OP_UNION: return tax ? "setop_union_tax" : "setop_union";
        default: return tax ? "setop_merge_tax" : "setop_merge";
    }
}

template <int OP, int VT>
int launch_fast_v(ukm_ctx* ctx, const SetopArgs& a) {
    constexpr size_t smem = (size_t)2 * (SO_THREADS * VT + 8) * 8;
    auto kern = setop_fast_kernel<OP, VT>;
    UKM_TRY(ukm_kernel_config(ctx, kern, smem, 0, nullptr));
    kern<<<a.num_tiles, SO_THREADS, smem, ctx->stream>>>(a);
    UKM_LAUNCHED(ctx);
    return UKM_OK;
}

template <int VT>
int launch_fast(ukm_ctx* ctx, int op, const SetopArgs& a) {
    switch (op) {
        case OP_INTER: return launch_fast_v<OP_INTER, VT>(ctx, a);
        case OP_DIFF: return launch_fast_v<OP_DIFF, VT>(ctx, a);
        case OP_UNION: return launch_fast_v<OP_UNION, VT>(ctx, a);
        default: return launch_fast_v<OP_MERGE, VT>(ctx, a);
    }
}

template <int OP, int NT, int VT, int SLOTS, int MINB>
int launch_pipe_v(ukm_ctx* ctx, SetopArgs a) {
    constexpr size_t smem = (size_t)SLOTS * (NT * VT + 8) * 8;
    auto kern = setop_pipe_kernel<OP, NT, VT, SLOTS, MINB>;
    int ctas_per_sm = 0;
    UKM_TRY(ukm_kernel_config(ctx, kern, smem, NT + PIPE_AUX, &ctas_per_sm));
    // persistent grid: every CTA must be resident (tiles are chained in index order)
    int grid = ctas_per_sm * ctx->sm_count;
    if (grid > PIPE_MAX_GRID) grid = PIPE_MAX_GRID;
    if (grid > a.num_tiles) grid = a.num_tiles;
    return ukm_launch_coop(ctx, kern, grid, NT + PIPE_AUX, smem, a);  // co-residency guaranteed by the driver
}

template <int NT, int VT, int SLOTS, int MINB>
int launch_pipe(ukm_ctx* ctx, int op, const SetopArgs& a) {
    switch (op) {
        case OP_INTER: return launch_pipe_v<OP_INTER, NT, VT, SLOTS, MINB>(ctx, a);
        case OP_DIFF: return launch_pipe_v<OP_DIFF, NT, VT, SLOTS, MINB>(ctx, a);
        case OP_UNION: return launch_pipe_v<OP_UNION, NT, VT, SLOTS, MINB>(ctx, a);
        default: return launch_pipe_v<OP_MERGE, NT, VT, SLOTS, MINB>(ctx, a);
    }
}

// Pipeline shapes: consumer threads x elements per thread x ring slots.  UKM_SETOP_PIPE picks one by
// index for A/B runs ("off" = the one-tile-per-CTA kernel).
struct PipeCfg {
    int nt, vt, slots;
};
const PipeCfg kPipeCfgs[] = {{256, 15, 3}, {256, 13, 4}, {384, 10, 3}, {384, 11, 3}, {512, 7, 3}, {256, 11, 4}};
constexpr int kNumPipeCfgs = 6;
int pipe_cfg() {  // -1 = off
    const char* e = getenv("UKM_SETOP_PIPE");
    if (!e) return 1;  // 256 x 13, 4 slots: best union time on B200 (profiles/)
    if (e[0] == 'o') return -1;
    int v = atoi(e);
    return (v >= 0 && v < kNumPipeCfgs) ? v : 0;
}

// merged elements per thread of the keys-only kernel; UKM_SETOP_VT overrides for A/B runs
int fast_vt() {
    const char* e = getenv("UKM_SETOP_VT");
    int v = e ? atoi(e) : 15;
    return (v == 11 || v == 15 || v == 19 || v == 23) ? v : 15;
}

// |B| >= SKEW * |A| => look A up in B instead of walking B (UKM_SETOP_SKEW overrides; 0 disables)
long long search_skew() {
    const char* e = getenv("UKM_SETOP_SKEW");
    return e ? atoll(e) : 6;  // C3 on B200: 12.95 ms per inter at 6, 13.2 ms at 3 and at 10..16, 15.7 ms without the look-up path
}

// How the look-up kernel finds its positions (UKM_SEARCH_MODE overrides; see setop_search_kernel): 0 bisection of
// the tile's window, 1 interpolated guess + gallop per item, 2 the same with items 1.. anchored on item 0.
// C3 on B200 (tools/exp_search.py): inter 12.89 / 14.54 / 12.67 ms, diff 13.32 / 15.01 / 13.15 ms for 0 / 1 / 2.
int search_mode() {
    const char* e = getenv("UKM_SEARCH_MODE");
    const int v = e ? atoi(e) : 2;
    return (v >= 0 && v <= 2) ? v : 2;
}

template <int OP>
void launch_search(ukm_ctx* ctx, bool tax, int mode, const SetopArgs& a) {
#define UKM_SS(T, M) setop_search_kernel<OP, T, M><<<a.num_tiles, SS_THREADS, 0, ctx->stream>>>(a)
    if (tax) {
        if (mode == 1) UKM_SS(true, 1);
        else if (mode == 2) UKM_SS(true, 2);
        else UKM_SS(true, 0);
    } else {
        if (mode == 1) UKM_SS(false, 1);
        else if (mode == 2) UKM_SS(false, 2);
        else UKM_SS(false, 0);
    }
#undef UKM_SS
}

// out buffers must hold: INTER min(nA,nB); DIFF nA; UNION/MERGE nA+nB.  *n_out gets the count
// (host value, after a stream sync).
int setop2(ukm_ctx* ctx, int op, const DevSet& A, const DevSet& B, bool tax, bool cnt, unsigned flags, uint32_t threshold,
           DevSet* out) {
    const long long nA = (long long)A.n, nB = (long long)B.n;
    const long long total = nA + nB;
    if (total == 0) {
        out->n = 0;
        return UKM_OK;
    }
    const long long skew = search_skew();
    const bool use_search = (op == OP_INTER || op == OP_DIFF) && !cnt && nA > 0 && skew > 0 && nB >= skew * nA;
    const bool use_fast = !use_search && !tax && !cnt;
    const int pc = pipe_cfg();
    const bool use_pipe = use_fast && pc >= 0;
    const int vt = fast_vt();
    const int tile_elems = use_search ? SS_TILE
                         : use_pipe ? kPipeCfgs[pc].nt * kPipeCfgs[pc].vt
                         : use_fast ? SO_THREADS * vt : SO_TILE;
    const int num_tiles = (int)(((use_search ? nA : total) + tile_elems - 1) / tile_elems);
    ukm_tmp tmp(ctx);
    long long* d_part = nullptr;
    uint64_t* d_status = nullptr;
    UKM_TRY(tmp.alloc(&d_part, (size_t)2 * (num_tiles + 1)));
    UKM_TRY(tmp.alloc(&d_status, (size_t)num_tiles + 4));
    // tile counter and output total live in the tail of the status allocation (zeroed together)
    uint32_t* d_counter = reinterpret_cast<uint32_t*>(d_status + num_tiles);
    unsigned long long* d_total = reinterpret_cast<unsigned long long*>(d_status + num_tiles + 1);
    UKM_CUDA(ctx, cudaMemsetAsync(d_status, 0, ((size_t)num_tiles + 4) * sizeof(uint64_t), ctx->stream));

    SetopArgs a;
    a.A = A.k; a.tA = A.t; a.cA = A.c; a.nA = nA; a.gA = A.g;
    a.B = B.k; a.tB = B.t; a.cB = B.c; a.nB = nB; a.gB = B.g;
    a.part = d_part;
    a.outK = out->k; a.outT = out->t; a.outC = out->c;
    a.status = d_status; a.tile_counter = d_counter; a.total_out = d_total;
    a.num_tiles = num_tiles;
    a.flags = flags;
    a.null_mode = 0;
#ifdef UKM_MEASURE  // measurement build only (make EXTRA=-DUKM_MEASURE): the null mode produces no valid result
    {
        const char* e = getenv("UKM_SETOP_NULL");
        a.null_mode = (e && e[0] == '1') ? 1 : 0;
    }
#endif
    a.threshold = threshold;
    a.tax = ukm_taxdev(ctx);
    a.err = ctx->d_err;
    {
        // algorithmic bytes: both inputs read once (+ the output, added below)
        ukm_stat_scope st(ctx, op_name(op, tax), (double)total * (tax ? 12.0 : 8.0));
        int r;
        if (use_search) {
            search_partition_kernel<<<(num_tiles + 1 + 127) / 128, 128, 0, ctx->stream>>>(A.k, nA, B.k, nB, num_tiles, d_part);
            UKM_LAUNCHED(ctx);
            if (op == OP_INTER) launch_search<OP_INTER>(ctx, tax, search_mode(), a);
            else launch_search<OP_DIFF>(ctx, tax, search_mode(), a);
            UKM_LAUNCHED(ctx);
            r = UKM_OK;
        } else {
            setop_partition_kernel<<<(num_tiles + 1 + 127) / 128, 128, 0, ctx->stream>>>(A.k, nA, B.k, nB, num_tiles, tile_elems,
                                                                                         op != OP_MERGE, d_part, ctx->d_err);
            UKM_LAUNCHED(ctx);
            if (use_pipe) {
                switch (pc) {
                    case 1: r = launch_pipe<256, 13, 4, 2>(ctx, op, a); break;
                    case 2: r = launch_pipe<384, 10, 3, 2>(ctx, op, a); break;
                    case 3: r = launch_pipe<384, 11, 3, 2>(ctx, op, a); break;
                    case 4: r = launch_pipe<512, 7, 3, 2>(ctx, op, a); break;
                    case 5: r = launch_pipe<256, 11, 4, 2>(ctx, op, a); break;
                    default: r = launch_pipe<256, 15, 3, 2>(ctx, op, a); break;
                }
            } else if (use_fast) {
                r = vt == 11 ? launch_fast<11>(ctx, op, a) : vt == 19 ? launch_fast<19>(ctx, op, a)
                  : vt == 23 ? launch_fast<23>(ctx, op, a) : launch_fast<15>(ctx, op, a);
            } else {
                switch (op) {
                    case OP_INTER: r = launch_setop<OP_INTER>(ctx, tax, false, a); break;
                    case OP_DIFF: r = launch_setop<OP_DIFF>(ctx, tax, false, a); break;
                    case OP_UNION: r = launch_setop<OP_UNION>(ctx, tax, cnt, a); break;
                    default: r = launch_setop<OP_MERGE>(ctx, tax, false, a); break;
                }
            }
        }
        UKM_TRY(r);
    }
    UKM_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch, d_total, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    UKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    out->n = (size_t)ctx->h_scratch[0];
    if (ctx->stats_on && !ctx->pending.empty()) ctx->pending.back().bytes += (double)out->n * (tax ? 12.0 : 8.0);
    return UKM_OK;
}

int alloc_set(ukm_tmp& tmp, DevSet* s, size_t cap, bool tax, bool cnt) {
    UKM_TRY(tmp.alloc(&s->k, cap + 2));
    if (tax) UKM_TRY(tmp.alloc(&s->t, cap + 4));
    if (cnt) UKM_TRY(tmp.alloc(&s->c, cap + 4));
    s->n = 0;
    return UKM_OK;
}
void free_set(ukm_tmp& tmp, DevSet* s) {
    if (s->k) tmp.free_now(s->k);
    if (s->t) tmp.free_now(s->t);
    if (s->c) tmp.free_now(s->c);
    *s = DevSet();
}

int stage_set(ukm_ctx* ctx, ukm_tmp& tmp, const ukm_span* in, bool tax, DevSet* s, bool* owned, bool validate = false) {
    ukm_dspan d;
    size_t before = tmp.ptrs.size();
    // a span without per-code taxids carries its global taxid as a scalar (never materialised)
    UKM_TRY(ukm_stage_in(ctx, tmp, in, tax && in->taxids != nullptr, &d));
    *owned = tmp.ptrs.size() != before;  // something was allocated for this span
    s->k = d.keys;
    s->t = d.taxids;
    s->c = nullptr;
    s->n = d.n;
    s->g = in->global_taxid;
    if (validate && d.n > 1) UKM_TRY(ukm_dev_check_sorted_unique(ctx, d.keys, d.n));
    return UKM_OK;
}
// free whatever stage_set allocated for this span (device spans are left alone)
void unstage_set(ukm_tmp& tmp, const ukm_span* in, DevSet* s) {
    if (in->where != UKM_DEVICE) {
        if (s->k) tmp.free_now(s->k);
        if (s->t) tmp.free_now(s->t);
    } else if (s->t && s->t != in->taxids) {
        tmp.free_now(s->t);
    }
    *s = DevSet();
}

// a set that still carries its taxid as a scalar gets a real per-code array (only needed when it is
// handed back unchanged or fed to a kernel without the scalar path)
int materialize_tax(ukm_ctx* ctx, ukm_tmp& tmp, DevSet* s) {
    if (s->t) return UKM_OK;
    UKM_TRY(tmp.alloc(&s->t, s->n + 4));
    return ukm_dev_fill_u32(ctx, s->t, s->g, s->n);
}

int check_args(ukm_ctx* ctx, const ukm_span* in, int n_in, ukm_span* out, const char* what) {
    if (!ctx) return UKM_E_ARG;
    if (!in || n_in < 1 || !out) return ukm_fail(ctx, UKM_E_ARG, "%s: need at least one input span and an output span", what);
    for (int i = 0; i < n_in; ++i)
        if (in[i].n > ((size_t)1 << 40)) return ukm_fail(ctx, UKM_E_ARG, "%s: input %d too large", what, i);
    UKM_TRY(ukm_begin_call(ctx));
    return UKM_OK;
}

int need_tax(ukm_ctx* ctx, unsigned flags, const char* what) {
    if ((flags & (UKM_F_TAXID | UKM_F_MIX_TAXID)) && !ctx->tax.parent)
        return ukm_fail(ctx, UKM_E_NO_TAXONOMY, "%s: taxids requested but no taxonomy loaded (ukm_set_taxonomy)", what);
    return UKM_OK;
}

// UKM_NWAY_FORCE=1: take the N-way filter whenever it applies, whatever the size ratios (tests)
bool nway_force() {
    const char* e = getenv("UKM_NWAY_FORCE");
    return e && e[0] == '1';
}
// The N-way inter / diff filter is correct but, measured on B200 (C3: 23.3-24.3 ms against 12.7 ms for the
// file-by-file passes), slower: with so little work per tile the kernel runs at the pace of the grid-wide
// output-offset hand-off (16.8 ms with NO work per tile, UKM_NWAY_NULL=1).  It stays opt-in
// (UKM_NWAY_FILTER=1, or UKM_NWAY_FORCE=1) until its outputs are decoupled from that hand-off.
bool nway_filter_enabled() {
    const char* e = getenv("UKM_NWAY_FILTER");
    return (e && e[0] == '1') || nway_force();
}

// running op in file order: cur = in[0]; cur = cur OP in[i]   (inter.go / diff.go iteration).
// Keys-only runs of sorted files go through the single-pass N-way filter (nway.cu) up to seven files at a
// time: the result is the same set the file-by-file iteration produces, every input is read once.
int run_chain(ukm_ctx* ctx, int op, const ukm_span* in, int n_in, unsigned flags, ukm_span* out, const char* what) {
    const bool tax = (flags & (UKM_F_TAXID | UKM_F_MIX_TAXID)) != 0;
    ukm_tmp tmp(ctx);
    DevSet cur, nxt, F;
    bool owned;
    const bool validate = (flags & UKM_F_VALIDATE) != 0;
    UKM_TRY(stage_set(ctx, tmp, &in[0], tax, &cur, &owned, validate));
    bool cur_is_input = true;  // cur aliases the staged in[0]
    const size_t cap = in[0].n;
    DevSet bufs[2];
    int which = 0;
    // keys-only runs: the single-pass filter over file-0 chunks (nfilter.cu, default) or the older rank-tiled one
    const bool nfil = ukm_nfilter_enabled() && !nway_filter_enabled();
    bool nway = !tax && (nfil || (ukm_nway_enabled() && nway_filter_enabled()));
    int i = 1;
    while (i < n_in) {
        if (op == OP_INTER) {
            // Whole-file quirks of the reference: an empty first file panics (inter.go:208), an empty later file ends the
            // loop and KEEPS the current set (inter.go:211-215, quirk B-3).  A key-range SLICE of a file (UKM_F_SHARD:
            // multi-GPU shards, streamed key ranges) can be empty when the file is not; there the plain set semantics
            // apply -- an empty input makes the intersection of that range empty -- so that the concatenation of the
            // slices' results equals the whole-file result.
            if (flags & UKM_F_SHARD) {
                if (in[i].n == 0) cur.n = 0;
            } else {
                if (cur.n == 0 && cur_is_input) return ukm_fail(ctx, UKM_E_PANIC, "%s: first input is empty (inter.go:208 panics)", what);
                if (in[i].n == 0) break;  // inter.go:211-215: flagBreak keeps the current set (quirk B-3)
            }
        }
        if (op == OP_DIFF && in[i].n == 0) { ++i; continue; }
        if (cur.n == 0) break;
        if (nway) {
            // the files the next N-way pass can take: sorted, non-empty (an empty file ends inter, diff skips it)
            int idx[NW_FANIN];
            int g = 0, j = i;
            size_t total = cur.n, largest = 0;
            while (j < n_in && g < NW_FANIN - 1) {
                if (in[j].n == 0) {
                    if (op == OP_INTER) break;
                    ++j;
                    continue;
                }
                if (op == OP_DIFF && !in[j].sorted) break;
                idx[g++] = j++;
                total += in[idx[g - 1]].n;
                if (in[idx[g - 1]].n > largest) largest = in[idx[g - 1]].n;
            }
            // worth it when file 0's share of a tile fits the hash table and it is not so sparse that looking its
            // keys up (setop_search_kernel) touches only a fraction of the other files
            // UKM_NFILTER_FORCE=1 (tests): take the chunk filter for any number of subjects and any size ratio
            const bool nf_force = nfil && getenv("UKM_NFILTER_FORCE") && getenv("UKM_NFILTER_FORCE")[0] == '1';
            if ((g >= 2 || (nf_force && g >= 1)) && ((cur.n * 3 <= total && cur.n * 64 >= largest) || nway_force() || nf_force)) {
                DevSet G[NW_FANIN];
                const uint64_t* ks[NW_FANIN];
                size_t ns[NW_FANIN];
                ks[0] = cur.k;
                ns[0] = cur.n;
                for (int q = 0; q < g; ++q) {
                    UKM_TRY(stage_set(ctx, tmp, &in[idx[q]], false, &G[q], &owned, validate));
                    ks[q + 1] = G[q].k;
                    ns[q + 1] = G[q].n;
                }
                if (bufs[which].k == nullptr) UKM_TRY(alloc_set(tmp, &bufs[which], cap, tax, false));
                nxt = bufs[which];
                bool fell_back = false;
                size_t n_o = 0;
                if (nfil) UKM_TRY(ukm_nfilter(ctx, op == OP_INTER, ks, ns, g + 1, nxt.k, &n_o, &fell_back));
                else UKM_TRY(ukm_nway_filter(ctx, op == OP_INTER, ks, ns, g + 1, nxt.k, &n_o, &fell_back));
                for (int q = 0; q < g; ++q) unstage_set(tmp, &in[idx[q]], &G[q]);
                if (!fell_back) {
                    if (cur_is_input) {
                        unstage_set(tmp, &in[0], &cur);
                        cur_is_input = false;
                    }
                    nxt.n = n_o;
                    bufs[which].n = n_o;
                    cur = nxt;
                    which ^= 1;
                    i = j;
                    continue;
                }
                nway = false;  // inputs that cannot be tiled: file by file from here on
            }
        }
        UKM_TRY(stage_set(ctx, tmp, &in[i], tax, &F, &owned, validate && !(op == OP_DIFF && !in[i].sorted)));
        if (op == OP_DIFF && !in[i].sorted) {
            // diff.go:341-367 handles an unsorted subject through a map; here: sort a private copy,
            // then drop duplicate codes (a map delete is idempotent)
            DevSet S;
            UKM_TRY(alloc_set(tmp, &S, F.n, tax, false));
            UKM_CUDA(ctx, cudaMemcpyAsync(S.k, F.k, F.n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
            if (tax && F.t) UKM_CUDA(ctx, cudaMemcpyAsync(S.t, F.t, F.n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
            else if (tax) UKM_TRY(ukm_dev_fill_u32(ctx, S.t, F.g, F.n));
            unstage_set(tmp, &in[i], &F);
            UKM_TRY(ukm_dev_sort(ctx, S.k, tax ? S.t : nullptr, S.n = in[i].n, 64));
            DevSet U;
            UKM_TRY(alloc_set(tmp, &U, S.n, tax, false));
            size_t nu = 0;
            UKM_TRY(ukm_dev_fold(ctx, UKM_FOLD_UNIQUE, S.k, S.t, S.n, tax, U.k, U.t, &nu));
            // taxid of a deduplicated subject code: LCA over its occurrences
            U.n = nu;
            free_set(tmp, &S);
            F = U;
            owned = true;
            if (bufs[which].k == nullptr) UKM_TRY(alloc_set(tmp, &bufs[which], cap, tax, false));
            nxt = bufs[which];
            UKM_TRY(setop2(ctx, op, cur, F, tax, false, flags, 0, &nxt));
            free_set(tmp, &F);
        } else {
            if (bufs[which].k == nullptr) UKM_TRY(alloc_set(tmp, &bufs[which], cap, tax, false));
            nxt = bufs[which];
            UKM_TRY(setop2(ctx, op, cur, F, tax, false, flags, 0, &nxt));
            unstage_set(tmp, &in[i], &F);
        }
        if (cur_is_input) {
            unstage_set(tmp, &in[0], &cur);
            cur_is_input = false;
        }
        bufs[which].n = nxt.n;
        cur = nxt;
        which ^= 1;
        ++i;
    }
    UKM_TRY(ukm_check_dev_error(ctx, what));
    if (tax) UKM_TRY(materialize_tax(ctx, tmp, &cur));
    return ukm_deliver(ctx, cur.k, tax ? cur.t : nullptr, cur.n, out);
}

// balanced tree of two-way passes (union, common, merge)
int run_tree(ukm_ctx* ctx, int op, const ukm_span* in, int n_in, unsigned flags, bool cnt, uint32_t threshold, int fold_mode,
             ukm_span* out, const char* what) {
    const bool tax = (flags & UKM_F_TAXID) != 0;
    ukm_tmp tmp(ctx);
    std::vector<DevSet> level(n_in);
    std::vector<int> src(n_in);  // index into `in` while the set is still a staged input, else -1
    for (int i = 0; i < n_in; ++i) {
        bool owned;
        UKM_TRY(stage_set(ctx, tmp, &in[i], tax, &level[i], &owned, (flags & UKM_F_VALIDATE) != 0 && op != OP_MERGE));
        src[i] = i;
    }
    // a single input still goes through one pass against an empty set so thresholds apply
    if (level.size() == 1 && cnt) {
        level.push_back(DevSet());
        src.push_back(-2);
    }
    // keys-only unions go through the single-pass N-way kernel (nway.cu), up to 8 sets per pass
    bool nway = op == OP_UNION && !tax && !cnt && ukm_nway_enabled();
    while (level.size() > 1) {
        std::vector<DevSet> next;
        std::vector<int> nsrc;
        const size_t fan = nway ? (size_t)NW_FANIN : 2;
        const bool last = level.size() <= fan;
        for (size_t i = 0; i < level.size(); i += fan) {
            const size_t gsz = level.size() - i < fan ? level.size() - i : fan;
            if (gsz == 1) {
                next.push_back(level[i]);
                nsrc.push_back(src[i]);
                continue;
            }
            DevSet o;
            size_t bound = 0;
            for (size_t j = i; j < i + gsz; ++j) bound += level[j].n;
            bool direct = false;
            if (last && fold_mode == UKM_FOLD_PLAIN && out->where == UKM_DEVICE && out->cap >= bound && out->keys &&
                (!tax || out->taxids)) {
                // final pass writes straight into the caller's device buffers (no extra copy)
                o.k = out->keys;
                o.t = tax ? out->taxids : nullptr;
                if (cnt) UKM_TRY(tmp.alloc(&o.c, bound + 4));
                direct = true;
            } else {
                UKM_TRY(alloc_set(tmp, &o, bound, tax, cnt));
            }
            // two sets: the two-way pipeline is the faster kernel (4.1 vs 5.1 ms per 1e9 keys on B200)
            if (nway && (gsz > 2 || nway_force())) {
                const uint64_t* ks[NW_FANIN];
                size_t ns[NW_FANIN];
                for (size_t j = 0; j < gsz; ++j) {
                    ks[j] = level[i + j].k;
                    ns[j] = level[i + j].n;
                }
                bool fell_back = false;
                size_t n_o = 0;
                if (ukm_nunion_enabled()) UKM_TRY(ukm_nunion(ctx, ks, ns, (int)gsz, o.k, &n_o, &fell_back));
                else UKM_TRY(ukm_nway_union(ctx, ks, ns, (int)gsz, o.k, &n_o, &fell_back));
                if (fell_back) {
                    // inputs that cannot be tiled (not duplicate-free): the rest of the tree runs two-way passes
                    if (!direct) free_set(tmp, &o);
                    for (size_t j = i; j < level.size(); ++j) {
                        next.push_back(level[j]);
                        nsrc.push_back(src[j]);
                    }
                    nway = false;
                    break;
                }
                o.n = n_o;
            } else {
                UKM_TRY(setop2(ctx, op, level[i], level[i + 1], tax, cnt, flags, last ? threshold : 0, &o));
            }
            for (size_t j = i; j < i + gsz; ++j) {
                if (src[j] >= 0) unstage_set(tmp, &in[src[j]], &level[j]);
                else if (src[j] == -1) free_set(tmp, &level[j]);
            }
            next.push_back(o);
            nsrc.push_back(-1);
        }
        level.swap(next);
        src.swap(nsrc);
    }
    UKM_TRY(ukm_check_dev_error(ctx, what));
    DevSet r = level[0];
    if (tax) UKM_TRY(materialize_tax(ctx, tmp, &r));
    if (fold_mode != UKM_FOLD_PLAIN) {
        DevSet f;
        UKM_TRY(alloc_set(tmp, &f, r.n + 2, tax, false));
        size_t nf = 0;
        UKM_TRY(ukm_dev_fold(ctx, fold_mode, r.k, r.t, r.n, tax, f.k, f.t, &nf));
        f.n = nf;
        r = f;
    }
    return ukm_deliver(ctx, r.k, tax ? r.t : nullptr, r.n, out);
}


// inter AND diff of the same sorted device-resident spans in ONE pass (nfilter.cu, NFOP_BOTH): after the first subject
// every key of file 0 is a candidate of exactly one of the two results, so both come out of one read of the inputs.
// `shard`: the spans are key-range slices (an empty slice empties that range's intersection and subtracts nothing).
// *done = false: the fused pass does not apply (more than eight files, an empty WHOLE file -- the whole-file rules of
// inter differ --, file 0 far sparser than the subjects); the caller runs the two operations one after the other.
int fused_inter_diff(ukm_ctx* ctx, const ukm_span* in, int n_in, unsigned flags, bool shard, ukm_span* out_i, ukm_span* out_d,
                     bool* done) {
    *done = false;
    if (!ukm_nfilter_enabled() || n_in < 2 || n_in > NW_FANIN) return UKM_OK;
    for (int f = 0; f < n_in; ++f) {
        if (in[f].where != UKM_DEVICE || !in[f].sorted) return UKM_OK;
        if (!shard && in[f].n == 0) return UKM_OK;
    }
    const uint64_t* ks[NW_FANIN];
    size_t ns[NW_FANIN];
    for (int f = 0; f < n_in; ++f) {
        ks[f] = in[f].keys;
        ns[f] = in[f].n;
        if ((flags & UKM_F_VALIDATE) && in[f].n > 1) UKM_TRY(ukm_dev_check_sorted_unique(ctx, in[f].keys, in[f].n));
    }
    ukm_tmp tmp(ctx);
    const size_t cap = in[0].n;
    uint64_t *d_i = nullptr, *d_d = nullptr;
    const bool direct_i = out_i->where == UKM_DEVICE && out_i->cap >= cap && out_i->keys;
    const bool direct_d = out_d->where == UKM_DEVICE && out_d->cap >= cap && out_d->keys;
    if (direct_i) d_i = out_i->keys;
    else UKM_TRY(tmp.alloc(&d_i, cap + 2));
    if (direct_d) d_d = out_d->keys;
    else UKM_TRY(tmp.alloc(&d_d, cap + 2));
    size_t n_i = 0, n_d = 0;
    bool declined = false;
    UKM_TRY(ukm_nfilter_both(ctx, ks, ns, n_in, d_i, &n_i, d_d, &n_d, &declined));
    if (declined) return UKM_OK;
    UKM_TRY(ukm_check_dev_error(ctx, "ukm_setops_stream(inter+diff)"));
    UKM_TRY(ukm_deliver(ctx, d_i, nullptr, n_i, out_i));
    UKM_TRY(ukm_deliver(ctx, d_d, nullptr, n_d, out_d));
    *done = true;
    return UKM_OK;
}

// union, inter AND diff of the same sorted device-resident spans from one pass (nway.cu: the union kernel's last merge
// level sees how many files hold every key).  Same applicability rules as fused_inter_diff, 3..8 files.
int fused_union_inter_diff(ukm_ctx* ctx, const ukm_span* in, int n_in, unsigned flags, bool shard, ukm_span* out_u, ukm_span* out_i,
                           ukm_span* out_d, bool* done) {
    *done = false;
    const char* e = getenv("UKM_FUSE3");
    if ((e && e[0] == '0') || !ukm_nway_enabled() || n_in < 3 || n_in > NW_FANIN) return UKM_OK;
    size_t total = 0;
    for (int f = 0; f < n_in; ++f) {
        if (in[f].where != UKM_DEVICE || !in[f].sorted) return UKM_OK;
        if (!shard && in[f].n == 0) return UKM_OK;
        total += in[f].n;
    }
    if (in[0].n == 0) return UKM_OK;
    const uint64_t* ks[NW_FANIN];
    size_t ns[NW_FANIN];
    for (int f = 0; f < n_in; ++f) {
        ks[f] = in[f].keys;
        ns[f] = in[f].n;
        if ((flags & UKM_F_VALIDATE) && in[f].n > 1) UKM_TRY(ukm_dev_check_sorted_unique(ctx, in[f].keys, in[f].n));
    }
    ukm_tmp tmp(ctx);
    uint64_t *d_u = nullptr, *d_i = nullptr, *d_d = nullptr;
    if (out_u->where == UKM_DEVICE && out_u->cap >= total && out_u->keys) d_u = out_u->keys;
    else UKM_TRY(tmp.alloc(&d_u, total + 2));
    if (out_i->where == UKM_DEVICE && out_i->cap >= in[0].n && out_i->keys) d_i = out_i->keys;
    else UKM_TRY(tmp.alloc(&d_i, in[0].n + 2));
    if (out_d->where == UKM_DEVICE && out_d->cap >= in[0].n && out_d->keys) d_d = out_d->keys;
    else UKM_TRY(tmp.alloc(&d_d, in[0].n + 2));
    size_t n_u = 0, n_i = 0, n_d = 0;
    bool fell_back = false;
    UKM_TRY(ukm_nway_union3(ctx, ks, ns, n_in, d_u, &n_u, d_i, &n_i, d_d, &n_d, &fell_back));
    if (fell_back) return UKM_OK;
    UKM_TRY(ukm_check_dev_error(ctx, "ukm_setops_stream(union+inter+diff)"));
    UKM_TRY(ukm_deliver(ctx, d_u, nullptr, n_u, out_u));
    UKM_TRY(ukm_deliver(ctx, d_i, nullptr, n_i, out_i));
    UKM_TRY(ukm_deliver(ctx, d_d, nullptr, n_d, out_d));
    *done = true;
    return UKM_OK;
}

// ---------------------------------------------------------------------------------------------------
// streamed operations on HOST inputs: every input byte crosses PCIe once, uploads / kernels / downloads overlap
// ---------------------------------------------------------------------------------------------------
// All set operations are key-local, so a step over host-resident files can run as a stream of key ranges: the slices
// of range c + 1 are uploaded (copy-in stream) while the requested operations run on range c (context stream) and the
// results of range c - 1 travel back (copy-out stream).  Ranges are quantiles of the largest input; the slices of a
// range are found by binary search on the host arrays.  Device memory: two sets of slice buffers + two sets of result
// buffers, sized for the largest range.
constexpr size_t STREAM_CHUNK_BYTES = (size_t)2 << 30;  // input bytes per key range (UKM_STREAM_CHUNK_MB overrides)

struct StreamPlan {
    int K = 1;                               // key ranges
    std::vector<std::vector<size_t>> off;    // off[f][c] .. off[f][c+1] = slice of file f in range c
};

void stream_plan(const ukm_span* in, int n_in, size_t chunk_bytes, StreamPlan* pl) {
    size_t total = 0, big = 0;
    int bigf = 0;
    for (int f = 0; f < n_in; ++f) {
        total += in[f].n;
        if (in[f].n > big) { big = in[f].n; bigf = f; }
    }
    size_t K = (total * 8 + chunk_bytes - 1) / chunk_bytes;
    if (K < 1) K = 1;
    if (K > big && big > 0) K = big;
    if (K > 4096) K = 4096;
    pl->K = (int)K;
    pl->off.assign(n_in, std::vector<size_t>(K + 1, 0));
    for (int f = 0; f < n_in; ++f) pl->off[f][K] = in[f].n;
    for (size_t c = 1; c < K; ++c) {
        const uint64_t cut = in[bigf].keys[big * c / K];  // range c starts at this key
        for (int f = 0; f < n_in; ++f) {
            const uint64_t* b = in[f].keys;
            size_t lo = pl->off[f][c - 1], hi = in[f].n;  // lower_bound(cut), monotone in c
            while (lo < hi) {
                const size_t mid = lo + ((hi - lo) >> 1);
                if (b[mid] < cut) lo = mid + 1;
                else hi = mid;
            }
            pl->off[f][c] = lo;
        }
    }
}

int stream_run(ukm_ctx* ctx, const ukm_span* in, int n_in, const int* ops, int n_ops, unsigned flags, ukm_span* outs) {
    size_t chunk_bytes = STREAM_CHUNK_BYTES;
    if (const char* e = getenv("UKM_STREAM_CHUNK_MB")) {
        const long v = atol(e);
        if (v > 0) chunk_bytes = (size_t)v << 20;
    }
    StreamPlan pl;
    stream_plan(in, n_in, chunk_bytes, &pl);
    const int K = pl.K;
    // the reference's whole-file rules, applied once on the file sizes (the per-range calls run with UKM_F_SHARD)
    int n_inter = n_in;  // inter.go:211-215: the loop ends at the first empty later file and keeps the current set (B-3)
    for (int f = 1; f < n_in; ++f)
        if (in[f].n == 0) { n_inter = f; break; }
    for (int k = 0; k < n_ops; ++k)
        if (ops[k] == UKM_OP_INTER && in[0].n == 0)
            return ukm_fail(ctx, UKM_E_PANIC, "ukm_setops_stream: first input is empty (inter.go:208 panics)");
    int k_inter = -1, k_diff = -1, k_union = -1;  // the first of each kind: candidates for the fused passes
    for (int k = n_ops - 1; k >= 0; --k) {
        if (ops[k] == UKM_OP_INTER) k_inter = k;
        if (ops[k] == UKM_OP_DIFF) k_diff = k;
        if (ops[k] == UKM_OP_UNION) k_union = k;
    }
    // buffers: the largest range decides
    size_t max_in = 0, max_f0 = 0;
    for (int c = 0; c < K; ++c) {
        size_t sum = 0;
        for (int f = 0; f < n_in; ++f) sum += pl.off[f][c + 1] - pl.off[f][c] + 2;  // + 2: 16-byte aligned slice starts
        if (sum > max_in) max_in = sum;
        const size_t f0 = pl.off[0][c + 1] - pl.off[0][c];
        if (f0 > max_f0) max_f0 = f0;
    }
    if (!ctx->copy_in) {
        UKM_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
        UKM_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
        for (int q = 0; q < 2; ++q) {
            UKM_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_in[q], cudaEventDisableTiming));
            UKM_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_out[q], cudaEventDisableTiming));
        }
        UKM_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_misc, cudaEventDisableTiming));
    }
    ukm_tmp tmp(ctx);
    uint64_t* d_in[2] = {nullptr, nullptr};
    std::vector<uint64_t*> d_out[2];
    std::vector<size_t> cap_out(n_ops);
    for (int k = 0; k < n_ops; ++k) cap_out[k] = (ops[k] == UKM_OP_UNION ? max_in : max_f0) + 2;
    for (int q = 0; q < 2; ++q) {
        UKM_TRY(tmp.alloc(&d_in[q], max_in + 2));
        d_out[q].resize(n_ops);
        for (int k = 0; k < n_ops; ++k) UKM_TRY(tmp.alloc(&d_out[q][k], cap_out[k]));
    }
    // the allocations are ordered on the context stream: the copy streams start after them
    UKM_CUDA(ctx, cudaEventRecord(ctx->ev_misc, ctx->stream));
    UKM_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_in, ctx->ev_misc, 0));
    UKM_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_out, ctx->ev_misc, 0));
    std::vector<ukm_span> dsp(n_in);
    std::vector<size_t> written(n_ops, 0);
    auto upload = [&](int c) -> int {
        const int q = c & 1;
        size_t pos = 0;
        for (int f = 0; f < n_in; ++f) {
            const size_t lo = pl.off[f][c], n = pl.off[f][c + 1] - lo;
            if (n) UKM_CUDA(ctx, cudaMemcpyAsync(d_in[q] + pos, in[f].keys + lo, n * 8, cudaMemcpyHostToDevice, ctx->copy_in));
            pos += (n + 1) & ~(size_t)1;
        }
        UKM_CUDA(ctx, cudaEventRecord(ctx->ev_in[q], ctx->copy_in));
        return UKM_OK;
    };
    int status = UKM_OK;
    UKM_TRY(upload(0));
    for (int c = 0; c < K && status == UKM_OK; ++c) {
        const int q = c & 1;
        // range c + 1 goes up while range c is computed (its buffer set was last read by range c - 1, which is complete:
        // every operation below ends with a synchronisation of the context stream)
        if (c + 1 < K) UKM_TRY(upload(c + 1));
        UKM_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_in[q], 0));
        if (c >= 2) UKM_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_out[q], 0));  // result set q: its download has left
        size_t pos = 0;
        for (int f = 0; f < n_in; ++f) {
            const size_t n = pl.off[f][c + 1] - pl.off[f][c];
            dsp[f] = in[f];
            dsp[f].keys = d_in[q] + pos;
            dsp[f].taxids = nullptr;
            dsp[f].n = dsp[f].cap = n;
            dsp[f].where = UKM_DEVICE;
            pos += (n + 1) & ~(size_t)1;
        }
        size_t n_res[8] = {0};
        bool have[8] = {false};
        if (k_inter >= 0 && k_diff >= 0 && k_union >= 0 && n_inter == n_in) {
            // all three results of the range from one pass over its slices
            ukm_span ou, oi, od;
            memset(&ou, 0, sizeof ou);
            memset(&oi, 0, sizeof oi);
            memset(&od, 0, sizeof od);
            ou.keys = d_out[q][k_union]; ou.cap = cap_out[k_union]; ou.where = UKM_DEVICE;
            oi.keys = d_out[q][k_inter]; oi.cap = cap_out[k_inter]; oi.where = UKM_DEVICE;
            od.keys = d_out[q][k_diff]; od.cap = cap_out[k_diff]; od.where = UKM_DEVICE;
            bool done = false;
            status = fused_union_inter_diff(ctx, dsp.data(), n_in, flags, true, &ou, &oi, &od, &done);
            if (status == UKM_OK && done) {
                n_res[k_union] = ou.n;
                n_res[k_inter] = oi.n;
                n_res[k_diff] = od.n;
                have[k_union] = have[k_inter] = have[k_diff] = true;
            }
        }
        if (status == UKM_OK && !have[k_inter >= 0 ? k_inter : 0] && k_inter >= 0 && k_diff >= 0 && n_inter == n_in) {
            // inter and diff of the range in one pass over its slices
            ukm_span oi, od;
            memset(&oi, 0, sizeof oi);
            memset(&od, 0, sizeof od);
            oi.keys = d_out[q][k_inter]; oi.cap = cap_out[k_inter]; oi.where = UKM_DEVICE;
            od.keys = d_out[q][k_diff]; od.cap = cap_out[k_diff]; od.where = UKM_DEVICE;
            bool done = false;
            status = fused_inter_diff(ctx, dsp.data(), n_in, flags, true, &oi, &od, &done);
            if (status == UKM_OK && done) {
                n_res[k_inter] = oi.n;
                n_res[k_diff] = od.n;
                have[k_inter] = have[k_diff] = true;
            }
        }
        for (int k = 0; k < n_ops && status == UKM_OK; ++k) {
            if (have[k]) continue;
            ukm_span o;
            memset(&o, 0, sizeof o);
            o.keys = d_out[q][k];
            o.cap = cap_out[k];
            o.where = UKM_DEVICE;
            if (ops[k] == UKM_OP_INTER) status = run_chain(ctx, OP_INTER, dsp.data(), n_inter, (flags & UKM_F_VALIDATE) | UKM_F_SHARD, &o, "ukm_setops_stream(inter)");
            else if (ops[k] == UKM_OP_DIFF) status = run_chain(ctx, OP_DIFF, dsp.data(), n_in, flags & UKM_F_VALIDATE, &o, "ukm_setops_stream(diff)");
            else status = run_tree(ctx, OP_UNION, dsp.data(), n_in, flags & UKM_F_VALIDATE, false, 0, UKM_FOLD_PLAIN, &o, "ukm_setops_stream(union)");
            n_res[k] = o.n;
        }
        if (status != UKM_OK) break;
        // results of range c travel back while range c + 1 is computed
        UKM_CUDA(ctx, cudaEventRecord(ctx->ev_misc, ctx->stream));
        UKM_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_out, ctx->ev_misc, 0));
        for (int k = 0; k < n_ops; ++k) {
            if (written[k] + n_res[k] > outs[k].cap) {
                outs[k].n = written[k] + n_res[k];
                status = ukm_fail(ctx, UKM_E_CAPACITY, "ukm_setops_stream: output %d needs more than %zu elements", k, outs[k].cap);
                break;
            }
            if (n_res[k])
                UKM_CUDA(ctx, cudaMemcpyAsync(outs[k].keys + written[k], d_out[q][k], n_res[k] * 8,
                                              outs[k].where == UKM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->copy_out));
            written[k] += n_res[k];
        }
        UKM_CUDA(ctx, cudaEventRecord(ctx->ev_out[q], ctx->copy_out));
    }
    // drain both copy streams before the buffers go back to the pool (also on errors)
    cudaStreamSynchronize(ctx->copy_in);
    cudaStreamSynchronize(ctx->copy_out);
    if (status != UKM_OK) return status;
    for (int k = 0; k < n_ops; ++k) outs[k].n = written[k];
    return UKM_OK;
}

// can this call take the streamed path: keys only, every input in host memory, sorted subjects
bool stream_eligible(const ukm_span* in, int n_in, unsigned flags) {
    if (flags & (UKM_F_TAXID | UKM_F_MIX_TAXID | UKM_F_COMPARE_TAXID)) return false;
    size_t total = 0;
    for (int f = 0; f < n_in; ++f) {
        if (in[f].where == UKM_DEVICE || !in[f].sorted) return false;
        if (in[f].n && !in[f].keys) return false;
        total += in[f].n;
    }
    const char* e = getenv("UKM_STREAM");
    if (e && e[0] == '0') return false;
    size_t min_bytes = (size_t)64 << 20;  // small inputs: one upload is as fast and saves the bookkeeping
    if (const char* m = getenv("UKM_STREAM_MIN_MB")) min_bytes = (size_t)atol(m) << 20;  // tests
    return total * 8 >= min_bytes;
}

}  // namespace

extern "C" int ukm_inter(ukm_ctx* ctx, const ukm_span* in, int n_in, unsigned flags, ukm_span* out) {
    UKM_TRY(check_args(ctx, in, n_in, out, "ukm_inter"));
    UKM_TRY(need_tax(ctx, flags & ~UKM_F_MIX_TAXID, "ukm_inter"));
    if (!(flags & UKM_F_SHARD) && stream_eligible(in, n_in, flags)) {
        const int op = UKM_OP_INTER;
        return stream_run(ctx, in, n_in, &op, 1, flags, out);
    }
    return run_chain(ctx, OP_INTER, in, n_in, flags, out, "ukm_inter");
}

extern "C" int ukm_diff(ukm_ctx* ctx, const ukm_span* in, int n_in, unsigned flags, ukm_span* out) {
    UKM_TRY(check_args(ctx, in, n_in, out, "ukm_diff"));
    if ((flags & UKM_F_COMPARE_TAXID) && (flags & UKM_F_TAXID)) UKM_TRY(need_tax(ctx, flags, "ukm_diff"));
    if (stream_eligible(in, n_in, flags)) {
        const int op = UKM_OP_DIFF;
        return stream_run(ctx, in, n_in, &op, 1, flags, out);
    }
    return run_chain(ctx, OP_DIFF, in, n_in, flags & ~UKM_F_MIX_TAXID, out, "ukm_diff");
}

extern "C" int ukm_union(ukm_ctx* ctx, const ukm_span* in, int n_in, unsigned flags, ukm_span* out) {
    UKM_TRY(check_args(ctx, in, n_in, out, "ukm_union"));
    UKM_TRY(need_tax(ctx, flags & UKM_F_TAXID, "ukm_union"));
    if (stream_eligible(in, n_in, flags)) {
        const int op = UKM_OP_UNION;
        return stream_run(ctx, in, n_in, &op, 1, flags, out);
    }
    return run_tree(ctx, OP_UNION, in, n_in, flags & (UKM_F_TAXID | UKM_F_VALIDATE), false, 0, UKM_FOLD_PLAIN, out, "ukm_union");
}

extern "C" int ukm_common(ukm_ctx* ctx, const ukm_span* in, int n_in, unsigned flags, uint16_t threshold, ukm_span* out) {
    UKM_TRY(check_args(ctx, in, n_in, out, "ukm_common"));
    UKM_TRY(need_tax(ctx, flags & UKM_F_TAXID, "ukm_common"));
    return run_tree(ctx, OP_UNION, in, n_in, flags & (UKM_F_TAXID | UKM_F_VALIDATE), true, threshold, UKM_FOLD_PLAIN, out, "ukm_common");
}

extern "C" int ukm_merge_sorted(ukm_ctx* ctx, int mode, const ukm_span* in, int n_in, unsigned flags, ukm_span* out) {
    UKM_TRY(check_args(ctx, in, n_in, out, "ukm_merge_sorted"));
    if (mode < UKM_FOLD_PLAIN || mode > UKM_FOLD_REPEATED_CHUNK) return ukm_fail(ctx, UKM_E_ARG, "ukm_merge_sorted: bad mode");
    if (mode != UKM_FOLD_PLAIN) UKM_TRY(need_tax(ctx, flags & UKM_F_TAXID, "ukm_merge_sorted"));
    return run_tree(ctx, OP_MERGE, in, n_in, flags & UKM_F_TAXID, false, 0, mode, out, "ukm_merge_sorted");
}

extern "C" int ukm_setops_stream(ukm_ctx* ctx, const ukm_span* in, int n_in, const int* ops, int n_ops, unsigned flags,
                                 ukm_span* outs) {
    if (!ctx) return UKM_E_ARG;
    if (!ops || n_ops < 1 || n_ops > 8 || !outs) return ukm_fail(ctx, UKM_E_ARG, "ukm_setops_stream: 1..8 operations and their output spans");
    UKM_TRY(check_args(ctx, in, n_in, outs, "ukm_setops_stream"));
    for (int k = 0; k < n_ops; ++k) {
        if (ops[k] != UKM_OP_INTER && ops[k] != UKM_OP_DIFF && ops[k] != UKM_OP_UNION)
            return ukm_fail(ctx, UKM_E_ARG, "ukm_setops_stream: unknown operation %d", ops[k]);
        if (outs[k].cap && !outs[k].keys) return ukm_fail(ctx, UKM_E_ARG, "ukm_setops_stream: output %d has keys == NULL", k);
    }
    if (flags & (UKM_F_TAXID | UKM_F_MIX_TAXID | UKM_F_COMPARE_TAXID))
        return ukm_fail(ctx, UKM_E_ARG, "ukm_setops_stream: keys only (taxid-carrying operations go through ukm_inter / ukm_diff / ukm_union)");
    bool host = true, dev = true, sorted = true;
    for (int f = 0; f < n_in; ++f) {
        host &= in[f].where != UKM_DEVICE;
        dev &= in[f].where == UKM_DEVICE;
        sorted &= in[f].sorted != 0;
    }
    if (!host && !dev) return ukm_fail(ctx, UKM_E_ARG, "ukm_setops_stream: all inputs must live in the same memory space");
    if (host && sorted) return stream_run(ctx, in, n_in, ops, n_ops, flags, outs);
    // device-resident inputs (nothing to stream) or an unsorted diff subject: the plain calls, one after the other --
    // except that inter and diff of the same files share one pass over the inputs
    bool have[8] = {false};
    {
        int k_inter = -1, k_diff = -1;
        for (int k = n_ops - 1; k >= 0; --k) {
            if (ops[k] == UKM_OP_INTER) k_inter = k;
            if (ops[k] == UKM_OP_DIFF) k_diff = k;
        }
        int k_union = -1;
        for (int k = n_ops - 1; k >= 0; --k)
            if (ops[k] == UKM_OP_UNION) k_union = k;
        if (dev && k_inter >= 0 && k_diff >= 0 && k_union >= 0) {
            bool done = false;
            UKM_TRY(fused_union_inter_diff(ctx, in, n_in, flags, (flags & UKM_F_SHARD) != 0, &outs[k_union], &outs[k_inter], &outs[k_diff], &done));
            if (done) have[k_union] = have[k_inter] = have[k_diff] = true;
        }
        if (dev && k_inter >= 0 && k_diff >= 0 && !have[k_inter]) {
            bool done = false;
            UKM_TRY(fused_inter_diff(ctx, in, n_in, flags, (flags & UKM_F_SHARD) != 0, &outs[k_inter], &outs[k_diff], &done));
            if (done) have[k_inter] = have[k_diff] = true;
        }
    }
    for (int k = 0; k < n_ops; ++k) {
        if (have[k]) continue;
        int r;
        if (ops[k] == UKM_OP_INTER) r = run_chain(ctx, OP_INTER, in, n_in, flags & (UKM_F_VALIDATE | UKM_F_SHARD), &outs[k], "ukm_setops_stream(inter)");
        else if (ops[k] == UKM_OP_DIFF) r = run_chain(ctx, OP_DIFF, in, n_in, flags & UKM_F_VALIDATE, &outs[k], "ukm_setops_stream(diff)");
        else r = run_tree(ctx, OP_UNION, in, n_in, flags & UKM_F_VALIDATE, false, 0, UKM_FOLD_PLAIN, &outs[k], "ukm_setops_stream(union)");
        UKM_TRY(r);
    }
    return UKM_OK;
}
