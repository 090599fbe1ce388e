// nfilter.cu -- single-pass N-way inter / diff of sorted duplicate-free k-mer streams (keys only).
//
// Replaces the file-by-file iteration of inter.go:205-286 (mc <- mc AND file_i, compact after every file) and
// diff.go:380-435 (two-pointer walk + map rebuild per subject) for up to eight files per pass with ONE read of every
// input: the result is file 0 filtered by membership in files 1..nf-1 (inter: found in all, diff: found in none), which
// is exactly what the reference's iteration produces on duplicate-free inputs.
//
// Tiles are chunks of FILE 0 (M = 8 * SUB consecutive keys); the matching segment of every other file is
// [lower_bound(F_f, first key of the chunk), lower_bound(F_f, first key of the next chunk)) -- one direct search per
// file and boundary (nfilter_partition_kernel), no multi-sequence selection.  A persistent CTA streams its tiles
// (round-robin) through a ring of shared-memory slots: the loader warp brings the eight segments in with 1-D TMA bulk
// copies several tiles ahead; each of the eight consumer warps owns SUB file-0 keys of the tile, narrows every other
// segment to its own key range, and runs the reference's per-file early exit on its keys -- survivors are
// re-compacted inside the warp after every file, the last files are probed as (survivor, file) pairs.  The warp's
// result is ONE 64-bit survivor mask per SUB keys, written to global memory: no output offsets, no look-back, no
// block-wide barrier, no dependency between CTAs (the kernel is correct under any scheduling).  A second, small pass
// (count / scan / gather) turns the masks into the compact sorted result.
//
// A segment that does not fit the slot (file f is locally much denser than file 0) is not loaded: the warps look their
// keys up in global memory for that file and that tile.
#include <stdlib.h>

#include "common.cuh"
#include "nway_core.cuh"
#include "nfilter_core.cuh"

namespace {

constexpr int NF_WARPS = 8;                       // consumer warps per CTA
constexpr int NF_THREADS = NF_WARPS * 32 + 32;    // + the loader warp (warp 0)
constexpr int NF_S0 = 64;                         // partition: every NF_S0-th boundary is searched in the whole file first

enum { NFOP_INTER = 0, NFOP_DIFF = 1, NFOP_BOTH = 2 };  // BOTH: inter and diff of the same files in one pass

// ---------------------------------------------------------------------------------------------------
// partition: bounds[t].pos[f] = first element of file f that belongs to tile t; bounds[num_tiles] = the end
// ---------------------------------------------------------------------------------------------------
struct NfPartArgs {
    NwFiles F;
    long long n0;
    int M;          // file-0 keys per tile
    int num_tiles;
};

// Level 0 (parent = 0): boundaries at multiples of `stride` and the end boundary, searched in the whole files, starting
// at the position file 0 suggests.  Level 1: every other boundary inside the bracket of its two level-0 neighbours.
__global__ void nfilter_partition_kernel(const NfPartArgs a, NwBound* __restrict__ bounds, int stride, int parent) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long t = i * stride;
    NwBound lo, hi, out;
    if (parent == 0) {
        const long long first_past = ((long long)a.num_tiles + stride - 1) / stride;  // first i with i * stride >= num_tiles
        if (i > first_past) return;
        if (i == first_past) t = a.num_tiles;  // this thread produces the end boundary
#pragma unroll
        for (int f = 0; f < NW_MAX; ++f) {
            lo.pos[f] = 0;
            hi.pos[f] = f < a.F.nf ? a.F.n[f] : 0;
        }
    } else {
        if (t >= a.num_tiles || t % parent == 0) return;
        const long long pl = t / parent * parent;
        const long long ph = pl + parent < a.num_tiles ? pl + parent : a.num_tiles;
        lo = bounds[pl];
        hi = bounds[ph];
    }
    const bool end = t == a.num_tiles;
    const long long p0 = end ? a.n0 : t * a.M;
    const uint64_t key = a.F.k[0][end ? a.n0 - 1 : p0];
    const double frac = parent == 0 ? (double)p0 / (double)a.n0
                                    : (double)(p0 - lo.pos[0]) / (double)(hi.pos[0] - lo.pos[0] > 0 ? hi.pos[0] - lo.pos[0] : 1);
    lo.pos[0] = hi.pos[0] = p0;  // file 0 is cut by position, not searched
    nw_cut_at(a.F, key, lo, hi, frac, &out);
    if (end) {
        // everything up to and INCLUDING the last key of file 0 belongs to the last tile
#pragma unroll
        for (int f = 1; f < NW_MAX; ++f)
            if (f < a.F.nf && out.pos[f] < a.F.n[f] && a.F.k[f][out.pos[f]] == key) ++out.pos[f];
    }
    out.pos[0] = p0;
    bounds[t] = out;
}

// ---------------------------------------------------------------------------------------------------
// the tile kernel
// ---------------------------------------------------------------------------------------------------
struct NfArgs {
    NwFiles F;
    const NwBound* bounds;
    unsigned long long* masks;  // one word per SUB keys of file 0: bit j = key j of the group survives
    unsigned long long* masks2; // NFOP_BOTH: masks = survivors of inter, masks2 = survivors of diff
    int num_tiles;
    int null_mode;  // measurement aids (UKM_MEASURE builds, results NOT valid): 1 = consumers do no work, 2 = also no loads, 3 = narrowing only
    int* err;
};

struct NfGeom {
    int n[NW_MAX];          // elements of file f in the slot; -1: the segment stayed in global memory
    int off[NW_MAX];        // first element of the segment in the slot
    long long gpos[NW_MAX]; // global segment [gpos, gpos + gn)
    long long gn[NW_MAX];
    int all_smem;           // every segment of the tile is in the slot (the fast path of the consumers)
};

struct NfSub {  // a warp's share of a segment
    long long lo, n;
};

__device__ __forceinline__ int nf_slice_h(const uint64_t* g, long long start) {
    return (int)((reinterpret_cast<uintptr_t>(g + start) & 15u) >> 3);
}

// Branch-free searches with a fixed probe sequence (P = the largest power of two <= n, first probe at n - P, then
// steps P/2 .. 1): ~8 instructions per probe and, when n is the same for every lane, no divergence.
// number of elements of seg[0, n) below x
template <typename IDX>
__device__ __forceinline__ IDX nf_lower_bound(const uint64_t* seg, IDX n, uint64_t x) {
    if (n <= 0) return 0;
    const IDX P = sizeof(IDX) == 8 ? (IDX)(1ull << (63 - __clzll((long long)n))) : (IDX)(1u << (31 - __clz((int)n)));
    IDX c = seg[n - P] < x ? n - P + 1 : 0;
    for (IDX step = P >> 1; step > 0; step >>= 1)
        if (seg[c + step - 1] < x) c += step;
    return c;
}
// does x occur in seg[0, n)
template <typename IDX>
__device__ __forceinline__ bool nf_contains(const uint64_t* seg, IDX n, uint64_t x) {
    if (n <= 0) return false;
    const IDX P = sizeof(IDX) == 8 ? (IDX)(1ull << (63 - __clzll((long long)n))) : (IDX)(1u << (31 - __clz((int)n)));
    IDX pos = seg[n - P] <= x ? n - P : 0;  // the last element <= x, if there is one
    for (IDX step = P >> 1; step > 0; step >>= 1)
        if (seg[pos + step] <= x) pos += step;
    return seg[pos] == x;
}
// two keys against the same segment: two independent load chains, one loop
__device__ __forceinline__ void nf_contains2(const uint64_t* seg, int n, uint64_t x0, uint64_t x1, bool* f0, bool* f1) {
    if (n <= 0) {
        *f0 = *f1 = false;
        return;
    }
    const int P = 1 << (31 - __clz(n));
    const uint64_t v = seg[n - P];
    int p0 = v <= x0 ? n - P : 0, p1 = v <= x1 ? n - P : 0;
    for (int step = P >> 1; step > 0; step >>= 1) {
        const uint64_t v0 = seg[p0 + step], v1 = seg[p1 + step];
        if (v0 <= x0) p0 += step;
        if (v1 <= x1) p1 += step;
    }
    *f0 = seg[p0] == x0;
    *f1 = seg[p1] == x1;
}

// floor(q / R) = (q * nf_inv[R]) >> 10 for q < 128, R in 1..7
__constant__ unsigned short nf_inv[8] = {0, 1024, 512, 342, 256, 205, 171, 147};

template <int SUB, int CAP>
struct NfShape {
    static constexpr int M = NF_WARPS * SUB;
    static constexpr int SLOT_E = (CAP + 2 * NW_MAX + 2 + 1) & ~1;  // + per-segment alignment slack + read-past padding
};

// ---- one warp, SUB keys of file 0, every segment in shared memory: the fast path ---------------------------------
// (the per-lane searches nf_find / nf_find2 / nf_rank live in nfilter_core.cuh, shared with the host model)

// OP = NFOP_BOTH computes inter and diff together: after file 1 every key is a candidate of exactly ONE of them (found: it
// can only still be in the intersection, not found: only in the difference), so the later files cost one extra dense round
// instead of a second pass over all inputs.  *type gets the candidates' kind (bit set: inter); the survivors of inter are
// alive & type, those of diff alive & ~type.
template <int OP, int SUB>
__device__ __forceinline__ unsigned long long nf_tile_fast(const uint64_t* slot, const NfGeom& g, uint8_t* list, int w, int cnt, int nf,
                                                           unsigned lane, bool narrow_only, unsigned long long* type) {
    const unsigned FULL = 0xffffffffu;
    const unsigned lt = lanemask_lt();
    const uint32_t slot_a = smem_u32(slot);
    const int fq = (int)lane & 7;
    // lane f (and f + 8, f + 16, f + 24): the tile's segment of file f
    const int tN = fq < nf ? g.n[fq] : 0;
    const uint32_t tseg = slot_a + (unsigned)(fq < nf ? g.off[fq] : 0) * 8u;
    const uint32_t f0a = slot_a + (unsigned)(g.off[0] + w * SUB) * 8u;
    // ---- narrow every segment to this warp's key range: lanes 1..7 the lower end, lanes 9..15 the upper end ----
    uint32_t sub_a;  // lane f: shared address and length of file f's share
    int sub_n;
    {
        const bool upper = lane >= 8;
        const bool to_end = w * SUB + cnt >= g.n[0];  // the last warp of the tile takes the rest
        const bool act = lane < 16 && fq >= 1 && fq < nf;
        const int nn = act ? tN : 0;
        const int lgmax = nf_lg(__reduce_max_sync(FULL, nn));
        const uint64_t key = nf_lds(f0a + ((upper && !to_end) ? (unsigned)cnt * 8u : 0u));
        int pos = nf_rank(tseg, (upper && to_end) ? 0 : nn, key, lgmax);
        if (upper && to_end) pos = nn;
        const int hi = __shfl_down_sync(FULL, pos, 8);
        sub_a = tseg + (unsigned)pos * 8u;
        sub_n = hi - pos;
    }
    unsigned alo = cnt >= 32 ? FULL : ((1u << cnt) - 1u);
    unsigned ahi = (SUB > 32 && cnt > 32) ? (cnt >= 64 ? FULL : ((1u << (cnt - 32)) - 1u)) : 0u;
    unsigned tlo = 0, thi = 0;  // NFOP_BOTH: the inter candidates
    if (narrow_only) return ((unsigned long long)ahi << 32) | alo;
    // ---- file 1: every key is still there, the lane takes its own keys ----
    int f = 1;
    {
        const uint32_t seg = __shfl_sync(FULL, sub_a, 1);
        const int n = __shfl_sync(FULL, sub_n, 1);
        const int lgmax = nf_lg(n);
        const uint64_t x0 = nf_lds(f0a + (lane < (unsigned)cnt ? lane : 0u) * 8u);
        bool fd0, fd1 = false;
        if (SUB > 32 && cnt > 32) {
            const uint64_t x1 = nf_lds(f0a + ((int)lane + 32 < cnt ? lane + 32 : 0u) * 8u);
            nf_find2(seg, n, x0, x1, lgmax, &fd0, &fd1);
        } else {
            fd0 = nf_find(seg, n, x0, lgmax);
        }
        if (OP == NFOP_BOTH) {
            tlo = __ballot_sync(FULL, fd0);
            if (SUB > 32) thi = __ballot_sync(FULL, fd1);
        } else {
            alo &= ~__ballot_sync(FULL, OP == NFOP_INTER ? !fd0 : fd0);
            if (SUB > 32) ahi &= ~__ballot_sync(FULL, OP == NFOP_INTER ? !fd1 : fd1);
        }
        f = 2;
    }
    // a survivor leaves when a file decides against it: inter candidates on a miss, diff candidates on a hit
    auto drops = [&](unsigned ix, bool found) -> bool {
        if (OP == NFOP_INTER) return !found;
        if (OP == NFOP_DIFF) return found;
        const bool is_inter = ((ix < 32 ? tlo >> ix : thi >> (ix - 32)) & 1u) != 0;
        return is_inter ? !found : found;
    };
    // ---- the other files: survivors are re-listed every round; few survivors take several files per round ----
    const uint32_t list_a = smem_u32(list);
    while (f < nf && (alo | ahi)) {
        const int a = __popc(alo) + __popc(ahi);
        if (alo >> lane & 1u) list[__popc(alo & lt)] = (uint8_t)lane;
        if (SUB > 32 && (ahi >> lane & 1u)) list[__popc(alo) + __popc(ahi & lt)] = (uint8_t)(lane + 32);
        __syncwarp();
        unsigned klo = 0, khi = 0;  // this lane's verdicts: survivors to drop
        if (a > 32) {
            // one file, two survivors per lane
            const uint32_t seg = __shfl_sync(FULL, sub_a, f);
            const int n = __shfl_sync(FULL, sub_n, f);
            const int lgmax = nf_lg(n);
            const bool v1 = (int)lane + 32 < a;
            unsigned i0, i1;
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(i0) : "r"(list_a + lane));
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(i1) : "r"(list_a + (v1 ? lane + 32 : lane)));
            bool fd0, fd1;
            nf_find2(seg, n, nf_lds(f0a + i0 * 8u), nf_lds(f0a + i1 * 8u), lgmax, &fd0, &fd1);
            if (drops(i0, fd0)) {
                if (i0 < 32) klo |= 1u << i0;
                else khi |= 1u << (i0 - 32);
            }
            if (v1 && drops(i1, fd1)) {
                if (i1 < 32) klo |= 1u << i1;
                else khi |= 1u << (i1 - 32);
            }
            f += 1;
        } else {
            // (survivor, file) pairs: 1, 2, 4 or 8 files per round, whatever fills the warp
            const int lgr = a > 16 ? 0 : a > 8 ? 1 : a > 4 ? 2 : 3;
            const int si = (int)lane >> lgr, ff = f + ((int)lane & ((1 << lgr) - 1));
            const bool valid = si < a && ff < nf;
            const uint32_t seg = __shfl_sync(FULL, sub_a, ff & 7);
            const int n_ff = __shfl_sync(FULL, sub_n, ff & 7);
            const int n = valid ? n_ff : 0;
            const int lgmax = nf_lg(__reduce_max_sync(FULL, n));
            unsigned ix;
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(ix) : "r"(list_a + (valid ? si : 0)));
            const bool fd = nf_find(seg, n, nf_lds(f0a + ix * 8u), lgmax);
            if (valid && drops(ix, fd)) {
                if (ix < 32) klo = 1u << ix;
                else khi = 1u << (ix - 32);
            }
            f += 1 << lgr;
        }
        alo &= ~__reduce_or_sync(FULL, klo);
        if (SUB > 32) ahi &= ~__reduce_or_sync(FULL, khi);
        __syncwarp();  // the list is rewritten in the next round
    }
    if (OP == NFOP_BOTH) *type = ((unsigned long long)thi << 32) | tlo;
    return ((unsigned long long)ahi << 32) | alo;
}

// ---- the general path of a warp: some segment of the tile stayed in global memory ------------------------------
template <int OP, int SUB>
__device__ __noinline__ unsigned long long nf_tile_generic(const uint64_t* slot, const NfGeom& g, const uint64_t* const* s_fk, NfSub* sub,
                                                           uint8_t* list, int w, int cnt, int nf, unsigned lane, unsigned long long* type) {
    const unsigned lt = lanemask_lt();
    const int n0t = g.n[0];
    unsigned alo = cnt >= 32 ? 0xffffffffu : ((1u << cnt) - 1u);
    unsigned ahi = (SUB > 32 && cnt > 32) ? (cnt >= 64 ? 0xffffffffu : ((1u << (cnt - 32)) - 1u)) : 0u;
    unsigned tlo = 0, thi = 0;  // NFOP_BOTH: the inter candidates (found in file 1)
    auto drops = [&](unsigned ix, bool found) -> bool {
        if (OP == NFOP_INTER) return !found;
        if (OP == NFOP_DIFF) return found;
        const bool is_inter = ((ix < 32 ? tlo >> ix : thi >> (ix - 32)) & 1u) != 0;
        return is_inter ? !found : found;
    };
    const uint64_t* f0 = slot + g.off[0] + w * SUB;
    // ---- narrow every segment to this warp's key range: lanes 0..7 the lower end, lanes 8..15 the upper end ----
    {
        const int f = (int)lane & 7;
        const bool upper = lane >= 8;
        long long pos = 0;
        if (lane < 16 && f >= 1 && f < nf) {
            const bool to_end = upper && (w * SUB + cnt >= n0t);  // the last warp of the tile takes the rest
            const uint64_t key = upper ? (to_end ? 0 : f0[cnt]) : f0[0];
            if (g.n[f] >= 0) pos = to_end ? g.n[f] : nf_lower_bound<int>(slot + g.off[f], g.n[f], key);
            else pos = to_end ? g.gn[f] : nf_lower_bound<long long>(s_fk[f] + g.gpos[f], g.gn[f], key);
        }
        const long long hi = __shfl_down_sync(0xffffffffu, pos, 8);
        if (lane < 8 && f >= 1 && f < nf) {
            sub[f].lo = pos;
            sub[f].n = hi - pos;
        }
    }
    __syncwarp();
    // ---- file by file with early exit; after the first file the survivors are re-listed every round ----
    int f = 1;
    bool first = true;
    while (f < nf && (alo | ahi)) {
        const int a = __popc(alo) + __popc(ahi);
        const int R = nf - f;
        unsigned klo = 0, khi = 0;  // this lane's verdicts: survivors to drop
        if (first) {
            // every key is still there: the lane takes its own two keys (no list)
            first = false;
            const long long slo = sub[f].lo, sn = sub[f].n;
            const uint64_t x0 = f0[lane < (unsigned)cnt ? lane : 0];
            const uint64_t x1 = f0[(SUB > 32 && (int)lane + 32 < cnt) ? lane + 32 : 0];
            bool fd0, fd1 = false;
            if (g.n[f] >= 0) {
                const uint64_t* seg = slot + g.off[f] + (int)slo;
                if (SUB > 32 && cnt > 32) nf_contains2(seg, (int)sn, x0, x1, &fd0, &fd1);
                else fd0 = nf_contains<int>(seg, (int)sn, x0);
            } else {
                const uint64_t* seg = s_fk[f] + g.gpos[f] + slo;
                fd0 = nf_contains<long long>(seg, sn, x0);
                if (SUB > 32 && cnt > 32) fd1 = nf_contains<long long>(seg, sn, x1);
            }
            if (OP == NFOP_BOTH) {
                tlo = __ballot_sync(0xffffffffu, fd0);
                thi = __ballot_sync(0xffffffffu, fd1);
            } else {
                if (OP == NFOP_INTER ? !fd0 : fd0) klo = 1u << lane;
                if (OP == NFOP_INTER ? !fd1 : fd1) khi = 1u << lane;
            }
            f += 1;
        } else {
            if (alo >> lane & 1u) list[__popc(alo & lt)] = (uint8_t)lane;
            if (SUB > 32 && (ahi >> lane & 1u)) list[__popc(alo) + __popc(ahi & lt)] = (uint8_t)(lane + 32);
            __syncwarp();
            if (a * R > 64 || R == 1) {
                // one file, every survivor
                const long long slo = sub[f].lo, sn = sub[f].n;
                const int i0 = list[lane < (unsigned)a ? lane : 0];
                const int i1 = list[(int)lane + 32 < a ? lane + 32 : 0];
                bool fd0, fd1 = false;
                if (g.n[f] >= 0) {
                    const uint64_t* seg = slot + g.off[f] + (int)slo;
                    if (a > 32) nf_contains2(seg, (int)sn, f0[i0], f0[i1], &fd0, &fd1);
                    else fd0 = nf_contains<int>(seg, (int)sn, f0[i0]);
                } else {
                    const uint64_t* seg = s_fk[f] + g.gpos[f] + slo;
                    fd0 = nf_contains<long long>(seg, sn, f0[i0]);
                    if (a > 32) fd1 = nf_contains<long long>(seg, sn, f0[i1]);
                }
                if ((int)lane < a && drops((unsigned)i0, fd0)) {
                    if (i0 < 32) klo |= 1u << i0;
                    else khi |= 1u << (i0 - 32);
                }
                if ((int)lane + 32 < a && drops((unsigned)i1, fd1)) {
                    if (i1 < 32) klo |= 1u << i1;
                    else khi |= 1u << (i1 - 32);
                }
                f += 1;
            } else {
                // the remaining files together: one (survivor, file) pair per lane and round (at most two rounds)
                const int pairs = a * R;
                const unsigned inv = nf_inv[R];
                for (int qd = (int)lane; qd < pairs; qd += 32) {
                    const int si = (int)(((unsigned)qd * inv) >> 10), ff = f + (qd - si * R);
                    const int ix = list[si];
                    const long long slo = sub[ff].lo, sn = sub[ff].n;
                    bool fd;
                    if (g.n[ff] >= 0) fd = nf_contains<int>(slot + g.off[ff] + (int)slo, (int)sn, f0[ix]);
                    else fd = nf_contains<long long>(s_fk[ff] + g.gpos[ff] + slo, sn, f0[ix]);
                    if (drops((unsigned)ix, fd)) {
                        if (ix < 32) klo |= 1u << ix;
                        else khi |= 1u << (ix - 32);
                    }
                }
                f = nf;
            }
        }
        alo &= ~__reduce_or_sync(0xffffffffu, klo);
        if (SUB > 32) ahi &= ~__reduce_or_sync(0xffffffffu, khi);
        __syncwarp();  // the list is rewritten in the next round
    }
    if (OP == NFOP_BOTH) *type = ((unsigned long long)thi << 32) | tlo;
    return ((unsigned long long)ahi << 32) | alo;
}

template <int OP, int SUB, int CAP, int SLOTS, int MINB>
__global__ void __launch_bounds__(NF_THREADS, MINB) nfilter_kernel(const NfArgs p) {
    using SH = NfShape<SUB, CAP>;
    extern __shared__ __align__(16) unsigned char nf_smem[];
    uint64_t* s_slots = reinterpret_cast<uint64_t*>(nf_smem);  // SLOTS * SLOT_E
    __shared__ __align__(8) uint64_t full_bar[SLOTS], empty_bar[SLOTS];
    __shared__ NfGeom s_geom[SLOTS];
    __shared__ const uint64_t* s_fk[NW_MAX];
    __shared__ NfSub s_sub[NF_WARPS][NW_MAX];
    __shared__ uint8_t s_list[NF_WARPS][64];

    const int G = gridDim.x;
    const int n_my = (p.num_tiles - (int)blockIdx.x + G - 1) / G;
    const int nf = p.F.nf;
    if (threadIdx.x == 0) {
        for (int s = 0; s < SLOTS; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], NF_WARPS);
        }
#pragma unroll
        for (int f = 0; f < NW_MAX; ++f) s_fk[f] = p.F.k[f];
        mbar_fence_init();
    }
    __syncthreads();
    const unsigned lane = lane_id();

    if (threadIdx.x < 32) {
        // ================= loader warp: four tiles per batch, lane = 8 * (tile in batch) + file =================
        // The cut positions of a batch are fetched two batches ahead and the unaligned head / tail elements one batch
        // ahead, so the global-memory latency of that bookkeeping is paid once per four tiles and never between tiles.
        const int q = (int)lane >> 3, f = (int)lane & 7;
        const bool mine = f < nf;
        const uint64_t* fk = mine ? s_fk[f] : nullptr;
        const int n_batches = (n_my + 3) >> 2;
        long long lo_a = 0, nn_a = 0, lo_b = 0, nn_b = 0;
        uint64_t hv_a = 0, tv_a = 0;
        auto fetch_bounds = [&](int b, long long* lo, long long* nn) {
            *lo = 0;
            *nn = 0;
            const int li = b * 4 + q;
            if (mine && li < n_my) {
                const int t = (int)blockIdx.x + li * G;
                *lo = p.bounds[t].pos[f];
                *nn = p.bounds[t + 1].pos[f] - *lo;
            }
        };
        auto fetch_edges = [&](long long lo, long long nn, uint64_t* hv, uint64_t* tv) {
            *hv = 0;
            *tv = 0;
            if (mine && nn > 0 && nn <= CAP && p.null_mode != 2) {
                *hv = fk[lo];
                *tv = fk[lo + nn - 1];
            }
        };
        fetch_bounds(0, &lo_a, &nn_a);
        fetch_edges(lo_a, nn_a, &hv_a, &tv_a);
        fetch_bounds(1, &lo_b, &nn_b);
        for (int b = 0; b < n_batches; ++b) {
            uint64_t hv_b, tv_b;
            long long lo_c, nn_c;
            fetch_edges(lo_b, nn_b, &hv_b, &tv_b);  // in flight while this batch is issued
            fetch_bounds(b + 2, &lo_c, &nn_c);
            const long long lo = lo_a, nn = nn_a;
            if (__any_sync(0xffffffffu, nn < 0)) {  // cannot happen: the cuts of a sorted file are monotone
                if (lane == 0) atomicExch(p.err, (int)UKM_E_INTERNAL);
            }
            const int h = (mine && nn > 0) ? nf_slice_h(fk, lo) : 0;
            const int padded = (mine && nn > 0 && nn <= CAP) ? (int)((h + nn + 1) & ~1ll) : 0;
            for (int qq = 0; qq < 4; ++qq) {
                const int li = b * 4 + qq;
                if (li >= n_my) break;
                const int s = li % SLOTS, u = li / SLOTS;
                if (u > 0 && !mbar_wait(&empty_bar[s], (unsigned)(u - 1) & 1u)) {
                    if (lane == 0) atomicExch(p.err, (int)UKM_E_INTERNAL);
                }
                // greedy layout in file order: a segment goes into the slot if it still fits, else it stays in global memory
                int base = 0, my_base = 0;
                bool my_fit = false;
#pragma unroll
                for (int ff = 0; ff < NW_MAX; ++ff) {
                    const int src = qq * 8 + ff;
                    const int pf = __shfl_sync(0xffffffffu, padded, src);
                    const long long nnf = __shfl_sync(0xffffffffu, nn, src);
                    const bool fit = nnf <= 0 || (nnf <= CAP && base + pf <= CAP);
                    if ((int)lane == src) {
                        my_fit = fit;
                        my_base = base;
                    }
                    if (fit) base += pf;
                }
                const bool act = q == qq && mine;
                uint64_t* slot = s_slots + (size_t)s * SH::SLOT_E;
                if (p.null_mode == 2) {  // measurement aid: the barrier ring alone
                    if (act) s_geom[s].n[f] = 0;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full_bar[s]);
                    continue;
                }
                const int n = (act && my_fit && nn > 0) ? (int)nn : 0;
                int head = 0, body = 0;
                if (n > 0) {
                    head = h ? 1 : 0;
                    body = (n - head) & ~1;
                    if (head) slot[my_base + h] = hv_a;
                    if (head + body < n) slot[my_base + h + n - 1] = tv_a;
                }
                unsigned bytes = (unsigned)body * 8u;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, d);
                const bool all_fit = __all_sync(0xffffffffu, q != qq || !mine || my_fit);
                if (lane == 0) s_geom[s].all_smem = all_fit ? 1 : 0;
                if (act) {
                    s_geom[s].n[f] = my_fit ? n : -1;
                    s_geom[s].off[f] = my_base + h;
                    s_geom[s].gpos[f] = lo;
                    s_geom[s].gn[f] = nn;
                }
                __syncwarp();
                if (lane == 0) mbar_expect_tx(&full_bar[s], bytes);  // arrive (release: plain stores + geometry) + tx count
                __syncwarp();
                if (body) tma_load_1d(slot + my_base + h + head, fk + lo + head, (unsigned)body * 8u, &full_bar[s]);
            }
            lo_a = lo_b; nn_a = nn_b; hv_a = hv_b; tv_a = tv_b;
            lo_b = lo_c; nn_b = nn_c;
        }
        return;
    }

    // ================= consumer warps: SUB keys of file 0 each, no block-wide synchronisation =================
    const int w = ((int)threadIdx.x >> 5) - 1;
    for (int i = 0; i < n_my; ++i) {
        const int s = i % SLOTS, u = i / SLOTS;
        const int tile = (int)blockIdx.x + i * G;
        const uint64_t* slot = s_slots + (size_t)s * SH::SLOT_E;
        if (!mbar_wait(&full_bar[s], (unsigned)u & 1u)) {
            if (lane == 0) atomicExch(p.err, (int)UKM_E_INTERNAL);
        }
        const NfGeom& g = s_geom[s];
        int cnt = g.n[0] - w * SUB;
        cnt = cnt < 0 ? 0 : (cnt > SUB ? SUB : cnt);
        unsigned long long alive = 0, type = 0;
        if (cnt > 0 && p.null_mode != 2) {
            if (p.null_mode == 1) alive = cnt >= 64 ? ~0ull : ((1ull << cnt) - 1ull);
            else if (g.all_smem) alive = nf_tile_fast<OP, SUB>(slot, g, s_list[w], w, cnt, nf, lane, p.null_mode == 3, &type);
            else alive = nf_tile_generic<OP, SUB>(slot, g, s_fk, s_sub[w], s_list[w], w, cnt, nf, lane, &type);
        }
        if (lane == 0) {
            if (OP == NFOP_BOTH) {
                p.masks[(size_t)tile * NF_WARPS + w] = alive & type;
                p.masks2[(size_t)tile * NF_WARPS + w] = alive & ~type;
            } else {
                p.masks[(size_t)tile * NF_WARPS + w] = alive;
            }
        }
        __syncwarp();  // every lane is past its reads of the slot
        if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
}

// ---------------------------------------------------------------------------------------------------
// masks -> compact sorted result: per-block survivor counts, one scan, gather
// ---------------------------------------------------------------------------------------------------
constexpr int NG_THREADS = 256;
constexpr int NG_PER = 8;                            // mask words per thread
constexpr int NG_BLOCK = NG_THREADS * NG_PER;        // mask words per block

__global__ void __launch_bounds__(NG_THREADS) nfilter_count_kernel(const unsigned long long* __restrict__ masks, size_t n_masks,
                                                                   unsigned long long* __restrict__ block_sums) {
    __shared__ unsigned s_w[NG_THREADS / 32];
    const size_t base = (size_t)blockIdx.x * NG_BLOCK;
    unsigned c = 0;
#pragma unroll
    for (int q = 0; q < NG_PER; ++q) {
        const size_t i = base + (size_t)q * NG_THREADS + threadIdx.x;
        if (i < n_masks) c += (unsigned)__popcll(masks[i]);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int q = 0; q < NG_THREADS / 32; ++q) t += s_w[q];
        block_sums[blockIdx.x] = t;
    }
}

// exclusive scan of nb block sums in place (one CTA), total -> *total
__global__ void __launch_bounds__(1024) nfilter_scan_kernel(unsigned long long* __restrict__ sums, int nb, unsigned long long* __restrict__ total) {
    __shared__ unsigned long long s_w[33];
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + (int)threadIdx.x;
        const unsigned long long v = i < nb ? sums[i] : 0ull;
        unsigned long long incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
            if ((threadIdx.x & 31) >= (unsigned)d) incl += t;
        }
        if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned long long x = s_w[threadIdx.x];
            unsigned long long xi = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, xi, d);
                if (threadIdx.x >= (unsigned)d) xi += t;
            }
            s_w[threadIdx.x] = xi - x;
            if (threadIdx.x == 31) s_w[32] = xi;
        }
        __syncthreads();
        if (i < nb) sums[i] = s_carry + s_w[threadIdx.x >> 5] + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += s_w[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}

// block b writes the survivors of its NG_BLOCK mask words; a warp owns 32 * NG_PER consecutive words
template <int SUB>
__global__ void __launch_bounds__(NG_THREADS) nfilter_gather_kernel(const unsigned long long* __restrict__ masks, size_t n_masks,
                                                                    const uint64_t* __restrict__ F0, const unsigned long long* __restrict__ block_offs,
                                                                    uint64_t* __restrict__ out) {
    __shared__ unsigned s_w[NG_THREADS / 32 + 1];
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const size_t wbase = (size_t)blockIdx.x * NG_BLOCK + (size_t)wid * 32 * NG_PER;
    unsigned long long m[NG_PER];
    unsigned c = 0;
#pragma unroll
    for (int q = 0; q < NG_PER; ++q) {
        const size_t i = wbase + (size_t)q * 32 + lane;
        m[q] = i < n_masks ? masks[i] : 0ull;
        c += (unsigned)__popcll(m[q]);
    }
    unsigned wtot = c;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) wtot += __shfl_xor_sync(0xffffffffu, wtot, d);
    if (lane == 0) s_w[wid] = wtot;
    __syncthreads();
    unsigned long long o = block_offs[blockIdx.x];
    for (unsigned q = 0; q < wid; ++q) o += s_w[q];
    if (wtot == 0) return;
#pragma unroll
    for (int q = 0; q < NG_PER; ++q) {
        const unsigned cq = (unsigned)__popcll(m[q]);
        unsigned incl = cq;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (unsigned)d) incl += t;
        }
        const unsigned excl = incl - cq;
        unsigned nz = __ballot_sync(0xffffffffu, cq != 0);
        while (nz) {
            const int src = __ffs(nz) - 1;
            nz &= nz - 1;
            const unsigned long long mm = __shfl_sync(0xffffffffu, m[q], src);
            const unsigned eo = __shfl_sync(0xffffffffu, excl, src);
            const size_t kbase = (wbase + (size_t)q * 32 + src) * SUB;
            if (mm >> lane & 1ull) out[o + eo + __popcll(mm & ((1ull << lane) - 1ull))] = F0[kbase + lane];
            if (SUB > 32 && (mm >> (lane + 32) & 1ull)) out[o + eo + __popcll(mm & ((1ull << (lane + 32)) - 1ull))] = F0[kbase + lane + 32];
        }
        o += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
int nf_env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// OP = NFOP_BOTH: outK / n_out = inter, outK2 / n_out2 = diff
template <int OP, int SUB, int CAP, int SLOTS, int MINB>
int launch_nfilter(ukm_ctx* ctx, NfArgs a, NfPartArgs pa, ukm_tmp& tmp, const uint64_t* F0, uint64_t* outK, size_t* n_out,
                   uint64_t* outK2, size_t* n_out2) {
    using SH = NfShape<SUB, CAP>;
    constexpr size_t smem = (size_t)SLOTS * SH::SLOT_E * 8;
    auto kern = nfilter_kernel<OP, SUB, CAP, SLOTS, MINB>;
    int per_sm = 0;
    UKM_TRY(ukm_kernel_config(ctx, kern, smem, NF_THREADS, &per_sm));
    pa.M = SH::M;
    const int num_tiles = (int)((pa.n0 + SH::M - 1) / SH::M);
    pa.num_tiles = num_tiles;
    const size_t n_masks = (size_t)num_tiles * NF_WARPS;
    const int nb = (int)((n_masks + NG_BLOCK - 1) / NG_BLOCK);
    NwBound* d_bounds = nullptr;
    unsigned long long* d_masks = nullptr;
    unsigned long long* d_sums = nullptr;
    UKM_TRY(tmp.alloc(&d_bounds, (size_t)num_tiles + 1));
    UKM_TRY(tmp.alloc(&d_masks, n_masks * (OP == NFOP_BOTH ? 2 : 1)));
    UKM_TRY(tmp.alloc(&d_sums, ((size_t)nb + 1) * (OP == NFOP_BOTH ? 2 : 1)));
    {
        const int n0 = num_tiles / NF_S0 + 2;
        nfilter_partition_kernel<<<(n0 + 63) / 64, 64, 0, ctx->stream>>>(pa, d_bounds, NF_S0, 0);
        UKM_LAUNCHED(ctx);
        nfilter_partition_kernel<<<(num_tiles + 127) / 128, 128, 0, ctx->stream>>>(pa, d_bounds, 1, NF_S0);
        UKM_LAUNCHED(ctx);
    }
    a.bounds = d_bounds;
    a.masks = d_masks;
    a.masks2 = d_masks + n_masks;
    a.num_tiles = num_tiles;
    int grid = per_sm * ctx->sm_count;
    if (grid > num_tiles) grid = num_tiles;
    kern<<<grid, NF_THREADS, smem, ctx->stream>>>(a);
    UKM_LAUNCHED(ctx);
    nfilter_count_kernel<<<nb, NG_THREADS, 0, ctx->stream>>>(d_masks, n_masks, d_sums);
    UKM_LAUNCHED(ctx);
    nfilter_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_sums, nb, d_sums + nb);
    UKM_LAUNCHED(ctx);
    nfilter_gather_kernel<SUB><<<nb, NG_THREADS, 0, ctx->stream>>>(d_masks, n_masks, F0, d_sums, outK);
    UKM_LAUNCHED(ctx);
    UKM_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch, d_sums + nb, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (OP == NFOP_BOTH) {
        unsigned long long* d_sums2 = d_sums + nb + 1;
        nfilter_count_kernel<<<nb, NG_THREADS, 0, ctx->stream>>>(a.masks2, n_masks, d_sums2);
        UKM_LAUNCHED(ctx);
        nfilter_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_sums2, nb, d_sums2 + nb);
        UKM_LAUNCHED(ctx);
        nfilter_gather_kernel<SUB><<<nb, NG_THREADS, 0, ctx->stream>>>(a.masks2, n_masks, F0, d_sums2, outK2);
        UKM_LAUNCHED(ctx);
        UKM_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch + 1, d_sums2 + nb, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    UKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_out = (size_t)ctx->h_scratch[0];
    if (OP == NFOP_BOTH) *n_out2 = (size_t)ctx->h_scratch[1];
    tmp.free_now(d_bounds);
    tmp.free_now(d_masks);
    tmp.free_now(d_sums);
    return UKM_OK;
}

template <int OP>
int launch_nfilter_sub(ukm_ctx* ctx, int sub, const NfArgs& a, const NfPartArgs& pa, ukm_tmp& tmp, const uint64_t* F0, uint64_t* outK,
                       size_t* n_out, uint64_t* outK2 = nullptr, size_t* n_out2 = nullptr) {
    // slot capacity x ring depth x CTAs per SM: 2 x 36 KB x 3 (default: C3 inter 10.0 ms on B200); UKM_NFILTER_CFG = 1: 4 x 27 KB x 2
    // (19.5 ms: tiles of 256 file-0 keys), 2: 3 x 36 KB x 2 (11.4 ms: 16 instead of 24 consumer warps per SM) -- A/B runs
    const int cfg = nf_env_int("UKM_NFILTER_CFG", 0);
#define NF_LAUNCH(CAP, SLOTS, MINB)                                                                                   \
    switch (sub) {                                                                                                    \
        case 64: return launch_nfilter<OP, 64, CAP, SLOTS, MINB>(ctx, a, pa, tmp, F0, outK, n_out, outK2, n_out2);    \
        case 32: return launch_nfilter<OP, 32, CAP, SLOTS, MINB>(ctx, a, pa, tmp, F0, outK, n_out, outK2, n_out2);    \
        case 16: return launch_nfilter<OP, 16, CAP, SLOTS, MINB>(ctx, a, pa, tmp, F0, outK, n_out, outK2, n_out2);    \
        default: return launch_nfilter<OP, 8, CAP, SLOTS, MINB>(ctx, a, pa, tmp, F0, outK, n_out, outK2, n_out2);     \
    }
    if (cfg == 1) { NF_LAUNCH(3392, 4, 2) }
    if (cfg == 2) { NF_LAUNCH(4576, 3, 2) }
    NF_LAUNCH(4576, 2, 3)
#undef NF_LAUNCH
}

}  // namespace

bool ukm_nfilter_enabled() {
    const char* e = getenv("UKM_NFILTER");
    return !(e && e[0] == '0');
}

// keys[0] filtered by membership in keys[1..nf-1] (2..8 sorted duplicate-free device arrays): inter keeps the keys
// found in every other array, diff the keys found in none.  outK capacity >= n[0].  *declined = true (nothing
// written) when file 0 is so sparse against the others that streaming them is the wrong algorithm (the caller looks
// file 0's keys up file by file instead).
int ukm_nfilter(ukm_ctx* ctx, bool inter, const uint64_t* const* keys, const size_t* n, int nf, uint64_t* outK, size_t* n_out,
                bool* declined) {
    *declined = false;
    *n_out = 0;
    if (nf < 2 || nf > NW_MAX) return ukm_fail(ctx, UKM_E_ARG, "nfilter: 2..8 inputs");
    if (n[0] == 0) return UKM_OK;
    NfArgs a;
    NfPartArgs pa;
    long long total = 0;
    for (int f = 0; f < NW_MAX; ++f) {
        a.F.k[f] = f < nf ? keys[f] : nullptr;
        a.F.n[f] = f < nf ? (long long)n[f] : 0;
        total += a.F.n[f];
    }
    a.F.nf = nf;
    a.err = ctx->d_err;
    a.null_mode = 0;
#ifdef UKM_MEASURE  // measurement build only (make measure): the null modes produce no valid result
    a.null_mode = nf_env_int("UKM_NFILTER_NULL", 0);
#endif
    pa.F = a.F;
    pa.n0 = (long long)n[0];
    // file-0 keys per warp so that a tile (8 warps) fits the slot with room for the spread of the segment sizes
    const double ratio = (double)total / (double)n[0];
    const double cap = nf_env_int("UKM_NFILTER_CFG", 0) == 1 ? 3392.0 : 4576.0;
    int sub = 0;
    for (int c = 64; c >= 8; c >>= 1)
        if ((double)(NF_WARPS * c) * ratio <= 0.9 * cap) { sub = c; break; }
    const int force = nf_env_int("UKM_NFILTER_SUB", 0);
    if (force == 64 || force == 32 || force == 16 || force == 8) sub = force;
    if (sub == 0) {
        *declined = true;
        return UKM_OK;
    }
    ukm_tmp tmp(ctx);
    {
        ukm_stat_scope st(ctx, inter ? "setop_inter_nway" : "setop_diff_nway", (double)total * 8.0);
        if (inter) UKM_TRY(launch_nfilter_sub<NFOP_INTER>(ctx, sub, a, pa, tmp, keys[0], outK, n_out));
        else UKM_TRY(launch_nfilter_sub<NFOP_DIFF>(ctx, sub, a, pa, tmp, keys[0], outK, n_out));
    }
    if (ctx->stats_on && !ctx->pending.empty()) ctx->pending.back().bytes += (double)*n_out * 8.0;
    return UKM_OK;
}

// inter AND diff of the same files in one pass (every input read once for the two results): outI / outD capacity >= n[0].
int ukm_nfilter_both(ukm_ctx* ctx, const uint64_t* const* keys, const size_t* n, int nf, uint64_t* outI, size_t* n_i, uint64_t* outD,
                     size_t* n_d, bool* declined) {
    *declined = false;
    *n_i = *n_d = 0;
    if (nf < 2 || nf > NW_MAX) return ukm_fail(ctx, UKM_E_ARG, "nfilter: 2..8 inputs");
    if (n[0] == 0) return UKM_OK;
    NfArgs a;
    NfPartArgs pa;
    long long total = 0;
    for (int f = 0; f < NW_MAX; ++f) {
        a.F.k[f] = f < nf ? keys[f] : nullptr;
        a.F.n[f] = f < nf ? (long long)n[f] : 0;
        total += a.F.n[f];
    }
    a.F.nf = nf;
    a.err = ctx->d_err;
    a.null_mode = 0;
    pa.F = a.F;
    pa.n0 = (long long)n[0];
    const double ratio = (double)total / (double)n[0];
    const double cap = nf_env_int("UKM_NFILTER_CFG", 0) == 1 ? 3392.0 : 4576.0;
    int sub = 0;
    for (int c = 64; c >= 8; c >>= 1)
        if ((double)(NF_WARPS * c) * ratio <= 0.9 * cap) { sub = c; break; }
    const int force = nf_env_int("UKM_NFILTER_SUB", 0);
    if (force == 64 || force == 32 || force == 16 || force == 8) sub = force;
    if (sub == 0) {
        *declined = true;
        return UKM_OK;
    }
    ukm_tmp tmp(ctx);
    {
        ukm_stat_scope st(ctx, "setop_inter_diff_nway", (double)total * 8.0);  // the inputs once (+ both outputs, added below)
        UKM_TRY(launch_nfilter_sub<NFOP_BOTH>(ctx, sub, a, pa, tmp, keys[0], outI, n_i, outD, n_d));
    }
    if (ctx->stats_on && !ctx->pending.empty()) ctx->pending.back().bytes += (double)(*n_i + *n_d) * 8.0;
    return UKM_OK;
}

// masks (one 64-bit word per 64 keys of F0, bit j = key 64 * word + j is kept) -> compact sorted result.  Shared with the
// N-way union, whose inter / diff results ride along as such masks (nway.cu).
int ukm_masks_gather(ukm_ctx* ctx, ukm_tmp& tmp, const unsigned long long* d_masks, size_t n_masks, const uint64_t* F0, uint64_t* outK,
                     size_t* n_out) {
    *n_out = 0;
    if (n_masks == 0) return UKM_OK;
    const int nb = (int)((n_masks + NG_BLOCK - 1) / NG_BLOCK);
    unsigned long long* d_sums = nullptr;
    UKM_TRY(tmp.alloc(&d_sums, (size_t)nb + 1));
    nfilter_count_kernel<<<nb, NG_THREADS, 0, ctx->stream>>>(d_masks, n_masks, d_sums);
    UKM_LAUNCHED(ctx);
    nfilter_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_sums, nb, d_sums + nb);
    UKM_LAUNCHED(ctx);
    nfilter_gather_kernel<64><<<nb, NG_THREADS, 0, ctx->stream>>>(d_masks, n_masks, F0, d_sums, outK);
    UKM_LAUNCHED(ctx);
    UKM_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch + 2, d_sums + nb, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    UKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_out = (size_t)ctx->h_scratch[2];
    tmp.free_now(d_sums);
    return UKM_OK;
}
