// kmer.cu -- k-mer iterators on the device and the `count` pipeline.
//
// Replaces sketches.NewKmerIterator/NextKmer -> kmers.Encode/RevComp/Canonical
// (count.go:321,363) and sketches.NewHashIterator/NextHash -> nthash.NTHi.Next
// (count.go:319,361), then count's dedup map + sort (count.go:373,434-436,531-595) as
// generate -> [scaled filter] -> radix sort -> unique fold.
//
// Arithmetic:
//   code: A0 C1 G2 T3, first base most significant; IUPAC degenerate -> alphabetically first
//         base; other bytes -> UKM_E_ILLEGAL_BASE.  canonical = min(code, revcomp).
//         rolling: code = ((prev & mask) << 2) | b ; rc = (prevRC >> 2) | ((b^3) << 2(k-1))
//   ntHash v1: seeds A 0x3c8bfbb395c60474 C 0x3193c18562a02b4c G 0x20323ed082572324
//         T 0x295549f54be24456, other 0;  fwd' = rol(fwd,1) ^ rol(seed[out],k) ^ seed[in];
//         rev' = ror(rev,1) ^ ror(seedc[out],1) ^ rol(seedc[in],k-1); canonical = min.
//
// Work decomposition: a CTA owns a tile of k-mer START positions in the concatenated base
// buffer; the tile's bases (+ k-1 halo) are staged in shared memory; each thread walks
// KM_CHUNK consecutive starts, computing the first k-mer of a run from scratch and rolling
// afterwards.  K-mers never span records; circular records wrap (reads outside the tile go
// to global memory).
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "select.cuh"

namespace {

constexpr int KM_THREADS = 256;
constexpr int KM_CHUNK = 68;  // starts per thread; 17 words: conflict-free byte walk across lanes
constexpr int KM_TILE = KM_THREADS * KM_CHUNK;
constexpr int KM_HALO = 64;

__device__ __forceinline__ uint64_t rol64d(uint64_t v, unsigned s) {
    s &= 63u;
    return s ? (v << s) | (v >> (64u - s)) : v;
}
__device__ __forceinline__ uint64_t ror64d(uint64_t v, unsigned s) {
    s &= 63u;
    return s ? (v >> s) | (v << (64u - s)) : v;
}

// byte -> class: 0..3 = A C G T(U) exactly; 4 = IUPAC degenerate mapping to A, 5 -> C, 6 -> G
// (2-bit code = cls & 3 after the table below); 255 = illegal for the 2-bit encoder.
__device__ __forceinline__ void build_lut(uint8_t* lut) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = 255;
    __syncthreads();
    if (threadIdx.x == 0) {
        const char* a0 = "MmVvHhRrDdWwNn";
        const char* c1 = "SsBbYy";
        const char* g2 = "Kk";
        for (const char* p = a0; *p; ++p) lut[(uint8_t)*p] = 4;  // -> A
        for (const char* p = c1; *p; ++p) lut[(uint8_t)*p] = 5;  // -> C
        for (const char* p = g2; *p; ++p) lut[(uint8_t)*p] = 6;  // -> G
        lut['A'] = lut['a'] = 0;
        lut['C'] = lut['c'] = 1;
        lut['G'] = lut['g'] = 2;
        lut['T'] = lut['t'] = lut['U'] = lut['u'] = 3;
    }
    __syncthreads();
}

struct KmerArgs {
    const uint8_t* bases;
    size_t n_bases;
    const unsigned long long* rec_off;  // n_rec + 1
    const unsigned long long* out_off;  // n_rec + 1: first output index of each record
    size_t n_rec;
    int k;
    int canonical;
    int circular;
    uint64_t* out;
    int* err;
    // FILTER mode (count): keep range_lo <= code <= range_hi only, append in arbitrary order
    uint64_t range_lo, range_hi;
    unsigned long long* cursor;  // number of codes appended so far
    unsigned long long cap;      // capacity of out; codes beyond it are counted but not stored
    unsigned tile_stride;        // process every tile_stride-th tile only (sampling pass); 1 = all
};

constexpr int KM_STAGE = 256;  // per-warp staging ring of the FILTER mode (flushed 128 codes at a time)

// FILTER = false: every code to out[out_off[r] + position] (record-then-position order).
// FILTER = true : codes inside [range_lo, range_hi] are compacted warp-wide through a shared-memory ring and
//                 appended in 1 KB bursts at a global cursor (order is irrelevant: `count` sorts next).
// The tile's bases are staged as CLASSES (the byte -> class table applied once per base, not once per use); a
// thread whose 68 starts and their k-1 trailing bases lie inside one record and inside the staged tile -- all but
// the threads at record or tile seams -- takes a fast path: 32-bit shared-memory indices, no record bookkeeping,
// one 16-byte seed load per incoming and per outgoing base, codes stored in aligned pairs.
template <bool HASHED, bool FILTER>
__global__ void __launch_bounds__(KM_THREADS) kmer_kernel(const KmerArgs p) {
    __shared__ uint8_t s_b[KM_TILE + KM_HALO + 16];  // class of every staged base
    __shared__ uint8_t s_lut[256];
    __shared__ uint64_t s_seed[4][8];  // [f_in, f_out(rol k), r_in(rol k-1), r_out(ror 1)][class 0..3, 4..7 = 0]
    __shared__ ulonglong2 s_in[8], s_out[8];  // {f_in, r_in}[class], {f_out, r_out}[class]: one load per base
    __shared__ uint64_t s_stage[FILTER ? (KM_THREADS / 32) * KM_STAGE : 1];

    const int tid = threadIdx.x;
    const unsigned lane = lane_id();
    const size_t B0 = (size_t)blockIdx.x * (FILTER ? (size_t)p.tile_stride : (size_t)1) * KM_TILE;
    const size_t tile_end = (B0 + KM_TILE + KM_HALO < p.n_bases) ? B0 + KM_TILE + KM_HALO : p.n_bases;
    const int tile_len = (int)(tile_end - B0);
    build_lut(s_lut);
    if (HASHED) {  // ntHash: every byte that is not A C G T(U) contributes 0 -> class 4
        for (int i = tid; i < 256; i += KM_THREADS) s_lut[i] = s_lut[i] < 4 ? s_lut[i] : 4;
    }
    if (HASHED && tid < 32) {
        const uint64_t S[4] = {0x3c8bfbb395c60474ull, 0x3193c18562a02b4cull, 0x20323ed082572324ull, 0x295549f54be24456ull};
        int t = tid >> 3, c = tid & 7;
        uint64_t v = 0;
        if (c < 4) {
            uint64_t sf = S[c], sr = S[3 - c];
            v = t == 0 ? sf : t == 1 ? rol64d(sf, (unsigned)p.k) : t == 2 ? rol64d(sr, (unsigned)(p.k - 1)) : ror64d(sr, 1);
        }
        s_seed[t][c] = v;
    }
    __syncthreads();
    for (int i = tid; i < tile_len; i += KM_THREADS) s_b[i] = s_lut[p.bases[B0 + i]];
    if (HASHED && tid < 8) {
        s_in[tid] = make_ulonglong2(s_seed[0][tid], s_seed[2][tid]);
        s_out[tid] = make_ulonglong2(s_seed[1][tid], s_seed[3][tid]);
    }
    __syncthreads();

    const size_t b0 = B0 + (size_t)tid * KM_CHUNK;
    size_t b_end = b0 + KM_CHUNK;
    if (b_end > B0 + KM_TILE) b_end = B0 + KM_TILE;
    if (b_end > p.n_bases) b_end = p.n_bases;
    const bool any = b0 < b_end;

    // record containing b0: last r with rec_off[r] <= b0
    size_t r = 0, rs = 0, re = 0;
    unsigned long long obase = 0;
    if (any) {
        size_t lo = 0, hi = p.n_rec;
        while (lo + 1 < hi) {
            size_t mid = (lo + hi) >> 1;
            if (p.rec_off[mid] <= b0) lo = mid;
            else hi = mid;
        }
        r = lo;
        rs = p.rec_off[r];
        re = p.rec_off[r + 1];
        obase = p.out_off[r];
    }

    const int k = p.k;
    const uint64_t kmask = (k >= 32) ? ~0ull : ((1ull << (2 * k)) - 1);
    uint64_t fw = 0, rv = 0;
    bool have = false;
    bool illegal = false;
    uint64_t* stage = s_stage + (FILTER ? (tid >> 5) * KM_STAGE : 0);
    unsigned staged = 0;  // codes waiting in this warp's ring (warp-uniform)

    // class of the base at absolute index q of the current record (q may run past `re` when circular)
    auto fetch = [&](size_t q) -> uint8_t {
        if (q >= re) q = rs + (q - re);
        return (q >= B0 && q < tile_end) ? s_b[q - B0] : s_lut[p.bases[q]];
    };
    // FILTER: append `n` staged codes (n <= KM_STAGE, warp-uniform) at the global cursor
    auto flush = [&](unsigned n) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(p.cursor, (unsigned long long)n);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (unsigned i = lane; i < n; i += 32)
            if (base + i < p.cap) p.out[base + i] = stage[i];
        __syncwarp();
    };
    // first k-mer of a run from scratch: `cls(i)` = class of its i-th base
    auto init = [&](auto cls) {
        fw = 0;
        rv = 0;
        if (HASHED) {
            for (int i = 0; i < k; ++i) {
                const uint8_t c = cls(i);
                const uint64_t sf = s_seed[0][c];
                const uint64_t sr = c < 4 ? s_seed[0][3 - c] : 0ull;
                fw ^= rol64d(sf, (unsigned)(k - 1 - i));
                rv ^= rol64d(sr, (unsigned)i);
            }
        } else {
            for (int i = 0; i < k; ++i) {
                const uint8_t c = cls(i);
                if (c == 255) illegal = true;
                const uint64_t v = (c >= 4 ? (uint64_t)(c - 4) : (uint64_t)c) & 3u;
                fw = ((fw << 2) | v) & kmask;
                rv = (rv >> 2) | ((v ^ 3u) << (2 * (k - 1)));
            }
        }
    };
    // next k-mer: base of class cin enters, base of class cout leaves
    auto roll = [&](uint8_t cin, uint8_t cout) {
        if (HASHED) {
            const ulonglong2 si = s_in[cin], so = s_out[cout];
            fw = ((fw << 1) | (fw >> 63)) ^ so.x ^ si.x;
            rv = ((rv >> 1) | (rv << 63)) ^ so.y ^ si.y;
        } else {
            if (cin == 255) illegal = true;
            const uint64_t v = (cin >= 4 ? (uint64_t)(cin - 4) : (uint64_t)cin) & 3u;
            fw = ((fw << 2) | v) & kmask;
            rv = (rv >> 2) | ((v ^ 3u) << (2 * (k - 1)));
        }
    };

    // fast path: the whole chunk and its trailing k-1 bases inside one record and inside the staged tile
    const bool fast = any && b_end == b0 + KM_CHUNK && b0 + KM_CHUNK + (size_t)k - 1 <= re && b0 + KM_CHUNK + (size_t)k - 1 <= tile_end;
    const int li = (int)(b0 - B0);  // shared-memory index of the chunk's first base
    uint64_t* const op = FILTER ? nullptr : p.out + (obase + (b0 - rs));
    const bool op_odd = (reinterpret_cast<uintptr_t>(op) & 15u) != 0;  // the first code is the second half of an aligned pair
    uint64_t pend = 0;  // first code of an aligned pair, waiting for its partner

    for (int j = 0; j < KM_CHUNK; ++j) {  // uniform trip count: the FILTER mode votes warp-wide every step
        bool emit = false;
        uint64_t code = 0;
        if (fast) {
            if (j == 0) init([&](int i) { return s_b[li + i]; });
            else roll(s_b[li + j + k - 1], s_b[li + j - 1]);
            code = (p.canonical && rv < fw) ? rv : fw;
            emit = true;
            if (!FILTER) {
                // codes j, j+1 form an aligned 16-byte pair when (j odd) == op_odd ... store pairs, singles at the ends
                const bool second = ((j & 1) != 0) != op_odd;  // this code completes a pair that started at j - 1
                if (second && j > 0) {
                    st_stream_u64x2(reinterpret_cast<ulonglong2*>(op + j - 1), make_ulonglong2(pend, code));
                } else if (!second && j + 1 < KM_CHUNK) {
                    pend = code;
                } else {
                    op[j] = code;  // j == 0 completing a pair that belongs to the thread before, or the last, unpaired code
                }
            }
        } else {
            const size_t b = b0 + (size_t)j;
            if (b < b_end) {
                while (b >= re) {
                    ++r;
                    rs = re;
                    re = p.rec_off[r + 1];
                    obase = p.out_off[r];
                    have = false;
                }
                const size_t L = re - rs;
                if (L < (size_t)k || (!p.circular && b + k > re)) {
                    have = false;
                } else {
                    if (!have) {
                        init([&](int i) { return fetch(b + i); });
                        have = true;
                    } else {
                        roll(fetch(b + k - 1), fetch(b - 1));
                    }
                    code = (p.canonical && rv < fw) ? rv : fw;
                    emit = true;
                    if (!FILTER) p.out[obase + (b - rs)] = code;
                }
            }
        }
        if (FILTER) {
            const bool keep = emit && code >= p.range_lo && code <= p.range_hi;
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) stage[staged + (unsigned)__popc(m & lanemask_lt())] = code;
            staged += (unsigned)__popc(m);
            __syncwarp();
            if (staged >= KM_STAGE - 32) {  // room for one more vote is gone: flush everything
                flush(staged);
                staged = 0;
            }
        }
    }
    if (FILTER && staged) flush(staged);
    if (illegal) atomicExch(p.err, (int)UKM_E_ILLEGAL_BASE);
}

struct LeGen {
    const uint64_t* keys;
    uint64_t max_hash;
    __device__ __forceinline__ uint64_t operator()(size_t i, bool* keep) const {
        uint64_t k = ld_stream_u64(keys + i);
        *keep = k <= max_hash;
        return k;
    }
};

// count -W w: one value per FULL window of w consecutive hashes of a record (sketches.NewMinimizerSketch /
// NextMinimizer, count.go:316-317,358-359): position i holds min(h[i .. i+w-1]); it is kept when it differs from
// the window before it (the dedup map of count.go:434-436 would drop the repeat anyway -- this only spares the
// sort 90% of its input).  Windows never span records.  Neighbouring threads read overlapping hashes: L1 serves them.
struct MinimizerGen {
    const uint64_t* h;
    const unsigned long long* out_off;  // first hash of every record, n_rec + 1 entries
    int n_rec;
    int w;
    int scaled;
    uint64_t max_hash;
    __device__ __forceinline__ uint64_t operator()(size_t i, bool* keep) const {
        int lo = 0, hi = n_rec;  // record of hash i: the last r with out_off[r] <= i
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (out_off[mid] <= i) lo = mid;
            else hi = mid;
        }
        const unsigned long long start = out_off[lo], end = out_off[lo + 1];
        *keep = false;
        if (i + (size_t)w > end) return 0;  // not a full window
        uint64_t m1 = ~0ull;                // min of h[i .. i+w-2], shared with the window before
        for (int j = 0; j + 1 < w; ++j) {
            const uint64_t v = h[i + j];
            m1 = v < m1 ? v : m1;
        }
        const uint64_t last = h[i + w - 1];
        const uint64_t m = m1 < last ? m1 : last;
        bool k = true;
        if (i > start) {
            const uint64_t p = h[i - 1];
            k = m != (p < m1 ? p : m1);
        }
        if (scaled && m > max_hash) k = false;  // count.go:373
        *keep = k;
        return m;
    }
};

// The same selection for windows up to MZ_WMAX, tiled: a CTA stages MZ_TILE + w hashes in shared memory, every
// thread owns MZ_R consecutive positions and builds their window minima from one pass over the w + MZ_R hashes they
// span (the part common to all of its windows once, suffix minima on the left, prefix minima on the right), so a
// position costs (w + MZ_R) / MZ_R shared-memory loads instead of w + 1 global ones.  Compaction as in select_kernel.
constexpr int MZ_THREADS = 256;
constexpr int MZ_R = 9;  // odd: the 8-byte accesses of neighbouring threads fall into different banks
constexpr int MZ_TILE = MZ_THREADS * MZ_R;
constexpr int MZ_WMAX = 4096;

struct MzArgs {
    const uint64_t* h;
    size_t n;
    const unsigned long long* out_off;
    int n_rec;
    int w;
    int scaled;
    uint64_t max_hash;
    uint64_t* out;
    uint64_t* status;
    uint32_t* tile_counter;
    unsigned long long* total_out;
    int num_tiles;
    int* err;
};

__device__ __forceinline__ uint64_t mz_min(uint64_t a, uint64_t b) { return a < b ? a : b; }

__global__ void __launch_bounds__(MZ_THREADS) minimizer_kernel(const MzArgs p) {
    constexpr int NW = MZ_THREADS / 32;
    extern __shared__ __align__(16) unsigned char mz_smem[];
    uint64_t* s_x = reinterpret_cast<uint64_t*>(mz_smem);  // s_x[q] = h[t0 - 1 + q], q < MZ_TILE + w + 1
    uint64_t* s_o = s_x + MZ_TILE + p.w + 1;               // staged output, MZ_TILE
    __shared__ unsigned s_scan[NW + 2];
    __shared__ int s_tile;
    __shared__ unsigned long long s_prefix;
    const int tid = threadIdx.x;
    if (tid == 0) s_tile = (int)atomicAdd(p.tile_counter, 1u);
    __syncthreads();
    const int tile = s_tile;
    const int w = p.w;
    const size_t t0 = (size_t)tile * MZ_TILE;
    const int nx = MZ_TILE + w + 1;
    for (int q = tid; q < nx; q += MZ_THREADS) {
        const size_t g = t0 + (size_t)q;  // hash index + 1
        s_x[q] = (g >= 1 && g - 1 < p.n) ? p.h[g - 1] : ~0ull;
    }
    __syncthreads();
    // W[j] = minimum of the window that starts at position i0 - 1 + j  (s_x[q0 + j .. q0 + j + w - 1])
    const size_t i0 = t0 + (size_t)tid * MZ_R;
    const int q0 = tid * MZ_R;
    uint64_t W[MZ_R + 1];
    if (w > MZ_R) {
        uint64_t mid = ~0ull;  // s_x[q0 + R .. q0 + w - 1]: inside every one of the R + 1 windows
        for (int q = q0 + MZ_R; q <= q0 + w - 1; ++q) mid = mz_min(mid, s_x[q]);
        uint64_t l = ~0ull;
#pragma unroll
        for (int j = MZ_R - 1; j >= 0; --j) {
            l = mz_min(l, s_x[q0 + j]);
            W[j] = mz_min(l, mid);
        }
        W[MZ_R] = mid;
        uint64_t r = ~0ull;
#pragma unroll
        for (int j = 1; j <= MZ_R; ++j) {
            r = mz_min(r, s_x[q0 + w + j - 1]);
            W[j] = mz_min(W[j], r);
        }
    } else {
#pragma unroll
        for (int j = 0; j <= MZ_R; ++j) {
            uint64_t m = ~0ull;
            for (int t = 0; t < w; ++t) m = mz_min(m, s_x[q0 + j + t]);
            W[j] = m;
        }
    }
    // record of position i0: the last r with out_off[r] <= i0
    int r = 0;
    {
        int lo = 0, hi = p.n_rec;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (p.out_off[mid] <= i0) lo = mid;
            else hi = mid;
        }
        r = lo;
    }
    unsigned long long rs = p.out_off[r], re = p.out_off[r + 1];
    unsigned mask = 0;
#pragma unroll
    for (int j = 0; j < MZ_R; ++j) {
        const size_t i = i0 + (size_t)j;
        if (i < p.n) {
            while (i >= re && r + 1 < p.n_rec) {
                ++r;
                rs = re;
                re = p.out_off[r + 1];
            }
            bool keep = i >= rs && i + (size_t)w <= re;          // a full window inside the record
            keep = keep && (i == rs || W[j + 1] != W[j]);        // differs from the window before it
            if (p.scaled && W[j + 1] > p.max_hash) keep = false;  // count.go:373
            if (keep) mask |= 1u << j;
        }
    }
    unsigned tile_total;
    const unsigned off = block_excl_scan_u32<MZ_THREADS>((unsigned)__popc(mask), s_scan, &tile_total);
    unsigned o = off;
#pragma unroll
    for (int j = 0; j < MZ_R; ++j)
        if (mask & (1u << j)) s_o[o++] = W[j + 1];
    const unsigned long long pre = tile_exclusive_prefix(p.status, tile, tile_total, p.err, &s_prefix);
    if (tid == 0 && tile == p.num_tiles - 1) *p.total_out = pre + tile_total;
    for (unsigned i = tid; i < tile_total; i += MZ_THREADS) p.out[pre + i] = s_o[i];
}

int minimizer_select(ukm_ctx* ctx, const MinimizerGen& g, size_t n, uint64_t* d_out, size_t* n_out) {
    *n_out = 0;
    if (n == 0) return UKM_OK;
    if (g.w > MZ_WMAX) return ukm_dev_select(ctx, g, n, d_out, n_out, "minimizer_window", 8.0 * (double)n);
    const int num_tiles = (int)((n + MZ_TILE - 1) / MZ_TILE);
    const size_t smem = ((size_t)2 * MZ_TILE + (size_t)g.w + 1) * sizeof(uint64_t);
    UKM_TRY(ukm_kernel_config(ctx, minimizer_kernel, ((size_t)2 * MZ_TILE + MZ_WMAX + 1) * sizeof(uint64_t), 0, nullptr));
    ukm_tmp tmp(ctx);
    uint64_t* d_status = nullptr;
    UKM_TRY(tmp.alloc(&d_status, (size_t)num_tiles + 4));
    MzArgs a;
    a.h = g.h; a.n = n; a.out_off = g.out_off; a.n_rec = g.n_rec; a.w = g.w; a.scaled = g.scaled; a.max_hash = g.max_hash;
    a.out = d_out;
    a.status = d_status;
    a.tile_counter = reinterpret_cast<uint32_t*>(d_status + num_tiles);
    a.total_out = reinterpret_cast<unsigned long long*>(d_status + num_tiles + 1);
    a.num_tiles = num_tiles;
    a.err = ctx->d_err;
    UKM_CUDA(ctx, cudaMemsetAsync(d_status, 0, ((size_t)num_tiles + 4) * sizeof(uint64_t), ctx->stream));
    {
        ukm_stat_scope st(ctx, "minimizer_window", 8.0 * (double)n);
        minimizer_kernel<<<num_tiles, MZ_THREADS, smem, ctx->stream>>>(a);
        UKM_LAUNCHED(ctx);
    }
    UKM_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch, a.total_out, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    UKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_out = (size_t)ctx->h_scratch[0];
    if (ctx->stats_on && !ctx->pending.empty()) ctx->pending.back().bytes += 8.0 * (double)*n_out;
    return UKM_OK;
}

// sequences staged on the device + the per-record output offsets
struct Prepared {
    KmerArgs a;
    unsigned long long total = 0;  // k-mers of all records
    size_t n_bases = 0;
    bool hashed = false;
};

int prepare(ukm_ctx* ctx, ukm_tmp& tmp, const uint8_t* bases, const uint64_t* rec_off, size_t n_rec, int k, unsigned flags, int where,
            Prepared* P, const char* what) {
    const bool hashed = (flags & UKM_F_HASHED) != 0;
    if (!rec_off) return ukm_fail(ctx, UKM_E_ARG, "%s: rec_off == NULL", what);
    if (k < 1 || (!hashed && k > 32) || k > 64) return ukm_fail(ctx, UKM_E_ARG, "%s: k=%d out of range", what, k);
    // the record table is a HOST array for UKM_HOST*; for UKM_DEVICE it lives on the device
    std::vector<unsigned long long> h_rec(n_rec + 1), h_out(n_rec + 1);
    if (where == UKM_DEVICE) {
        UKM_CUDA(ctx, cudaMemcpyAsync(h_rec.data(), rec_off, (n_rec + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
        UKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    } else {
        for (size_t i = 0; i <= n_rec; ++i) h_rec[i] = rec_off[i];
    }
    unsigned long long total = 0;
    for (size_t r = 0; r < n_rec; ++r) {
        if (h_rec[r + 1] < h_rec[r]) return ukm_fail(ctx, UKM_E_ARG, "%s: rec_off not monotone at %zu", what, r);
        unsigned long long L = h_rec[r + 1] - h_rec[r];
        h_out[r] = total;
        if (L >= (unsigned long long)k) total += (flags & UKM_F_CIRCULAR) ? L : L - k + 1;  // count.go:324-328: short records skipped
    }
    h_out[n_rec] = total;
    P->total = total;
    P->n_bases = n_rec ? (size_t)h_rec[n_rec] : 0;
    P->hashed = hashed;
    if (total == 0) return UKM_OK;
    if (n_rec && h_rec[0] != 0) return ukm_fail(ctx, UKM_E_ARG, "%s: rec_off[0] must be 0", what);
    if (!bases) return ukm_fail(ctx, UKM_E_ARG, "%s: bases == NULL", what);
    const uint8_t* d_bases = bases;
    if (where != UKM_DEVICE) {
        uint8_t* t;
        UKM_TRY(tmp.alloc(&t, P->n_bases + 16));
        UKM_CUDA(ctx, cudaMemcpyAsync(t, bases, P->n_bases, cudaMemcpyHostToDevice, ctx->stream));
        d_bases = t;
    }
    unsigned long long *d_rec, *d_out;
    UKM_TRY(tmp.alloc(&d_rec, n_rec + 1));
    UKM_TRY(tmp.alloc(&d_out, n_rec + 1));
    UKM_CUDA(ctx, cudaMemcpyAsync(d_rec, h_rec.data(), (n_rec + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    UKM_CUDA(ctx, cudaMemcpyAsync(d_out, h_out.data(), (n_rec + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    // pageable host sources (and the local vectors) must stay valid until the copies ran
    UKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    KmerArgs& a = P->a;
    memset(&a, 0, sizeof a);
    a.bases = d_bases;
    a.n_bases = P->n_bases;
    a.rec_off = d_rec;
    a.out_off = d_out;
    a.n_rec = n_rec;
    a.k = k;
    a.canonical = (flags & UKM_F_CANONICAL) ? 1 : 0;
    a.circular = (flags & UKM_F_CIRCULAR) ? 1 : 0;
    a.err = ctx->d_err;
    a.tile_stride = 1;
    return UKM_OK;
}

// every code of every record, record-then-position order (the iterators themselves)
int generate_ordered(ukm_ctx* ctx, Prepared& P, uint64_t* d_codes) {
    KmerArgs a = P.a;
    a.out = d_codes;
    const int grid = (int)((P.n_bases + KM_TILE - 1) / KM_TILE);
    ukm_stat_scope st(ctx, P.hashed ? "kmer_nthash" : "kmer_encode", (double)P.n_bases + 8.0 * (double)P.total);
    if (P.hashed) kmer_kernel<true, false><<<grid, KM_THREADS, 0, ctx->stream>>>(a);
    else kmer_kernel<false, false><<<grid, KM_THREADS, 0, ctx->stream>>>(a);
    UKM_LAUNCHED(ctx);
    return UKM_OK;
}

// codes inside [lo, hi] only, arbitrary order; *n_kept may exceed cap (then the caller retries with narrower ranges)
int generate_range(ukm_ctx* ctx, Prepared& P, uint64_t lo, uint64_t hi, uint64_t* d_codes, size_t cap, unsigned long long* d_cursor,
                   size_t* n_kept, unsigned tile_stride = 1) {
    KmerArgs a = P.a;
    a.out = d_codes;
    a.range_lo = lo;
    a.range_hi = hi;
    a.cursor = d_cursor;
    a.cap = cap;
    a.tile_stride = tile_stride;
    UKM_CUDA(ctx, cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long), ctx->stream));
    const size_t n_tiles = (P.n_bases + KM_TILE - 1) / KM_TILE;
    const int grid = (int)((n_tiles + tile_stride - 1) / tile_stride);
    {
        ukm_stat_scope st(ctx, P.hashed ? "kmer_nthash_range" : "kmer_encode_range", (double)P.n_bases);
        if (P.hashed) kmer_kernel<true, true><<<grid, KM_THREADS, 0, ctx->stream>>>(a);
        else kmer_kernel<false, true><<<grid, KM_THREADS, 0, ctx->stream>>>(a);
        UKM_LAUNCHED(ctx);
    }
    UKM_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch, d_cursor, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    UKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_kept = (size_t)ctx->h_scratch[0];
    if (ctx->stats_on && !ctx->pending.empty()) ctx->pending.back().bytes += 8.0 * (double)std::min(*n_kept, cap);
    return UKM_OK;
}

// k-mers per key-range pass of `count` (UKM_COUNT_PASS overrides: tests force several passes on small inputs)
size_t count_pass_limit() {
    const char* e = getenv("UKM_COUNT_PASS");
    if (e) {
        long long v = atoll(e);
        if (v > 0) return (size_t)v;
    }
    return (size_t)2500000000ull;  // 20 GB of codes + 20 GB radix ping-pong per pass
}

}  // namespace

extern "C" int ukm_kmers_seq(ukm_ctx* ctx, const uint8_t* bases, const uint64_t* rec_off, size_t n_rec, int k, unsigned flags,
                             uint64_t max_hash, int where, ukm_span* out) {
    if (!ctx) return UKM_E_ARG;
    if (!out) return ukm_fail(ctx, UKM_E_ARG, "ukm_kmers_seq: out == NULL");
    UKM_TRY(ukm_begin_call(ctx));
    ukm_tmp tmp(ctx);
    Prepared P;
    UKM_TRY(prepare(ctx, tmp, bases, rec_off, n_rec, k, flags, where, &P, "ukm_kmers_seq"));
    if (P.total == 0) return ukm_deliver(ctx, nullptr, nullptr, 0, out);
    uint64_t* d_codes = nullptr;
    UKM_TRY(tmp.alloc(&d_codes, (size_t)P.total + 2));
    UKM_TRY(generate_ordered(ctx, P, d_codes));
    size_t n = (size_t)P.total;
    if (flags & UKM_F_SCALED) {  // count.go:373: `if scaled && code > maxHash { continue }`
        uint64_t* d_kept;
        UKM_TRY(tmp.alloc(&d_kept, n + 2));
        size_t kept = 0;
        LeGen g{d_codes, max_hash};
        UKM_TRY(ukm_dev_select(ctx, g, n, d_kept, &kept, "scaled_filter", 8.0 * (double)n));
        d_codes = d_kept;
        n = kept;
    }
    UKM_TRY(ukm_check_dev_error(ctx, "ukm_kmers_seq"));
    return ukm_deliver(ctx, d_codes, nullptr, n, out);
}

// count = distinct codes, ascending.  The code space is cut into P equal key ranges; each pass regenerates the
// codes of its range (the scaled filter `code <= max_hash` is just a tighter upper bound), sorts them
// (count.go:581) and drops duplicates (the map of count.go:434-436); the passes' results concatenate into the
// globally sorted answer.  P = 1 unless the k-mers would not fit (config C4: 10^10 k-mers = 80 GB of codes).
extern "C" int ukm_count_seq(ukm_ctx* ctx, const uint8_t* bases, const uint64_t* rec_off, size_t n_rec, int k, unsigned flags,
                             uint64_t max_hash, int where, ukm_span* out) {
    if (!ctx) return UKM_E_ARG;
    if (!out) return ukm_fail(ctx, UKM_E_ARG, "ukm_count_seq: out == NULL");
    UKM_TRY(ukm_begin_call(ctx));
    ukm_tmp tmp(ctx);
    Prepared P;
    UKM_TRY(prepare(ctx, tmp, bases, rec_off, n_rec, k, flags, where, &P, "ukm_count_seq"));
    if (P.total == 0) return ukm_deliver(ctx, nullptr, nullptr, 0, out);
    const int key_bits = (flags & UKM_F_HASHED) ? 64 : 2 * k;
    const uint64_t space_max = key_bits == 64 ? ~0ull : ((1ull << key_bits) - 1);
    const uint64_t top = (flags & UKM_F_SCALED) ? std::min(max_hash, space_max) : space_max;  // largest code kept
    // expected codes kept ~ total * (top+1)/2^bits for hashes; for 2-bit codes the estimate is only a start
    const double frac = (flags & UKM_F_HASHED) ? ((double)top + 1.0) / 18446744073709551616.0 : 1.0;
    const size_t limit = count_pass_limit();
    size_t passes = (size_t)(((double)P.total * frac + (double)limit - 1) / (double)limit);
    if (passes < 1) passes = 1;
    unsigned long long* d_cursor = nullptr;
    UKM_TRY(tmp.alloc(&d_cursor, 2));
    const bool out_dev = out->where == UKM_DEVICE;
    cudaMemcpyKind kind = out_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    // Codes are not uniform over [0, top] (a canonical hash is the min of two hashes; 2-bit codes follow the base
    // composition), so the pass boundaries are quantiles of a sample: every `stride`-th tile of start positions
    // (~4M codes), sorted on the device.
    std::vector<uint64_t> sample;
    double sample_frac = 1.0;  // share of the sampled codes that are <= top
    const bool small = P.total <= ((size_t)32 << 20) && passes == 1;  // everything fits in one pass with cap = total
    if (!small) {
        const size_t n_tiles = (P.n_bases + KM_TILE - 1) / KM_TILE;
        const unsigned stride = (unsigned)std::max<size_t>(1, n_tiles / 128);
        const size_t scap = (n_tiles / stride + 2) * (size_t)KM_TILE;
        uint64_t* d_s = nullptr;
        UKM_TRY(tmp.alloc(&d_s, scap + 2));
        size_t ns = 0;
        UKM_TRY(generate_range(ctx, P, 0, top, d_s, scap, d_cursor, &ns, stride));
        if (ns > scap) ns = scap;
        if (ns > 1) UKM_TRY(ukm_dev_sort(ctx, d_s, nullptr, ns, key_bits));
        sample.resize(ns);
        if (ns) UKM_CUDA(ctx, cudaMemcpyAsync(sample.data(), d_s, ns * 8, cudaMemcpyDeviceToHost, ctx->stream));
        UKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        tmp.free_now(d_s);
        const double sampled_kmers = (double)P.total / (double)stride;
        if (sampled_kmers > 0) sample_frac = std::min(1.0, (double)ns / sampled_kmers);
        passes = (size_t)(((double)P.total * sample_frac + (double)limit - 1) / (double)limit);
        if (passes < 1) passes = 1;
    }
    for (int attempt = 0; attempt < 8; ++attempt, passes *= 2) {
        // per-pass buffers: 1.3x the expected share, retried with twice the passes when a range still overflows
        size_t cap = small ? (size_t)P.total : (size_t)((double)P.total * sample_frac / (double)passes * 1.3) + (1u << 20);
        if (cap > P.total) cap = (size_t)P.total;
        uint64_t *d_codes = nullptr, *d_uniq = nullptr;  // d_uniq: staging of a pass's distinct codes, only when they cannot be folded in place
        UKM_TRY(tmp.alloc(&d_codes, cap + 2));
        size_t written = 0;
        bool overflow = false;
        uint64_t lo = 0;
        for (size_t ps = 0; ps < passes && !overflow; ++ps) {
            // range ps = (quantile ps/passes, quantile (ps+1)/passes] of the sample; the last range ends at `top`
            uint64_t hi = top;
            if (ps + 1 < passes) {
                if (!sample.empty()) hi = sample[std::min(sample.size() - 1, (size_t)((double)sample.size() * (double)(ps + 1) / (double)passes))];
                else hi = (uint64_t)(((long double)top + 1.0L) / (long double)passes * (long double)(ps + 1)) - 1;
                if (hi > top) hi = top;
            }
            if (ps > 0 && hi < lo) continue;  // empty range (many equal sample values)
            size_t kept = 0;
            UKM_TRY(generate_range(ctx, P, lo, hi, d_codes, cap, d_cursor, &kept));
            UKM_TRY(ukm_check_dev_error(ctx, "ukm_count_seq"));
            if (kept > cap) {
                overflow = true;
                break;
            }
            if (kept) {
                UKM_TRY(ukm_dev_sort(ctx, d_codes, nullptr, kept, key_bits));
                size_t m = 0;
                if (out_dev && written + kept <= out->cap) {
                    // a device result is folded straight into place (no staging copy of the pass's 20 GB)
                    UKM_TRY(ukm_dev_fold(ctx, UKM_FOLD_UNIQUE, d_codes, nullptr, kept, false, out->keys + written, nullptr, &m));
                } else {
                    if (!d_uniq) UKM_TRY(tmp.alloc(&d_uniq, cap + 2));
                    UKM_TRY(ukm_dev_fold(ctx, UKM_FOLD_UNIQUE, d_codes, nullptr, kept, false, d_uniq, nullptr, &m));
                    if (written + m > out->cap) {
                        out->n = written + m;
                        return ukm_fail(ctx, UKM_E_CAPACITY, "ukm_count_seq: output needs more than %zu elements", out->cap);
                    }
                    if (m) UKM_CUDA(ctx, cudaMemcpyAsync(out->keys + written, d_uniq, m * sizeof(uint64_t), kind, ctx->stream));
                    UKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                }
                written += m;
            }
            if (hi == top) break;
            lo = hi + 1;
        }
        tmp.free_now(d_codes);
        if (d_uniq) tmp.free_now(d_uniq);
        if (!overflow) {
            UKM_TRY(ukm_check_dev_error(ctx, "ukm_count_seq"));
            out->n = written;
            return UKM_OK;
        }
    }
    return ukm_fail(ctx, UKM_E_INTERNAL, "ukm_count_seq: key ranges too skewed for the pass buffers");
}

// count -H -W w (count.go:100-114, 316-317, 358-359): the distinct sliding-window minima of the ntHash stream of
// every record, ascending.  hashes -> window minima that differ from their predecessor -> sort -> unique.
extern "C" int ukm_count_minimizer(ukm_ctx* ctx, const uint8_t* bases, const uint64_t* rec_off, size_t n_rec, int k, int w,
                                   unsigned flags, uint64_t max_hash, int where, ukm_span* out) {
    if (!ctx) return UKM_E_ARG;
    if (!out) return ukm_fail(ctx, UKM_E_ARG, "ukm_count_minimizer: out == NULL");
    if (w < 1) return ukm_fail(ctx, UKM_E_ARG, "ukm_count_minimizer: w=%d", w);
    UKM_TRY(ukm_begin_call(ctx));
    flags |= UKM_F_HASHED;  // count.go:105-109: -W switches -H on
    ukm_tmp tmp(ctx);
    Prepared P;
    UKM_TRY(prepare(ctx, tmp, bases, rec_off, n_rec, k, flags, where, &P, "ukm_count_minimizer"));
    if (P.total == 0) return ukm_deliver(ctx, nullptr, nullptr, 0, out);
    const size_t n = (size_t)P.total;
    uint64_t *d_h = nullptr, *d_c = nullptr;
    UKM_TRY(tmp.alloc(&d_h, n + 2));
    UKM_TRY(generate_ordered(ctx, P, d_h));
    UKM_TRY(tmp.alloc(&d_c, n + 2));
    size_t nc = 0;
    MinimizerGen g{d_h, P.a.out_off, (int)n_rec, w, (flags & UKM_F_SCALED) ? 1 : 0, max_hash};
    UKM_TRY(minimizer_select(ctx, g, n, d_c, &nc));
    tmp.free_now(d_h);
    UKM_TRY(ukm_check_dev_error(ctx, "ukm_count_minimizer"));
    if (nc == 0) return ukm_deliver(ctx, nullptr, nullptr, 0, out);
    UKM_TRY(ukm_dev_sort(ctx, d_c, nullptr, nc, 64));
    uint64_t* d_u = nullptr;
    UKM_TRY(tmp.alloc(&d_u, nc + 2));
    size_t m = 0;
    UKM_TRY(ukm_dev_fold(ctx, UKM_FOLD_UNIQUE, d_c, nullptr, nc, false, d_u, nullptr, &m));
    UKM_TRY(ukm_check_dev_error(ctx, "ukm_count_minimizer"));
    return ukm_deliver(ctx, d_u, nullptr, m, out);
}
