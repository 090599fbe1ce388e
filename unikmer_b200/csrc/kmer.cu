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
};

template <bool HASHED>
__global__ void __launch_bounds__(KM_THREADS) kmer_kernel(const KmerArgs p) {
    __shared__ uint8_t s_b[KM_TILE + KM_HALO + 16];
    __shared__ uint8_t s_lut[256];
    __shared__ uint64_t s_seed[4][8];  // [f_in, f_out(rol k), r_in(rol k-1), r_out(ror 1)][class 0..3, 4..7 = 0]

    const int tid = threadIdx.x;
    const size_t B0 = (size_t)blockIdx.x * KM_TILE;
    const size_t tile_end = (B0 + KM_TILE + KM_HALO < p.n_bases) ? B0 + KM_TILE + KM_HALO : p.n_bases;
    const int tile_len = (int)(tile_end - B0);
    build_lut(s_lut);
    for (int i = tid; i < tile_len; i += KM_THREADS) s_b[i] = p.bases[B0 + i];
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

    size_t b = B0 + (size_t)tid * KM_CHUNK;
    size_t b_end = b + KM_CHUNK;
    if (b_end > B0 + KM_TILE) b_end = B0 + KM_TILE;
    if (b_end > p.n_bases) b_end = p.n_bases;
    if (b >= b_end) return;

    // record containing b: last r with rec_off[r] <= b
    size_t lo = 0, hi = p.n_rec;
    while (lo + 1 < hi) {
        size_t mid = (lo + hi) >> 1;
        if (p.rec_off[mid] <= b) lo = mid;
        else hi = mid;
    }
    size_t r = lo;
    size_t rs = p.rec_off[r], re = p.rec_off[r + 1];
    unsigned long long obase = p.out_off[r];

    const int k = p.k;
    const uint64_t kmask = (k >= 32) ? ~0ull : ((1ull << (2 * k)) - 1);
    uint64_t fw = 0, rv = 0;
    bool have = false;
    bool illegal = false;

    // base at absolute index q of the current record (q may run past `re` when circular)
    auto fetch = [&](size_t q) -> uint8_t {
        if (q >= re) q = rs + (q - re);
        return (q >= B0 && q < tile_end) ? s_b[q - B0] : p.bases[q];
    };

    for (; b < b_end; ++b) {
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
            continue;
        }
        if (!have) {
            fw = 0;
            rv = 0;
            if (HASHED) {
                for (int i = 0; i < k; ++i) {
                    uint8_t c = s_lut[fetch(b + i)];
                    uint64_t sf = c < 4 ? s_seed[0][c] : 0ull;
                    uint64_t sr = c < 4 ? s_seed[0][3 - c] : 0ull;
                    fw ^= rol64d(sf, (unsigned)(k - 1 - i));
                    rv ^= rol64d(sr, (unsigned)i);
                }
            } else {
                for (int i = 0; i < k; ++i) {
                    uint8_t c = s_lut[fetch(b + i)];
                    if (c == 255) illegal = true;
                    uint64_t v = (c >= 4 ? (uint64_t)(c - 4) : (uint64_t)c) & 3u;
                    fw = ((fw << 2) | v) & kmask;
                    rv = (rv >> 2) | ((v ^ 3u) << (2 * (k - 1)));
                }
            }
            have = true;
        } else {
            uint8_t cin = s_lut[fetch(b + k - 1)];
            if (HASHED) {
                uint8_t cout = s_lut[fetch(b - 1)];
                uint64_t fi = cin < 4 ? s_seed[0][cin] : 0ull, fo = cout < 4 ? s_seed[1][cout] : 0ull;
                uint64_t ri = cin < 4 ? s_seed[2][cin] : 0ull, ro = cout < 4 ? s_seed[3][cout] : 0ull;
                fw = rol64d(fw, 1) ^ fo ^ fi;
                rv = ror64d(rv, 1) ^ ro ^ ri;
            } else {
                if (cin == 255) illegal = true;
                uint64_t v = (cin >= 4 ? (uint64_t)(cin - 4) : (uint64_t)cin) & 3u;
                fw = ((fw << 2) | v) & kmask;
                rv = (rv >> 2) | ((v ^ 3u) << (2 * (k - 1)));
            }
        }
        p.out[obase + (b - rs)] = (p.canonical && rv < fw) ? rv : fw;
    }
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

// generate every k-mer code / hash of every record into a fresh device buffer
int generate(ukm_ctx* ctx, ukm_tmp& tmp, const uint8_t* bases, const uint64_t* rec_off, size_t n_rec, int k, unsigned flags,
             uint64_t max_hash, int where, uint64_t** d_codes, size_t* n_codes, const char* what) {
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
    const size_t n_bases = n_rec ? (size_t)h_rec[n_rec] : 0;
    *n_codes = (size_t)total;
    *d_codes = nullptr;
    if (total == 0) return UKM_OK;
    if (n_rec && h_rec[0] != 0) return ukm_fail(ctx, UKM_E_ARG, "%s: rec_off[0] must be 0", what);
    if (!bases) return ukm_fail(ctx, UKM_E_ARG, "%s: bases == NULL", what);

    const uint8_t* d_bases = bases;
    if (where != UKM_DEVICE) {
        uint8_t* t;
        UKM_TRY(tmp.alloc(&t, n_bases + 16));
        UKM_CUDA(ctx, cudaMemcpyAsync(t, bases, n_bases, cudaMemcpyHostToDevice, ctx->stream));
        d_bases = t;
    }
    unsigned long long *d_rec, *d_out;
    UKM_TRY(tmp.alloc(&d_rec, n_rec + 1));
    UKM_TRY(tmp.alloc(&d_out, n_rec + 1));
    UKM_CUDA(ctx, cudaMemcpyAsync(d_rec, h_rec.data(), (n_rec + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    UKM_CUDA(ctx, cudaMemcpyAsync(d_out, h_out.data(), (n_rec + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    UKM_TRY(tmp.alloc(d_codes, (size_t)total + 2));

    KmerArgs a;
    a.bases = d_bases;
    a.n_bases = n_bases;
    a.rec_off = d_rec;
    a.out_off = d_out;
    a.n_rec = n_rec;
    a.k = k;
    a.canonical = (flags & UKM_F_CANONICAL) ? 1 : 0;
    a.circular = (flags & UKM_F_CIRCULAR) ? 1 : 0;
    a.out = *d_codes;
    a.err = ctx->d_err;
    const int grid = (int)((n_bases + KM_TILE - 1) / KM_TILE);
    {
        ukm_stat_scope st(ctx, hashed ? "kmer_nthash" : "kmer_encode", (double)n_bases + 8.0 * (double)total);
        if (hashed) kmer_kernel<true><<<grid, KM_THREADS, 0, ctx->stream>>>(a);
        else kmer_kernel<false><<<grid, KM_THREADS, 0, ctx->stream>>>(a);
        UKM_LAUNCHED(ctx);
    }
    // pageable host sources must stay valid until the copies above ran
    UKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));

    if (flags & UKM_F_SCALED) {  // count.go:373: `if scaled && code > maxHash { continue }`
        uint64_t* d_kept;
        UKM_TRY(tmp.alloc(&d_kept, (size_t)total + 2));
        size_t kept = 0;
        LeGen g{*d_codes, max_hash};
        UKM_TRY(ukm_dev_select(ctx, g, (size_t)total, d_kept, &kept, "scaled_filter", 8.0 * (double)total));
        tmp.free_now(*d_codes);
        *d_codes = d_kept;
        *n_codes = kept;
    }
    return UKM_OK;
}

}  // namespace

extern "C" int ukm_kmers_seq(ukm_ctx* ctx, const uint8_t* bases, const uint64_t* rec_off, size_t n_rec, int k, unsigned flags,
                             uint64_t max_hash, int where, ukm_span* out) {
    if (!ctx) return UKM_E_ARG;
    if (!out) return ukm_fail(ctx, UKM_E_ARG, "ukm_kmers_seq: out == NULL");
    UKM_CUDA(ctx, cudaSetDevice(ctx->device));
    ukm_tmp tmp(ctx);
    uint64_t* d_codes = nullptr;
    size_t n = 0;
    UKM_TRY(generate(ctx, tmp, bases, rec_off, n_rec, k, flags, max_hash, where, &d_codes, &n, "ukm_kmers_seq"));
    UKM_TRY(ukm_check_dev_error(ctx, "ukm_kmers_seq"));
    return ukm_deliver(ctx, d_codes, nullptr, n, out);
}

extern "C" int ukm_count_seq(ukm_ctx* ctx, const uint8_t* bases, const uint64_t* rec_off, size_t n_rec, int k, unsigned flags,
                             uint64_t max_hash, int where, ukm_span* out) {
    if (!ctx) return UKM_E_ARG;
    if (!out) return ukm_fail(ctx, UKM_E_ARG, "ukm_count_seq: out == NULL");
    UKM_CUDA(ctx, cudaSetDevice(ctx->device));
    ukm_tmp tmp(ctx);
    uint64_t* d_codes = nullptr;
    size_t n = 0;
    UKM_TRY(generate(ctx, tmp, bases, rec_off, n_rec, k, flags, max_hash, where, &d_codes, &n, "ukm_count_seq"));
    UKM_TRY(ukm_check_dev_error(ctx, "ukm_count_seq"));
    if (n == 0) return ukm_deliver(ctx, nullptr, nullptr, 0, out);
    const int key_bits = (flags & UKM_F_HASHED) ? 64 : 2 * k;
    UKM_TRY(ukm_dev_sort(ctx, d_codes, nullptr, n, key_bits));  // count.go:581
    uint64_t* d_uniq = nullptr;
    UKM_TRY(tmp.alloc(&d_uniq, n + 2));
    size_t m = 0;
    UKM_TRY(ukm_dev_fold(ctx, UKM_FOLD_UNIQUE, d_codes, nullptr, n, false, d_uniq, nullptr, &m));  // the map of count.go:434-436
    UKM_TRY(ukm_check_dev_error(ctx, "ukm_count_seq"));
    return ukm_deliver(ctx, d_uniq, nullptr, m, out);
}
