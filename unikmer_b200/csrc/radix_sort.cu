// radix_sort.cu -- LSB radix sort of uint64 k-mer codes (optionally with uint32 taxids).
//
// Replaces sortutil.Uint64s (sort.go:274,337,463; union.go:274,295; diff.go:587;
// common.go:344; count.go:581; split.go:311,383) and sorts.Quicksort(CodeTaxidSlice)
// (sort.go:268,331,457; split.go:305; record layout kmers.go:24-46).
//
// Algorithm (B200-first, not the reference's MSD/quick sorts): one histogram sweep over
// the keys for all digit positions, then one "onesweep" pass per 8-bit digit: every CTA
// ranks a tile of keys with warp match + per-warp shared-memory counters, chains its
// per-digit counts to its predecessors with a decoupled look-back (no separate scan pass,
// keys are read once and written once per pass), reorders the tile in shared memory and
// writes digit-contiguous runs to HBM.  Traffic: (1 + 2*passes) * 8 B per key.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int MAX_PASSES = 8;
constexpr size_t PORTION = (size_t)1 << 29;  // keys per look-back chain (30-bit status values)

struct PassPlan {
    int npass;
    int shift[MAX_PASSES];
    uint32_t mask[MAX_PASSES];
};

PassPlan make_plan(int key_bits) {
    if (key_bits <= 0 || key_bits > 64) key_bits = 64;
    PassPlan p{};
    p.npass = (key_bits + RADIX_BITS - 1) / RADIX_BITS;
    int left = key_bits;
    for (int i = 0; i < p.npass; ++i) {
        int b = left >= RADIX_BITS ? RADIX_BITS : left;
        p.shift[i] = i * RADIX_BITS;
        p.mask[i] = (1u << b) - 1;
        left -= b;
    }
    return p;
}

// ---- histogram of every digit position in one sweep --------------------------------
constexpr int HIST_THREADS = 512;

__global__ void __launch_bounds__(HIST_THREADS) radix_hist_kernel(const uint64_t* __restrict__ keys, size_t n, PassPlan plan,
                                                                   unsigned long long* __restrict__ ghist) {
    __shared__ uint32_t sh[MAX_PASSES * RADIX];
    for (int i = threadIdx.x; i < MAX_PASSES * RADIX; i += HIST_THREADS) sh[i] = 0;
    __syncthreads();
    // 128-bit loads over the 16-byte aligned body, scalar head/tail
    const size_t head = (n && ((reinterpret_cast<uintptr_t>(keys) >> 3) & 1)) ? 1 : 0;
    const size_t n2 = (n - head) / 2;
    const ulonglong2* k2 = reinterpret_cast<const ulonglong2*>(keys + head);
    size_t stride = (size_t)gridDim.x * HIST_THREADS;
    for (size_t i = (size_t)blockIdx.x * HIST_THREADS + threadIdx.x; i < n2; i += stride) {
        ulonglong2 v = ld_stream_u64x2(k2 + i);
#pragma unroll
        for (int p = 0; p < MAX_PASSES; ++p) {
            if (p < plan.npass) {
                atomicAdd(&sh[p * RADIX + ((uint32_t)(v.x >> plan.shift[p]) & plan.mask[p])], 1u);
                atomicAdd(&sh[p * RADIX + ((uint32_t)(v.y >> plan.shift[p]) & plan.mask[p])], 1u);
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < 2) {
        // at most one unaligned head element and one odd tail element
        size_t idx = threadIdx.x == 0 ? 0 : head + 2 * n2;
        bool take = threadIdx.x == 0 ? (head == 1) : (idx < n);
        if (take) {
            uint64_t v = keys[idx];
            for (int p = 0; p < plan.npass; ++p) atomicAdd(&sh[p * RADIX + ((uint32_t)(v >> plan.shift[p]) & plan.mask[p])], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < plan.npass * RADIX; i += HIST_THREADS) {
        uint32_t c = sh[i];
        if (c) atomicAdd(&ghist[i], (unsigned long long)c);
    }
}

// exclusive scan of each pass's 256 bins -> global digit bases.  One block per pass.
__global__ void __launch_bounds__(RADIX) radix_bases_kernel(const unsigned long long* __restrict__ ghist,
                                                            unsigned long long* __restrict__ bases) {
    __shared__ unsigned long long s[RADIX];
    int p = blockIdx.x, d = threadIdx.x;
    s[d] = ghist[p * RADIX + d];
    __syncthreads();
    if (d == 0) {
        unsigned long long acc = 0;
        for (int i = 0; i < RADIX; ++i) {
            unsigned long long c = s[i];
            s[i] = acc;
            acc += c;
        }
    }
    __syncthreads();
    bases[p * RADIX + d] = s[d];
}

// ---- onesweep pass ---------------------------------------------------------------------
#define OS_FLAG_PARTIAL (1u << 30)
#define OS_FLAG_INCLUSIVE (2u << 30)
#define OS_VALUE_MASK ((1u << 30) - 1)

template <int THREADS, int ITEMS, bool PAIRS, bool BALLOT_MATCH>
__global__ void __launch_bounds__(THREADS)
    onesweep_kernel(const uint64_t* __restrict__ kin, uint64_t* __restrict__ kout, const uint32_t* __restrict__ vin,
                    uint32_t* __restrict__ vout, size_t n, int num_tiles, int shift, uint32_t mask,
                    const unsigned long long* __restrict__ bases_in, unsigned long long* __restrict__ bases_out,
                    uint32_t* __restrict__ status, uint32_t* __restrict__ tile_counter, int* __restrict__ err) {
    constexpr int WARPS = THREADS / 32;
    constexpr int TILE = THREADS * ITEMS;
    static_assert(THREADS >= RADIX, "one thread per digit needed");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* s_keys = reinterpret_cast<uint64_t*>(smem_raw);                       // TILE
    unsigned long long* s_gbase = reinterpret_cast<unsigned long long*>(s_keys + TILE);  // RADIX
    uint32_t* s_whist = reinterpret_cast<uint32_t*>(s_gbase + RADIX);                // WARPS*RADIX
    uint32_t* s_dstart = s_whist + WARPS * RADIX;                                    // RADIX
    uint32_t* s_scan = s_dstart + RADIX;                                             // WARPS+2
    uint32_t* s_vals = s_scan + WARPS + 2;                                           // TILE (PAIRS)
    __shared__ int s_tile;

    const int tid = threadIdx.x;
    const unsigned lane = lane_id();
    const int warp = tid >> 5;

    if (tid == 0) s_tile = (int)atomicAdd(tile_counter, 1u);
    for (int i = tid; i < WARPS * RADIX; i += THREADS) s_whist[i] = 0;
    __syncthreads();
    const int tile = s_tile;
    const size_t tile_base = (size_t)tile * TILE;
    const int valid = (n - tile_base) < (size_t)TILE ? (int)(n - tile_base) : TILE;

    // 1. load, warp-striped (coalesced 256 B per warp request)
    uint64_t key[ITEMS];
    uint32_t val[PAIRS ? ITEMS : 1];
    {
        const int wofs = warp * 32 * ITEMS + (int)lane;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            int t = wofs + i * 32;
            key[i] = (t < valid) ? ld_stream_u64(kin + tile_base + t) : ~0ull;
        }
        if (PAIRS) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                int t = wofs + i * 32;
                val[i] = (t < valid) ? __ldg(vin + tile_base + t) : 0u;
            }
        }
    }

    // 2. rank inside the warp.  Phase A: peer masks of all items (independent: MATCH / ballots pipeline
    //    freely).  Phase B: the only serial part, one shared-memory counter update per item: every
    //    lane reads the (warp,digit) counter, the highest peer lane adds the peer count.
    uint32_t* wh = s_whist + warp * RADIX;
    uint32_t rank[ITEMS];
    const unsigned lt = lanemask_lt();
    unsigned peers[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t d = (uint32_t)(key[i] >> shift) & mask;
        if (BALLOT_MATCH) {
            unsigned m = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < RADIX_BITS; ++b) {
                const unsigned v = __ballot_sync(0xffffffffu, (d >> b) & 1u);
                m &= ((d >> b) & 1u) ? v : ~v;
            }
            peers[i] = m;
        } else {
            peers[i] = __match_any_sync(0xffffffffu, d);
        }
    }
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t d = (uint32_t)(key[i] >> shift) & mask;
        const uint32_t before = (uint32_t)__popc(peers[i] & lt);
        const uint32_t cnt = (uint32_t)__popc(peers[i]);
        const uint32_t old = wh[d];
        __syncwarp();
        if (before == cnt - 1) wh[d] = old + cnt;
        __syncwarp();
        rank[i] = old + before;
    }
    __syncthreads();

    // 3. per digit: exclusive offsets across warps, CTA total
    uint32_t bin = 0;
    if (tid < RADIX) {
        uint32_t sum = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            uint32_t c = s_whist[w * RADIX + tid];
            s_whist[w * RADIX + tid] = sum;
            sum += c;
        }
        bin = sum;
    }
    uint32_t total;
    uint32_t dstart = block_excl_scan_u32<THREADS>(bin, s_scan, &total);
    (void)total;

    // 4. decoupled look-back, one thread per digit.  (Publishing the counts early and chaining AFTER the reorder was
    //    measured slower, 72.7 vs 59.9 ms per 1e9 keys: successors then meet more PARTIAL words and walk further.)
    if (tid < RADIX) {
        uint32_t pub = bin;
        if ((uint32_t)tid == mask) pub -= (uint32_t)(TILE - valid);  // padding keys sit in the top digit
        uint32_t* my = status + (size_t)tile * RADIX + tid;
        uint32_t excl = 0;
        if (tile == 0) {
            st_relaxed_u32(my, OS_FLAG_INCLUSIVE | pub);
        } else {
            st_relaxed_u32(my, OS_FLAG_PARTIAL | pub);
            int j = tile - 1;
            unsigned spins = 0;
            while (true) {
                uint32_t w = ld_relaxed_u32(status + (size_t)j * RADIX + tid);
                uint32_t flag = w >> 30;
                if (flag == 0) {
                    if (++spins > UKM_WATCHDOG_SPINS) {
                        atomicExch(err, (int)UKM_E_INTERNAL);
                        break;
                    }
                    continue;
                }
                spins = 0;
                excl += w & OS_VALUE_MASK;
                if (flag == 2 || j == 0) break;
                --j;
            }
            st_relaxed_u32(my, OS_FLAG_INCLUSIVE | (excl + pub));
        }
        unsigned long long gb = bases_in[tid];
        s_gbase[tid] = gb + excl - dstart;
        s_dstart[tid] = dstart;
        if (tile == num_tiles - 1 && bases_out) bases_out[tid] = gb + excl + pub;
    }
    __syncthreads();

    // 5. reorder the tile in shared memory
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        uint32_t d = (uint32_t)(key[i] >> shift) & mask;
        uint32_t pos = s_dstart[d] + wh[d] + rank[i];
        s_keys[pos] = key[i];
        if (PAIRS) s_vals[pos] = val[i];
    }
    __syncthreads();

    // 6. digit-contiguous runs to HBM (consecutive threads -> consecutive addresses inside a run)
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        int p = tid + i * THREADS;
        if (p < valid) {
            uint64_t k = s_keys[p];
            uint32_t d = (uint32_t)(k >> shift) & mask;
            size_t dst = (size_t)(s_gbase[d] + (unsigned long long)p);
            kout[dst] = k;
            if (PAIRS) vout[dst] = s_vals[p];
        }
    }
}

template <int THREADS, int ITEMS, bool PAIRS>
size_t onesweep_smem() {
    constexpr int WARPS = THREADS / 32;
    constexpr int TILE = THREADS * ITEMS;
    size_t s = (size_t)TILE * 8 + RADIX * 8 + (size_t)WARPS * RADIX * 4 + RADIX * 4 + (WARPS + 2) * 4;
    if (PAIRS) s += (size_t)TILE * 4;
    return s + 16;
}

template <int THREADS, int ITEMS, bool PAIRS, bool BM>
int launch_onesweep_v(ukm_ctx* ctx, const uint64_t* kin, uint64_t* kout, const uint32_t* vin, uint32_t* vout, size_t n, int shift,
                    uint32_t mask, const unsigned long long* bases_in, unsigned long long* bases_out, uint32_t* status,
                    uint32_t* counter) {
    constexpr int TILE = THREADS * ITEMS;
    int num_tiles = (int)((n + TILE - 1) / TILE);
    size_t smem = onesweep_smem<THREADS, ITEMS, PAIRS>();
    auto kern = onesweep_kernel<THREADS, ITEMS, PAIRS, BM>;
    UKM_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    UKM_CUDA(ctx, cudaMemsetAsync(status, 0, (size_t)num_tiles * RADIX * sizeof(uint32_t), ctx->stream));
    kern<<<num_tiles, THREADS, smem, ctx->stream>>>(kin, kout, vin, vout, n, num_tiles, shift, mask, bases_in, bases_out, status,
                                                    counter, ctx->d_err);
    UKM_LAUNCHED(ctx);
    return UKM_OK;
}

// UKM_SORT_MATCH=any selects MATCH.ANY instead of the ballot-built peer masks (A/B runs)
template <int THREADS, int ITEMS, bool PAIRS>
int launch_onesweep(ukm_ctx* ctx, const uint64_t* kin, uint64_t* kout, const uint32_t* vin, uint32_t* vout, size_t n, int shift,
                    uint32_t mask, const unsigned long long* bases_in, unsigned long long* bases_out, uint32_t* status,
                    uint32_t* counter) {
    // peer masks from 8 ballots beat MATCH.ANY on B200 (59.8 vs 84.2 ms for 1e9 keys); "any" switches back
    const char* e = getenv("UKM_SORT_MATCH");
    if (!e || e[0] != 'a')
        return launch_onesweep_v<THREADS, ITEMS, PAIRS, true>(ctx, kin, kout, vin, vout, n, shift, mask, bases_in, bases_out, status, counter);
    return launch_onesweep_v<THREADS, ITEMS, PAIRS, false>(ctx, kin, kout, vin, vout, n, shift, mask, bases_in, bases_out, status, counter);
}

struct SortCfg {
    int threads, items;
};

SortCfg pick_cfg(bool pairs) {
    // tunable for A/B runs on the GPU box: UKM_SORT_CFG = 0..3
    static const SortCfg cfgs[] = {{256, 16}, {256, 24}, {512, 16}, {384, 20}};
    const char* e = getenv("UKM_SORT_CFG");
    int i = e ? atoi(e) : 2;  // 512 threads x 16 keys: fastest of the four on B200
    if (i < 0 || i > 3) i = 2;
    (void)pairs;
    return cfgs[i];
}

template <bool PAIRS>
int dispatch_onesweep(ukm_ctx* ctx, SortCfg c, const uint64_t* kin, uint64_t* kout, const uint32_t* vin, uint32_t* vout, size_t n,
                      int shift, uint32_t mask, const unsigned long long* bi, unsigned long long* bo, uint32_t* status,
                      uint32_t* counter) {
    if (c.threads == 256 && c.items == 16)
        return launch_onesweep<256, 16, PAIRS>(ctx, kin, kout, vin, vout, n, shift, mask, bi, bo, status, counter);
    if (c.threads == 256 && c.items == 24)
        return launch_onesweep<256, 24, PAIRS>(ctx, kin, kout, vin, vout, n, shift, mask, bi, bo, status, counter);
    if (c.threads == 512 && c.items == 16)
        return launch_onesweep<512, 16, PAIRS>(ctx, kin, kout, vin, vout, n, shift, mask, bi, bo, status, counter);
    return launch_onesweep<384, 20, PAIRS>(ctx, kin, kout, vin, vout, n, shift, mask, bi, bo, status, counter);
}

// ---- Go []CodeTaxid (16-byte AoS) <-> SoA -----------------------------------------------
__global__ void aos16_split_kernel(const ulonglong2* __restrict__ aos, uint64_t* __restrict__ k, uint32_t* __restrict__ v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        ulonglong2 r = aos[i];
        k[i] = r.x;
        v[i] = (uint32_t)r.y;  // little-endian: taxid is the low half of the second word
    }
}
__global__ void aos16_join_kernel(ulonglong2* __restrict__ aos, const uint64_t* __restrict__ k, const uint32_t* __restrict__ v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) aos[i] = make_ulonglong2(k[i], (unsigned long long)v[i]);
}

}  // namespace

// device-level sort used by every command pipeline; d_vals == nullptr sorts keys only
int ukm_dev_sort(ukm_ctx* ctx, uint64_t* d_keys, uint32_t* d_vals, size_t n, int key_bits) {
    if (n < 2) return UKM_OK;
    const bool pairs = d_vals != nullptr;
    PassPlan plan = make_plan(key_bits);
    SortCfg cfg = pick_cfg(pairs);
    const size_t tile = (size_t)cfg.threads * cfg.items;
    const size_t portion = (PORTION / tile) * tile;
    const size_t nportions = (n + portion - 1) / portion;
    const size_t max_tiles = ((n < portion ? n : portion) + tile - 1) / tile;

    ukm_tmp tmp(ctx);
    uint64_t* d_tmpk = nullptr;
    uint32_t* d_tmpv = nullptr;
    unsigned long long *d_hist = nullptr, *d_bases = nullptr;
    uint32_t *d_status = nullptr, *d_counters = nullptr;
    UKM_TRY(tmp.alloc(&d_tmpk, n));
    if (pairs) UKM_TRY(tmp.alloc(&d_tmpv, n));
    UKM_TRY(tmp.alloc(&d_hist, (size_t)MAX_PASSES * RADIX));
    UKM_TRY(tmp.alloc(&d_bases, (size_t)2 * MAX_PASSES * RADIX));
    UKM_TRY(tmp.alloc(&d_status, max_tiles * RADIX));
    UKM_TRY(tmp.alloc(&d_counters, (size_t)MAX_PASSES * nportions));
    UKM_CUDA(ctx, cudaMemsetAsync(d_hist, 0, (size_t)MAX_PASSES * RADIX * sizeof(unsigned long long), ctx->stream));
    UKM_CUDA(ctx, cudaMemsetAsync(d_counters, 0, (size_t)MAX_PASSES * nportions * sizeof(uint32_t), ctx->stream));

    {
        ukm_stat_scope st(ctx, "radix_hist", 8.0 * (double)n);
        radix_hist_kernel<<<ctx->sm_count * 4, HIST_THREADS, 0, ctx->stream>>>(d_keys, n, plan, d_hist);
        UKM_LAUNCHED(ctx);
        radix_bases_kernel<<<plan.npass, RADIX, 0, ctx->stream>>>(d_hist, d_bases);
        UKM_LAUNCHED(ctx);
    }

    uint64_t *ksrc = d_keys, *kdst = d_tmpk;
    uint32_t *vsrc = d_vals, *vdst = d_tmpv;
    for (int p = 0; p < plan.npass; ++p) {
        unsigned long long* b0 = d_bases + (size_t)p * RADIX;
        unsigned long long* b1 = d_bases + (size_t)(MAX_PASSES + p) * RADIX;
        for (size_t q = 0; q < nportions; ++q) {
            size_t off = q * portion;
            size_t cnt = (n - off) < portion ? (n - off) : portion;
            ukm_stat_scope st(ctx, pairs ? "onesweep_pairs" : "onesweep_keys", (pairs ? 24.0 : 16.0) * (double)cnt);
            int r;
            if (pairs)
                r = dispatch_onesweep<true>(ctx, cfg, ksrc + off, kdst, vsrc + off, vdst, cnt, plan.shift[p], plan.mask[p], b0,
                                            (q + 1 < nportions) ? b1 : nullptr, d_status, d_counters + (size_t)p * nportions + q);
            else
                r = dispatch_onesweep<false>(ctx, cfg, ksrc + off, kdst, nullptr, nullptr, cnt, plan.shift[p], plan.mask[p], b0,
                                             (q + 1 < nportions) ? b1 : nullptr, d_status, d_counters + (size_t)p * nportions + q);
            UKM_TRY(r);
            unsigned long long* t = b0;
            b0 = b1;
            b1 = t;
        }
        uint64_t* tk = ksrc;
        ksrc = kdst;
        kdst = tk;
        uint32_t* tv = vsrc;
        vsrc = vdst;
        vdst = tv;
    }
    if (ksrc != d_keys) {
        UKM_CUDA(ctx, cudaMemcpyAsync(d_keys, ksrc, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
        if (pairs) UKM_CUDA(ctx, cudaMemcpyAsync(d_vals, vsrc, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return UKM_OK;
}

static int sort_common(ukm_ctx* ctx, uint64_t* keys, uint32_t* taxids, size_t n, int key_bits, int where, const char* what) {
    if (!ctx) return UKM_E_ARG;
    if (n && !keys) return ukm_fail(ctx, UKM_E_ARG, "%s: keys == NULL", what);
    if (n < 2) return UKM_OK;
    UKM_TRY(ukm_begin_call(ctx));
    if (where == UKM_DEVICE) {
        UKM_TRY(ukm_dev_sort(ctx, keys, taxids, n, key_bits));
        return ukm_check_dev_error(ctx, what);
    }
    ukm_tmp tmp(ctx);
    uint64_t* dk = nullptr;
    uint32_t* dv = nullptr;
    UKM_TRY(tmp.alloc(&dk, n));
    UKM_CUDA(ctx, cudaMemcpyAsync(dk, keys, n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    if (taxids) {
        UKM_TRY(tmp.alloc(&dv, n));
        UKM_CUDA(ctx, cudaMemcpyAsync(dv, taxids, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    UKM_TRY(ukm_dev_sort(ctx, dk, dv, n, key_bits));
    UKM_CUDA(ctx, cudaMemcpyAsync(keys, dk, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (taxids) UKM_CUDA(ctx, cudaMemcpyAsync(taxids, dv, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    return ukm_check_dev_error(ctx, what);
}

extern "C" int ukm_sort_u64(ukm_ctx* ctx, uint64_t* keys, size_t n, int key_bits, int where) {
    return sort_common(ctx, keys, nullptr, n, key_bits, where, "ukm_sort_u64");
}

extern "C" int ukm_sort_pairs(ukm_ctx* ctx, uint64_t* keys, uint32_t* taxids, size_t n, int key_bits, int where) {
    if (ctx && n && !taxids) return ukm_fail(ctx, UKM_E_ARG, "ukm_sort_pairs: taxids == NULL");
    return sort_common(ctx, keys, taxids, n, key_bits, where, "ukm_sort_pairs");
}

extern "C" int ukm_sort_codetaxid16(ukm_ctx* ctx, void* aos16, size_t n, int key_bits) {
    if (!ctx) return UKM_E_ARG;
    if (n && !aos16) return ukm_fail(ctx, UKM_E_ARG, "ukm_sort_codetaxid16: NULL");
    if (n < 2) return UKM_OK;
    UKM_TRY(ukm_begin_call(ctx));
    ukm_tmp tmp(ctx);
    ulonglong2* d_aos = nullptr;
    uint64_t* dk = nullptr;
    uint32_t* dv = nullptr;
    UKM_TRY(tmp.alloc(&d_aos, n));
    UKM_TRY(tmp.alloc(&dk, n));
    UKM_TRY(tmp.alloc(&dv, n));
    UKM_CUDA(ctx, cudaMemcpyAsync(d_aos, aos16, n * 16, cudaMemcpyHostToDevice, ctx->stream));
    int g = ukm_grid_for(n, 256 * 4, ctx->sm_count);
    aos16_split_kernel<<<g, 256, 0, ctx->stream>>>(d_aos, dk, dv, n);
    UKM_LAUNCHED(ctx);
    UKM_TRY(ukm_dev_sort(ctx, dk, dv, n, key_bits));
    aos16_join_kernel<<<g, 256, 0, ctx->stream>>>(d_aos, dk, dv, n);
    UKM_LAUNCHED(ctx);
    UKM_CUDA(ctx, cudaMemcpyAsync(aos16, d_aos, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    return ukm_check_dev_error(ctx, "ukm_sort_codetaxid16");
}
