// fold.cu -- one-pass folds over a sorted slice, sortedness check, key-range partition.
//
// ukm_fold_sorted replaces the scan loops after the in-memory sort (sort.go:482-573) and
// the chunk writers dumpCodes2File / dumpCodesTaxids2File (util-sort.go:35-190):
// plain copy / -u first-of-run (+LCA over the run's taxids) / -d repeated codes.
// Single pass: every CTA flags run heads in a shared-memory tile, compacts with a block
// scan and chains tile totals with a decoupled look-back.
#include "common.cuh"
#include "lca.cuh"

namespace {

constexpr int FD_THREADS = 256;
constexpr int FD_ITEMS = 8;
constexpr int FD_TILE = FD_THREADS * FD_ITEMS;
#define FD_P(q) ((q) + (((q) + 7) >> 3))  // slot of halo-relative index q (q = element index + 1); FD_P(8t+1 .. 8t+8) are consecutive

template <int MODE, bool TAX>
__global__ void __launch_bounds__(FD_THREADS)
    fold_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ taxids, size_t n, uint64_t* __restrict__ outK,
                uint32_t* __restrict__ outT, uint64_t* __restrict__ status, uint32_t* __restrict__ tile_counter,
                unsigned long long* __restrict__ total_out, int num_tiles, TaxDev tax, int* __restrict__ err) {
    constexpr int NW = FD_THREADS / 32;
    // tile element i lives at s_k[FD_P(i + 1)]: one pad slot per 8 elements, so that the blocked walk (thread t owns
    // elements 8t .. 8t+7) strides 9 slots across a warp and every 8-byte access hits its own bank pair (the unpadded
    // layout was an 8-way conflict on each of the three loads per key: half of the kernel's time)
    __shared__ uint64_t s_k[FD_TILE + FD_TILE / 8 + 4];  // FD_P(0) = left halo, FD_P(1..TILE) tile, FD_P(TILE+1) right halo
    __shared__ uint64_t s_ok[FD_TILE + 2];     // staging: a tile emits at most valid+1 (a run emitting 2 has >= 2 elements)
    __shared__ uint32_t s_ot[TAX ? FD_TILE + 2 : 1];
    __shared__ unsigned s_scan[NW + 2];
    __shared__ int s_tile;
    __shared__ unsigned long long s_prefix;

    const int tid = threadIdx.x;
    if (tid == 0) s_tile = (int)atomicAdd(tile_counter, 1u);
    __syncthreads();
    const int tile = s_tile;
    const size_t base = (size_t)tile * FD_TILE;
    const int valid = (n - base) < (size_t)FD_TILE ? (int)(n - base) : FD_TILE;

    for (int i = tid; i < valid; i += FD_THREADS) s_k[FD_P(1 + i)] = ld_stream_u64(keys + base + i);
    if (tid == 0) {
        // halos: has_left / has_right say whether a neighbour exists at all
        s_k[FD_P(0)] = base > 0 ? keys[base - 1] : 0;
        s_k[FD_P(1 + valid)] = (base + valid < n) ? keys[base + valid] : 0;
    }
    __syncthreads();
    const bool has_left = base > 0;
    const bool has_right = base + valid < n;

    // blocked walk: thread owns elements [tid*ITEMS, tid*ITEMS+ITEMS)
    unsigned emit2 = 0;  // 2 bits per item: number of copies to emit (0,1,2)
    uint32_t lca[FD_ITEMS];
    unsigned cnt = 0;
#pragma unroll
    for (int j = 0; j < FD_ITEMS; ++j) {
        const int i = tid * FD_ITEMS + j;
        unsigned e = 0;
        lca[j] = 0;
        if (i < valid) {
            const uint64_t k = s_k[FD_P(1 + i)];
            const bool head = (i > 0 || has_left) ? (s_k[FD_P(i)] != k) : true;
            const bool next_same = (i + 1 < valid || has_right) ? (s_k[FD_P(2 + i)] == k) : false;
            if (head) {
                if (MODE == UKM_FOLD_UNIQUE) e = 1;
                else if (MODE == UKM_FOLD_REPEATED_FINAL) e = next_same ? 1 : 0;
                else e = next_same ? 2 : 1;  // REPEATED_CHUNK
                if (TAX && e) {
                    // LCA over the whole run (it may leave the tile: walk global memory)
                    size_t g = base + i;
                    uint32_t l = taxids[g];
                    if (next_same) {
                        for (size_t q = g + 1; q < n && keys[q] == k; ++q) l = lca_dev(tax, taxids[q], l);
                    }
                    lca[j] = l;
                }
            }
        }
        emit2 |= e << (2 * j);
        cnt += e;
    }
    unsigned tile_total;
    const unsigned off = block_excl_scan_u32<FD_THREADS>(cnt, s_scan, &tile_total);
    {
        unsigned o = off;
#pragma unroll
        for (int j = 0; j < FD_ITEMS; ++j) {
            const unsigned e = (emit2 >> (2 * j)) & 3u;
            if (e) {
                const uint64_t k = s_k[FD_P(1 + tid * FD_ITEMS + j)];
                s_ok[o] = k;
                if (TAX) s_ot[o] = lca[j];
                ++o;
                if (e == 2) {
                    s_ok[o] = k;
                    if (TAX) s_ot[o] = lca[j];
                    ++o;
                }
            }
        }
    }
    const unsigned long long pre = tile_exclusive_prefix(status, tile, tile_total, err, &s_prefix);
    if (tid == 0 && tile == num_tiles - 1) *total_out = pre + tile_total;
    for (unsigned i = tid; i < tile_total; i += FD_THREADS) {
        outK[pre + i] = s_ok[i];
        if (TAX) outT[pre + i] = s_ot[i];
    }
}

template <int MODE>
int launch_fold(ukm_ctx* ctx, bool tax, int num_tiles, const uint64_t* k, const uint32_t* t, size_t n, uint64_t* ok, uint32_t* ot,
                uint64_t* status, uint32_t* counter, unsigned long long* total) {
    if (tax)
        fold_kernel<MODE, true><<<num_tiles, FD_THREADS, 0, ctx->stream>>>(k, t, n, ok, ot, status, counter, total, num_tiles,
                                                                           ukm_taxdev(ctx), ctx->d_err);
    else
        fold_kernel<MODE, false><<<num_tiles, FD_THREADS, 0, ctx->stream>>>(k, t, n, ok, ot, status, counter, total, num_tiles,
                                                                            ukm_taxdev(ctx), ctx->d_err);
    UKM_LAUNCHED(ctx);
    return UKM_OK;
}

__global__ void check_sorted_unique_kernel(const uint64_t* __restrict__ keys, size_t n, int* __restrict__ err) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (; i + 1 < n; i += stride) bad |= keys[i] >= keys[i + 1];
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicExch(err, (int)UKM_E_NOT_SORTED_UNIQUE);
}

__global__ void partition_sorted_kernel(const uint64_t* __restrict__ keys, size_t n, const uint64_t* __restrict__ splitters,
                                        int n_split, unsigned long long* __restrict__ offsets) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_split) return;
    uint64_t x = splitters[s];
    size_t lo = 0, hi = n;
    while (lo < hi) {
        size_t mid = lo + ((hi - lo) >> 1);
        if (keys[mid] < x) lo = mid + 1;
        else hi = mid;
    }
    offsets[s] = lo;
}

}  // namespace

// output capacity: n elements (every mode emits at most one element per input element)
int ukm_dev_fold(ukm_ctx* ctx, int mode, const uint64_t* d_keys, const uint32_t* d_taxids, size_t n, bool has_taxid,
                 uint64_t* d_out_keys, uint32_t* d_out_taxids, size_t* n_out) {
    *n_out = 0;
    if (n == 0) return UKM_OK;
    if (mode == UKM_FOLD_PLAIN) {
        if (d_out_keys != d_keys) UKM_CUDA(ctx, cudaMemcpyAsync(d_out_keys, d_keys, n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        if (has_taxid && d_out_taxids != d_taxids)
            UKM_CUDA(ctx, cudaMemcpyAsync(d_out_taxids, d_taxids, n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        *n_out = n;
        return UKM_OK;
    }
    const int num_tiles = (int)((n + FD_TILE - 1) / FD_TILE);
    ukm_tmp tmp(ctx);
    uint64_t* d_status = nullptr;
    UKM_TRY(tmp.alloc(&d_status, (size_t)num_tiles + 4));
    uint32_t* d_counter = reinterpret_cast<uint32_t*>(d_status + num_tiles);
    unsigned long long* d_total = reinterpret_cast<unsigned long long*>(d_status + num_tiles + 1);
    UKM_CUDA(ctx, cudaMemsetAsync(d_status, 0, ((size_t)num_tiles + 4) * sizeof(uint64_t), ctx->stream));
    {
        ukm_stat_scope st(ctx, has_taxid ? "fold_tax" : "fold", (double)n * (has_taxid ? 12.0 : 8.0));
        int r;
        if (mode == UKM_FOLD_UNIQUE)
            r = launch_fold<UKM_FOLD_UNIQUE>(ctx, has_taxid, num_tiles, d_keys, d_taxids, n, d_out_keys, d_out_taxids, d_status, d_counter, d_total);
        else if (mode == UKM_FOLD_REPEATED_FINAL)
            r = launch_fold<UKM_FOLD_REPEATED_FINAL>(ctx, has_taxid, num_tiles, d_keys, d_taxids, n, d_out_keys, d_out_taxids, d_status, d_counter, d_total);
        else
            r = launch_fold<UKM_FOLD_REPEATED_CHUNK>(ctx, has_taxid, num_tiles, d_keys, d_taxids, n, d_out_keys, d_out_taxids, d_status, d_counter, d_total);
        UKM_TRY(r);
    }
    UKM_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch, d_total, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    UKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_out = (size_t)ctx->h_scratch[0];
    if (ctx->stats_on && !ctx->pending.empty()) ctx->pending.back().bytes += (double)*n_out * (has_taxid ? 12.0 : 8.0);
    return UKM_OK;
}

extern "C" int ukm_fold_sorted(ukm_ctx* ctx, int mode, const ukm_span* in, unsigned flags, ukm_span* out) {
    if (!ctx) return UKM_E_ARG;
    if (!in || !out) return ukm_fail(ctx, UKM_E_ARG, "ukm_fold_sorted: NULL span");
    if (mode < UKM_FOLD_PLAIN || mode > UKM_FOLD_REPEATED_CHUNK) return ukm_fail(ctx, UKM_E_ARG, "ukm_fold_sorted: bad mode");
    UKM_TRY(ukm_begin_call(ctx));
    const bool tax = (flags & UKM_F_TAXID) != 0;
    if (tax && mode != UKM_FOLD_PLAIN && !ctx->tax.parent)
        return ukm_fail(ctx, UKM_E_NO_TAXONOMY, "ukm_fold_sorted: taxids requested but no taxonomy loaded");
    ukm_tmp tmp(ctx);
    ukm_dspan d;
    UKM_TRY(ukm_stage_in(ctx, tmp, in, tax, &d));
    // Reference quirks that hinge on the sentinel `last = ^uint64(0)`:
    const uint64_t SENT = ~0ull;
    if (in->n == 0) {
        // sort.go:505-507 / util-sort.go:142-144: -u with taxids writes (last, lca) unconditionally;
        // util-sort.go:81-87,166-176: the chunk -d writer writes `last` once unconditionally.
        if ((mode == UKM_FOLD_UNIQUE && tax) || mode == UKM_FOLD_REPEATED_CHUNK) {
            uint64_t* dk;
            uint32_t* dt;
            UKM_TRY(tmp.alloc(&dk, 2));
            UKM_TRY(tmp.alloc(&dt, 2));
            UKM_CUDA(ctx, cudaMemcpyAsync(dk, &SENT, 8, cudaMemcpyHostToDevice, ctx->stream));
            UKM_CUDA(ctx, cudaMemsetAsync(dt, 0, 8, ctx->stream));
            return ukm_deliver(ctx, dk, tax ? dt : nullptr, 1, out);
        }
        return ukm_deliver(ctx, nullptr, nullptr, 0, out);
    }
    uint64_t first_key = 0;
    UKM_CUDA(ctx, cudaMemcpyAsync(&first_key, d.keys, 8, cudaMemcpyDeviceToHost, ctx->stream));
    UKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (mode == UKM_FOLD_UNIQUE && first_key == SENT) {
        // sorted => every code is ^uint64(0).  Without taxids the first code equals `last` and is
        // skipped (sort.go:542-549): empty output.  With taxids the run folds into lca = LCA(t, 0) = 0.
        if (!tax) return ukm_deliver(ctx, nullptr, nullptr, 0, out);
        uint64_t* dk;
        uint32_t* dt;
        UKM_TRY(tmp.alloc(&dk, 2));
        UKM_TRY(tmp.alloc(&dt, 2));
        UKM_CUDA(ctx, cudaMemcpyAsync(dk, &SENT, 8, cudaMemcpyHostToDevice, ctx->stream));
        UKM_CUDA(ctx, cudaMemsetAsync(dt, 0, 8, ctx->stream));
        return ukm_deliver(ctx, dk, dt, 1, out);
    }
    uint64_t* ok = nullptr;
    uint32_t* ot = nullptr;
    UKM_TRY(tmp.alloc(&ok, in->n + 2));
    if (tax) UKM_TRY(tmp.alloc(&ot, in->n + 2));
    size_t m = 0;
    UKM_TRY(ukm_dev_fold(ctx, mode, d.keys, d.taxids, d.n, tax, ok, ot, &m));
    UKM_TRY(ukm_check_dev_error(ctx, "ukm_fold_sorted"));
    return ukm_deliver(ctx, ok, tax ? ot : nullptr, m, out);
}

int ukm_dev_check_sorted_unique(ukm_ctx* ctx, const uint64_t* d_keys, size_t n) {
    if (n < 2) return UKM_OK;
    ukm_stat_scope st(ctx, "validate_sorted_unique", 8.0 * (double)n);
    check_sorted_unique_kernel<<<ukm_grid_for(n, 256 * 8, ctx->sm_count), 256, 0, ctx->stream>>>(d_keys, n, ctx->d_err);
    UKM_LAUNCHED(ctx);
    return UKM_OK;
}

extern "C" int ukm_check_sorted_unique(ukm_ctx* ctx, const ukm_span* in) {
    if (!ctx) return UKM_E_ARG;
    if (!in) return ukm_fail(ctx, UKM_E_ARG, "ukm_check_sorted_unique: NULL span");
    if (in->n < 2) return UKM_OK;
    UKM_TRY(ukm_begin_call(ctx));
    ukm_tmp tmp(ctx);
    ukm_dspan d;
    UKM_TRY(ukm_stage_in(ctx, tmp, in, false, &d));
    UKM_TRY(ukm_dev_check_sorted_unique(ctx, d.keys, d.n));
    return ukm_check_dev_error(ctx, "ukm_check_sorted_unique");
}

extern "C" int ukm_partition_sorted(ukm_ctx* ctx, const ukm_span* in, const uint64_t* splitters, int n_split, uint64_t* offsets) {
    if (!ctx) return UKM_E_ARG;
    if (!in || !offsets || n_split < 0 || (n_split && !splitters)) return ukm_fail(ctx, UKM_E_ARG, "ukm_partition_sorted: bad argument");
    UKM_TRY(ukm_begin_call(ctx));
    offsets[0] = 0;
    offsets[n_split + 1] = in->n;
    if (n_split == 0) return UKM_OK;
    if (in->where != UKM_DEVICE) {
        // host span: binary searches on the host, no transfer
        for (int s = 0; s < n_split; ++s) {
            size_t lo = 0, hi = in->n;
            while (lo < hi) {
                size_t mid = lo + ((hi - lo) >> 1);
                if (in->keys[mid] < splitters[s]) lo = mid + 1;
                else hi = mid;
            }
            offsets[s + 1] = lo;
        }
        return UKM_OK;
    }
    ukm_tmp tmp(ctx);
    uint64_t* d_split = nullptr;
    unsigned long long* d_off = nullptr;
    UKM_TRY(tmp.alloc(&d_split, (size_t)n_split));
    UKM_TRY(tmp.alloc(&d_off, (size_t)n_split));
    UKM_CUDA(ctx, cudaMemcpyAsync(d_split, splitters, (size_t)n_split * 8, cudaMemcpyHostToDevice, ctx->stream));
    partition_sorted_kernel<<<(n_split + 63) / 64, 64, 0, ctx->stream>>>(in->keys, in->n, d_split, n_split, d_off);
    UKM_LAUNCHED(ctx);
    UKM_CUDA(ctx, cudaMemcpyAsync(offsets + 1, d_off, (size_t)n_split * 8, cudaMemcpyDeviceToHost, ctx->stream));
    UKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return UKM_OK;
}
