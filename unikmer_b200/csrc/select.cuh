// select.cuh -- single-pass stream compaction: out = [gen.key(i) for i in [0,n) if gen.keep(i)],
// order preserved.  Block scan + decoupled look-back, output staged through shared memory.
#pragma once
#include "common.cuh"

constexpr int SEL_THREADS = 256;
constexpr int SEL_ITEMS = 8;
constexpr int SEL_TILE = SEL_THREADS * SEL_ITEMS;

template <typename Gen>
__global__ void __launch_bounds__(SEL_THREADS)
    select_kernel(Gen gen, size_t n, uint64_t* __restrict__ out, uint64_t* __restrict__ status, uint32_t* __restrict__ tile_counter,
                  unsigned long long* __restrict__ total_out, int num_tiles, int* __restrict__ err) {
    constexpr int NW = SEL_THREADS / 32;
    __shared__ uint64_t s_o[SEL_TILE];
    __shared__ unsigned s_scan[NW + 2];
    __shared__ int s_tile;
    __shared__ unsigned long long s_prefix;
    const int tid = threadIdx.x;
    if (tid == 0) s_tile = (int)atomicAdd(tile_counter, 1u);
    __syncthreads();
    const int tile = s_tile;
    const size_t base = (size_t)tile * SEL_TILE + (size_t)tid * SEL_ITEMS;
    uint64_t k[SEL_ITEMS];
    unsigned mask = 0;
#pragma unroll
    for (int j = 0; j < SEL_ITEMS; ++j) {
        size_t i = base + j;
        k[j] = 0;
        if (i < n) {
            bool keep;
            k[j] = gen(i, &keep);
            if (keep) mask |= 1u << j;
        }
    }
    unsigned tile_total;
    const unsigned off = block_excl_scan_u32<SEL_THREADS>((unsigned)__popc(mask), s_scan, &tile_total);
    unsigned o = off;
#pragma unroll
    for (int j = 0; j < SEL_ITEMS; ++j)
        if (mask & (1u << j)) s_o[o++] = k[j];
    const unsigned long long pre = tile_exclusive_prefix(status, tile, tile_total, err, &s_prefix);
    if (tid == 0 && tile == num_tiles - 1) *total_out = pre + tile_total;
    for (unsigned i = tid; i < tile_total; i += SEL_THREADS) out[pre + i] = s_o[i];
}

// host driver; *n_out is valid on return (stream synchronised)
template <typename Gen>
int ukm_dev_select(ukm_ctx* ctx, Gen gen, size_t n, uint64_t* d_out, size_t* n_out, const char* stat_name, double algo_bytes_in) {
    *n_out = 0;
    if (n == 0) return UKM_OK;
    const int num_tiles = (int)((n + SEL_TILE - 1) / SEL_TILE);
    ukm_tmp tmp(ctx);
    uint64_t* d_status = nullptr;
    UKM_TRY(tmp.alloc(&d_status, (size_t)num_tiles + 4));
    uint32_t* d_counter = reinterpret_cast<uint32_t*>(d_status + num_tiles);
    unsigned long long* d_total = reinterpret_cast<unsigned long long*>(d_status + num_tiles + 1);
    UKM_CUDA(ctx, cudaMemsetAsync(d_status, 0, ((size_t)num_tiles + 4) * sizeof(uint64_t), ctx->stream));
    {
        ukm_stat_scope st(ctx, stat_name, algo_bytes_in);
        select_kernel<Gen><<<num_tiles, SEL_THREADS, 0, ctx->stream>>>(gen, n, d_out, d_status, d_counter, d_total, num_tiles, ctx->d_err);
        UKM_LAUNCHED(ctx);
    }
    UKM_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch, d_total, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    UKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_out = (size_t)ctx->h_scratch[0];
    if (ctx->stats_on && !ctx->pending.empty()) ctx->pending.back().bytes += 8.0 * (double)*n_out;
    return UKM_OK;
}
