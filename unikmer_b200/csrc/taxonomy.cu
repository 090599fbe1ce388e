// taxonomy.cu -- device taxonomy tables for the LCA folds.
// Host side of loadTaxonomy (util.go:119-171): the caller parses nodes.dmp / merged.dmp
// (bio/taxdump) and hands over child->parent and old->new arrays; this builds the dense
// parent/merged/depth tables lca.cuh walks.
#include "common.cuh"
#include "lca.cuh"

TaxDev ukm_taxdev(const ukm_ctx* ctx) {
    TaxDev t;
    t.parent = ctx->tax.parent;
    t.merged = ctx->tax.merged;
    t.depth = ctx->tax.depth;
    t.n = ctx->tax.n;
    return t;
}

extern "C" int ukm_set_taxonomy(ukm_ctx* ctx, const uint32_t* parent, size_t n, const uint32_t* merged_from,
                                const uint32_t* merged_to, size_t n_merged) {
    if (!ctx) return UKM_E_ARG;
    if (!parent || n == 0) return ukm_fail(ctx, UKM_E_ARG, "ukm_set_taxonomy: empty parent table");
    if (n_merged && (!merged_from || !merged_to)) return ukm_fail(ctx, UKM_E_ARG, "ukm_set_taxonomy: merged arrays NULL");
    UKM_TRY(ukm_begin_call(ctx));
    size_t nn = n;
    for (size_t i = 0; i < n_merged; ++i)
        if ((size_t)merged_from[i] + 1 > nn) nn = (size_t)merged_from[i] + 1;
    if (nn > 0xffffffffull) return ukm_fail(ctx, UKM_E_ARG, "ukm_set_taxonomy: table too large");
    std::vector<uint32_t> par(nn, 0), mer(nn, 0), dep(nn, 0);
    memcpy(par.data(), parent, n * sizeof(uint32_t));
    for (size_t t = 0; t < n; ++t)
        if (par[t] >= n || (par[t] && !par[par[t]]))
            return ukm_fail(ctx, UKM_E_ARG, "ukm_set_taxonomy: parent of %zu (%u) is not a known taxid", t, par[t]);
    for (size_t i = 0; i < n_merged; ++i) mer[merged_from[i]] = merged_to[i];
    // depth by memoised climbing; 0xffffffff = not yet known
    const uint32_t UNK = 0xffffffffu;
    std::vector<uint32_t> d(nn, UNK);
    std::vector<uint32_t> path;
    for (size_t t = 0; t < n; ++t) {
        if (!par[t] || d[t] != UNK) continue;
        path.clear();
        uint32_t x = (uint32_t)t;
        while (d[x] == UNK && par[x] != x) {
            path.push_back(x);
            x = par[x];
            if (path.size() > nn) return ukm_fail(ctx, UKM_E_ARG, "ukm_set_taxonomy: cycle in the parent table at %zu", t);
        }
        uint32_t base = (par[x] == x && d[x] == UNK) ? (d[x] = 0) : d[x];
        for (size_t i = path.size(); i-- > 0;) d[path[i]] = ++base;
    }
    for (size_t t = 0; t < nn; ++t) dep[t] = d[t] == UNK ? 0 : d[t];

    UKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->tax.parent);
    cudaFree(ctx->tax.merged);
    cudaFree(ctx->tax.depth);
    ctx->tax = ukm_taxonomy_dev();
    UKM_CUDA(ctx, cudaMalloc(&ctx->tax.parent, nn * sizeof(uint32_t)));
    UKM_CUDA(ctx, cudaMalloc(&ctx->tax.merged, nn * sizeof(uint32_t)));
    UKM_CUDA(ctx, cudaMalloc(&ctx->tax.depth, nn * sizeof(uint32_t)));
    UKM_CUDA(ctx, cudaMemcpy(ctx->tax.parent, par.data(), nn * sizeof(uint32_t), cudaMemcpyHostToDevice));
    UKM_CUDA(ctx, cudaMemcpy(ctx->tax.merged, mer.data(), nn * sizeof(uint32_t), cudaMemcpyHostToDevice));
    UKM_CUDA(ctx, cudaMemcpy(ctx->tax.depth, dep.data(), nn * sizeof(uint32_t), cudaMemcpyHostToDevice));
    ctx->tax.n = (uint32_t)nn;
    return UKM_OK;
}

__global__ void lca_batch_kernel(TaxDev t, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint32_t* __restrict__ out,
                                 size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = lca_dev(t, a[i], b[i]);
}

extern "C" int ukm_lca_batch(ukm_ctx* ctx, const uint32_t* a, const uint32_t* b, size_t n, uint32_t* out, int where) {
    if (!ctx) return UKM_E_ARG;
    if (n && (!a || !b || !out)) return ukm_fail(ctx, UKM_E_ARG, "ukm_lca_batch: NULL");
    if (n == 0) return UKM_OK;
    UKM_TRY(ukm_begin_call(ctx));
    ukm_tmp tmp(ctx);
    const uint32_t *da = a, *db = b;
    uint32_t* dout = out;
    if (where != UKM_DEVICE) {
        uint32_t *ta, *tb;
        UKM_TRY(tmp.alloc(&ta, n));
        UKM_TRY(tmp.alloc(&tb, n));
        UKM_TRY(tmp.alloc(&dout, n));
        UKM_CUDA(ctx, cudaMemcpyAsync(ta, a, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        UKM_CUDA(ctx, cudaMemcpyAsync(tb, b, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        da = ta;
        db = tb;
    }
    lca_batch_kernel<<<ukm_grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(ukm_taxdev(ctx), da, db, dout, n);
    UKM_LAUNCHED(ctx);
    if (where != UKM_DEVICE) UKM_CUDA(ctx, cudaMemcpyAsync(out, dout, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    UKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return UKM_OK;
}
