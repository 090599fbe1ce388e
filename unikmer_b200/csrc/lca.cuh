// lca.cuh -- device restatement of bio/taxdump Taxonomy.LCA, the
// function behind the 14 taxondb.LCA call sites (union.go:199; inter.go:235,238;
// diff.go:362,407; common.go:265; count.go:386,408; sort.go:491,515;
// util-sort.go:128,151,325,374).
//
// Table layout in HBM (built by ukm_set_taxonomy): parent[t], merged[t], depth[t] for
// t in [0, n).  NCBI has ~2.6 M nodes => ~31 MB, resident in the 126 MB L2.
#pragma once
#include <stdint.h>

struct TaxDev {
    const uint32_t* parent;
    const uint32_t* merged;
    const uint32_t* depth;
    uint32_t n;
};

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t tax_resolve_dev(const TaxDev& t, uint32_t x) {
    if (x < t.n) {
        if (__ldg(t.parent + x)) return x;
        uint32_t y = __ldg(t.merged + x);
        if (y && y < t.n && __ldg(t.parent + y)) return y;
    }
    return 0;
}

// LCA(a,b): 0 if either is 0; a if a == b (no validity check); unknown id => 0.
static __device__ __noinline__ uint32_t lca_dev(const TaxDev t, uint32_t a, uint32_t b) {
    if (a == 0 || b == 0) return 0;
    if (a == b) return a;
    if (!t.parent) return 0;
    a = tax_resolve_dev(t, a);
    b = tax_resolve_dev(t, b);
    if (!a || !b) return 0;
    if (a == b) return a;
    uint32_t da = __ldg(t.depth + a), db = __ldg(t.depth + b);
    while (da > db) { a = __ldg(t.parent + a); --da; }
    while (db > da) { b = __ldg(t.parent + b); --db; }
    // bounded by the depth: both are at the same level of a rooted tree
    while (a != b && da > 0) { a = __ldg(t.parent + a); b = __ldg(t.parent + b); --da; }
    return a == b ? a : 0;
}
#endif

struct ukm_ctx;
TaxDev ukm_taxdev(const ukm_ctx* ctx);  // taxonomy.cu
