// common.cuh -- internal definitions shared by the libukm translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/ukm.h"

// ---------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------
struct ukm_taxonomy_dev {
    uint32_t* parent = nullptr;  // parent[t]; 0 = unknown
    uint32_t* merged = nullptr;  // merged[t] = new id or 0
    uint32_t* depth = nullptr;   // depth[t], root = 0
    uint32_t n = 0;              // table length
};

struct ukm_stat_acc {
    uint64_t launches = 0;
    double ms = 0;
    double bytes = 0;
};

struct ukm_pending_event {
    cudaEvent_t a, b;
    std::string name;
    double bytes;
};

struct ukm_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    std::string err;
    ukm_taxonomy_dev tax;
    // watchdog / error word written by kernels (bounded spins, sortedness checks)
    int* d_err = nullptr;
    int* h_err = nullptr;  // pinned
    // pinned scratch for small D2H results (counts)
    uint64_t* h_scratch = nullptr;  // pinned, 64 words
    // stats
    uint64_t launches = 0;  // kernels launched by this context
    bool stats_on = false;
    std::map<std::string, ukm_stat_acc> stats;
    std::vector<ukm_pending_event> pending;
    std::vector<cudaEvent_t> event_pool;
    // per-context (= per-device) kernel set-up: dynamic shared-memory opt-in done, resident CTAs per SM (-1 = not asked)
    std::map<const void*, int> kcfg;
    // a call that failed after a kernel wrote the device error word may have left it set: the next call clears it first
    bool err_stale = false;
    // streamed operations on host inputs (ukm_setops_stream): copy streams beside the compute stream, created on first use
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr}, ev_misc = nullptr;
};

int ukm_fail(ukm_ctx* ctx, int code, const char* fmt, ...);
// first thing every compute entry point does: make the context's device current and drop a device error word that an
// earlier, failed call left behind (so a stale NOT_SORTED_UNIQUE / ILLEGAL_BASE is never reported by an unrelated call)
int ukm_begin_call(ukm_ctx* ctx);

#define UKM_CUDA(ctx, call)                                                                        \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return ukm_fail((ctx), _e == cudaErrorMemoryAllocation ? UKM_E_NOMEM : UKM_E_CUDA,     \
                            "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));   \
    } while (0)

// after every kernel launch: count it (ukm_launch_count) and pick up launch errors
#define UKM_LAUNCHED(ctx)                       \
    do {                                        \
        (ctx)->launches++;                      \
        UKM_CUDA((ctx), cudaGetLastError());    \
    } while (0)

#define UKM_TRY(expr)              \
    do {                           \
        int _r = (expr);           \
        if (_r != UKM_OK) return _r; \
    } while (0)

// device arena: stream-ordered allocations on the ctx stream
int ukm_dev_alloc(ukm_ctx* ctx, void** p, size_t bytes);
void ukm_dev_free(ukm_ctx* ctx, void* p);

// RAII holder so early returns release temporaries
struct ukm_tmp {
    ukm_ctx* ctx;
    std::vector<void*> ptrs;
    explicit ukm_tmp(ukm_ctx* c) : ctx(c) {}
    ~ukm_tmp() {
        for (void* p : ptrs) ukm_dev_free(ctx, p);
    }
    template <typename T>
    int alloc(T** p, size_t count) {
        void* q = nullptr;
        int r = ukm_dev_alloc(ctx, &q, count * sizeof(T));
        if (r != UKM_OK) return r;
        ptrs.push_back(q);
        *p = static_cast<T*>(q);
        return UKM_OK;
    }
    // give up ownership of p (it becomes the caller's)
    bool release(void* p) {
        for (auto& q : ptrs)
            if (q == p) { q = ptrs.back(); ptrs.pop_back(); return true; }
        return false;
    }
    // frees p only if this holder owns it (never a caller's buffer)
    void free_now(void* p) {
        if (release(p)) ukm_dev_free(ctx, p);
    }
};

// stats: bracket a launch (or group of launches) of one kernel family
struct ukm_stat_scope {
    ukm_ctx* ctx;
    cudaEvent_t a = nullptr, b = nullptr;
    const char* name;
    double bytes;
    ukm_stat_scope(ukm_ctx* c, const char* n, double algo_bytes);
    ~ukm_stat_scope();
};

// check the device error word (one small D2H); returns UKM_OK or the recorded status
int ukm_check_dev_error(ukm_ctx* ctx, const char* what);

// staging of spans: bring an input span to the device (no copy if already there)
struct ukm_dspan {
    uint64_t* keys = nullptr;
    uint32_t* taxids = nullptr;
    size_t n = 0;
};
int ukm_stage_in(ukm_ctx* ctx, ukm_tmp& tmp, const ukm_span* in, bool want_taxids, ukm_dspan* out);
// deliver a device result into the caller's out span (capacity check, D2H if host)
int ukm_deliver(ukm_ctx* ctx, const uint64_t* d_keys, const uint32_t* d_taxids, size_t n, ukm_span* out);

// internal device-level entry points used across translation units
int ukm_dev_sort(ukm_ctx* ctx, uint64_t* d_keys, uint32_t* d_vals, size_t n, int key_bits);
int ukm_dev_fold(ukm_ctx* ctx, int mode, const uint64_t* d_keys, const uint32_t* d_taxids, size_t n,
                 bool has_taxid, uint64_t* d_out_keys, uint32_t* d_out_taxids, size_t* n_out);
int ukm_dev_fill_u32(ukm_ctx* ctx, uint32_t* d, uint32_t v, size_t n);
int ukm_dev_check_sorted_unique(ukm_ctx* ctx, const uint64_t* d_keys, size_t n);  // sets the device error word

// single-pass N-way union of 2..8 sorted duplicate-free device arrays (nway.cu)
bool ukm_nway_enabled();
int ukm_nway_union(ukm_ctx* ctx, const uint64_t* const* keys, const size_t* n, int nf, uint64_t* outK, size_t* n_out,
                   bool* fell_back);
int ukm_nway_filter(ukm_ctx* ctx, bool inter, const uint64_t* const* keys, const size_t* n, int nf, uint64_t* outK, size_t* n_out,
                    bool* fell_back);

// the same union with row-based merge levels (nunion.cu: conflict-free gathers + bitonic merge networks in registers)
bool ukm_nunion_enabled();
int ukm_nunion(ukm_ctx* ctx, const uint64_t* const* keys, const size_t* n, int nf, uint64_t* outK, size_t* n_out, bool* fell_back);

// single-pass N-way inter / diff over file-0 chunks (nfilter.cu): keys[0] filtered by membership in keys[1..nf-1]
bool ukm_nfilter_enabled();
int ukm_nfilter(ukm_ctx* ctx, bool inter, const uint64_t* const* keys, const size_t* n, int nf, uint64_t* outK, size_t* n_out,
                bool* declined);

// Opt `kern` in to `smem` bytes of dynamic shared memory on THIS context's device (the attribute is per device: a
// process may own several GPUs) and, if threads > 0, report how many CTAs of it are resident per SM.  Cached per context.
template <typename K>
int ukm_kernel_config(ukm_ctx* ctx, K kern, size_t smem, int threads, int* ctas_per_sm) {
    const void* key = reinterpret_cast<const void*>(kern);
    auto it = ctx->kcfg.find(key);
    if (it == ctx->kcfg.end()) {
        if (smem > 48 * 1024) UKM_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int nb = -1;
        if (threads > 0) {
            UKM_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem));
            if (nb < 1) return ukm_fail(ctx, UKM_E_INTERNAL, "a persistent kernel does not fit on an SM of device %d", ctx->device);
        }
        it = ctx->kcfg.emplace(key, nb).first;
    }
    if (ctas_per_sm) *ctas_per_sm = it->second;
    return UKM_OK;
}
// Persistent kernels whose CTAs wait on each other (grid-wide output-offset hand-off) are launched cooperatively: the
// driver then guarantees that the whole grid is co-resident, whatever else shares the GPU (other contexts, MPS).
template <typename K, typename A>
int ukm_launch_coop(ukm_ctx* ctx, K kern, int grid, int threads, size_t smem, A& args) {
    void* argv[] = {(void*)&args};
    UKM_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(threads), argv, smem, ctx->stream));
    ctx->launches++;
    return UKM_OK;
}

int ukm_nway_union3(ukm_ctx* ctx, const uint64_t* const* keys, const size_t* n, int nf, uint64_t* outK, size_t* n_out, uint64_t* outI,
                    size_t* n_i, uint64_t* outD, size_t* n_d, bool* fell_back);
int ukm_nfilter_both(ukm_ctx* ctx, const uint64_t* const* keys, const size_t* n, int nf, uint64_t* outI, size_t* n_i, uint64_t* outD,
                     size_t* n_d, bool* declined);

static inline int ukm_grid_for(size_t work, int per_block, int sm_count, int max_per_sm = 32) {
    size_t g = (work + per_block - 1) / per_block;
    size_t cap = (size_t)sm_count * max_per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// ---------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------
#ifdef __CUDACC__

#define UKM_WATCHDOG_SPINS (1u << 24)

__device__ __forceinline__ uint64_t sm64_dev(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ unsigned lane_id() {
    unsigned l;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// streaming 64-bit / 128-bit global accesses (read-once data: do not pollute L1)
__device__ __forceinline__ uint64_t ld_stream_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ ulonglong2 ld_stream_u64x2(const ulonglong2* p) {
    ulonglong2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream_u64x2(ulonglong2* p, ulonglong2 v) {
    asm volatile("st.global.L1::no_allocate.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v.x), "l"(v.y) : "memory");
}

// Look-back status words carry their whole payload in one 64-bit (or 32-bit) word, so relaxed
// (L1-bypassing) accesses are enough; ld.acquire.gpu would cost an L1 invalidate (CCTL.IVALL) per poll.
__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// acquire / release variants (kept for data-carrying hand-offs)
__device__ __forceinline__ uint64_t ld_acquire_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- mbarrier + 1-D TMA bulk copy (cp.async.bulk; SASS: UBLKCP / SYNCS) ------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// named barrier over `nthreads` threads (warp-specialised kernels: consumers only)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: returns false if the watchdog expired
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, unsigned parity) {
    for (unsigned spin = 0; spin < UKM_WATCHDOG_SPINS; ++spin) {
        if (mbar_try_wait(bar, parity)) return true;
        if (spin >= 2) __nanosleep(128);  // a long wait: stop competing for issue slots with the warps that have work
    }
    return false;
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// shared -> global bulk store (bulk_group completion); same alignment rules
__device__ __forceinline__ void tma_store_1d(void* gmem_dst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- warp / block scans ------------------------------------------------------------
__device__ __forceinline__ unsigned warp_incl_scan_u32(unsigned v) {
    unsigned l = lane_id();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, v, d);
        if (l >= (unsigned)d) v += t;
    }
    return v;
}

// exclusive block scan of one unsigned per thread; returns exclusive prefix, *total = block sum.
// `ws` is shared scratch of (NWARPS + 1) unsigned.  Contains two __syncthreads.
template <int NTHREADS>
__device__ __forceinline__ unsigned block_excl_scan_u32(unsigned v, unsigned* ws, unsigned* total) {
    constexpr int NW = NTHREADS / 32;
    unsigned incl = warp_incl_scan_u32(v);
    unsigned w = threadIdx.x >> 5, l = lane_id();
    if (l == 31) ws[w] = incl;
    __syncthreads();
    if (w == 0) {
        unsigned x = (l < NW) ? ws[l] : 0u;
        unsigned xi = warp_incl_scan_u32(x);
        if (l < NW) ws[l] = xi - x;
        if (l == NW - 1) ws[NW] = xi;
    }
    __syncthreads();
    *total = ws[NW];
    return ws[w] + incl - v;
}

// ---- decoupled look-back over one 64-bit status word per tile ----------------------
// word = flag << 62 | value ; flag 0 = empty, 1 = partial (tile aggregate), 2 = inclusive
#define UKM_LB_PARTIAL (1ull << 62)
#define UKM_LB_INCLUSIVE (2ull << 62)
#define UKM_LB_VALUE(x) ((x) & ((1ull << 62) - 1))

// Called by ALL lanes of one warp.  Publishes `aggregate` for `tile`, returns the exclusive
// prefix (sum of aggregates of tiles < tile) to every lane.  Bounded: on watchdog expiry sets
// *err and returns 0.
__device__ __forceinline__ uint64_t lookback_warp(uint64_t* status, int tile, uint64_t aggregate, int* err) {
    unsigned l = lane_id();
    if (tile == 0) {
        if (l == 0) st_relaxed_u64(&status[0], UKM_LB_INCLUSIVE | aggregate);
        return 0;
    }
    if (l == 0) st_relaxed_u64(&status[tile], UKM_LB_PARTIAL | aggregate);
    uint64_t prefix = 0;
    int base = tile - 1;  // lane l inspects tile base - l
    unsigned spins = 0;
    while (true) {
        int j = base - (int)l;
        uint64_t w = (j >= 0) ? ld_relaxed_u64(&status[j]) : UKM_LB_INCLUSIVE;  // before tile 0: inclusive 0
        unsigned flag = (unsigned)(w >> 62);
        unsigned empty = __ballot_sync(0xffffffffu, flag == 0);
        unsigned incl = __ballot_sync(0xffffffffu, flag == 2);
        // usable window: lanes below the first empty lane
        unsigned first_empty = empty ? (unsigned)(__ffs(empty) - 1) : 32u;
        unsigned first_incl = incl ? (unsigned)(__ffs(incl) - 1) : 32u;
        if (first_incl < first_empty) {
            // sum lanes 0..first_incl
            uint64_t v = (l <= first_incl) ? UKM_LB_VALUE(w) : 0ull;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            prefix += v;
            break;
        }
        if (first_empty > 0) {
            // consume the partials of lanes 0..first_empty-1 and move the window back
            uint64_t v = (l < first_empty) ? UKM_LB_VALUE(w) : 0ull;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            prefix += v;
            base -= (int)first_empty;
            spins = 0;
        } else {
            if (++spins > UKM_WATCHDOG_SPINS) {
                if (l == 0) atomicExch(err, (int)UKM_E_INTERNAL);
                prefix = 0;
                break;
            }
        }
    }
    if (l == 0) st_relaxed_u64(&status[tile], UKM_LB_INCLUSIVE | (prefix + aggregate));
    return prefix;
}

// 256-wide variant: every lane inspects 8 consecutive predecessors (nearest first), so one round
// covers every tile that can be resident at once (a 32-wide window needs ~N_resident/64 serial L2
// round trips when a wave of tiles publishes together: measured 20% of the kernel).  `publish`
// = also write this tile's PARTIAL word first.  Called by all lanes of one warp.
__device__ __forceinline__ uint64_t lookback_wide(uint64_t* status, int tile, uint64_t aggregate, bool publish, int* err) {
    const unsigned l = lane_id();
    if (tile == 0) {
        if (l == 0) st_relaxed_u64(&status[0], UKM_LB_INCLUSIVE | aggregate);
        return 0;
    }
    if (publish && l == 0) st_relaxed_u64(&status[tile], UKM_LB_PARTIAL | aggregate);
    uint64_t prefix = 0;
    int base = tile - 1;  // lane l inspects tiles base - 8l - m, m = 0..7
    unsigned spins = 0;
    while (true) {
        uint64_t w[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int j = base - 8 * (int)l - m;
            w[m] = (j >= 0) ? ld_relaxed_u64(&status[j]) : UKM_LB_INCLUSIVE;
        }
        uint64_t sum = 0;
        unsigned state = 0, taken = 8;  // state: 0 all partial, 1 stopped at an empty word, 2 hit an inclusive word
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const unsigned flag = (unsigned)(w[m] >> 62);
            if (state == 0) {
                if (flag == 0) {
                    state = 1;
                    taken = m;
                } else {
                    sum += UKM_LB_VALUE(w[m]);
                    if (flag == 2) {
                        state = 2;
                        taken = m + 1;
                    }
                }
            }
        }
        const unsigned stopped = __ballot_sync(0xffffffffu, state != 0);
        const unsigned f = stopped ? (unsigned)(__ffs(stopped) - 1) : 32u;  // first lane that stopped
        uint64_t v = (l <= f) ? sum : 0ull;                                  // lanes before it contribute fully
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        prefix += v;
        const unsigned fstate = __shfl_sync(0xffffffffu, state, f & 31u);
        const unsigned ftaken = __shfl_sync(0xffffffffu, taken, f & 31u);
        if (f < 32u && fstate == 2) break;
        const unsigned consumed = (f < 32u) ? 8u * f + ftaken : 256u;
        base -= (int)consumed;
        if (consumed == 0) {
            if (++spins > UKM_WATCHDOG_SPINS) {
                if (l == 0) atomicExch(err, (int)UKM_E_INTERNAL);
                prefix = 0;
                break;
            }
        } else {
            spins = 0;
        }
    }
    if (l == 0) st_relaxed_u64(&status[tile], UKM_LB_INCLUSIVE | (prefix + aggregate));
    return prefix;
}

// block_excl_scan_u32 over the NTHREADS threads of a named barrier (tid = index within that group)
template <int NTHREADS>
__device__ __forceinline__ unsigned group_excl_scan_u32(unsigned v, unsigned tid, unsigned* ws, unsigned* total, int bar_id) {
    constexpr int NW = NTHREADS / 32;
    unsigned incl = warp_incl_scan_u32(v);
    unsigned w = tid >> 5, l = lane_id();
    if (l == 31) ws[w] = incl;
    named_bar_sync(bar_id, NTHREADS);
    if (w == 0) {
        unsigned x = (l < NW) ? ws[l] : 0u;
        unsigned xi = warp_incl_scan_u32(x);
        if (l < NW) ws[l] = xi - x;
        if (l == NW - 1) ws[NW] = xi;
    }
    named_bar_sync(bar_id, NTHREADS);
    *total = ws[NW];
    return ws[w] + incl - v;
}

// All threads call this after staging their outputs: warp 0 chains the tile total, the barrier
// publishes both the prefix and the staged data.  `s_prefix` is one shared 64-bit word.
__device__ __forceinline__ uint64_t tile_exclusive_prefix(uint64_t* status, int tile, uint64_t tile_total, int* err,
                                                          unsigned long long* s_prefix) {
    if (threadIdx.x < 32) {
        uint64_t pre = lookback_warp(status, tile, tile_total, err);
        if (threadIdx.x == 0) *s_prefix = pre;
    }
    __syncthreads();
    return *s_prefix;
}

#endif  // __CUDACC__
