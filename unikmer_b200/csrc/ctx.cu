// ctx.cu -- context, device arena, error reporting, span staging, per-kernel statistics.
#include <stdarg.h>

#include "common.cuh"

int ukm_fail(ukm_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) {
        ctx->err = buf;
        ctx->err_stale = true;
    }
    return code;
}

int ukm_begin_call(ukm_ctx* ctx) {
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return ukm_fail(ctx, UKM_E_CUDA, "cudaSetDevice(%d): %s", ctx->device, cudaGetErrorString(e));
    if (ctx->err_stale) {
        e = cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream);
        if (e != cudaSuccess) return ukm_fail(ctx, UKM_E_CUDA, "clearing the device error word: %s", cudaGetErrorString(e));
        ctx->err_stale = false;
    }
    return UKM_OK;
}

static thread_local std::string g_create_err;

extern "C" const char* ukm_version(void) { return "unikmer-b200 0.1 (sm_100a)"; }

extern "C" const char* ukm_last_error(ukm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

extern "C" ukm_ctx* ukm_create(int device) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_err = std::string("ukm_create: no CUDA device: ") + cudaGetErrorString(e);
        return nullptr;
    }
    if (device < 0 || device >= ndev) {
        g_create_err = "ukm_create: bad device index";
        return nullptr;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) {
        g_create_err = std::string("ukm_create: cudaSetDevice: ") + cudaGetErrorString(e);
        return nullptr;
    }
    ukm_ctx* ctx = new ukm_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_create_err = std::string("ukm_create: cudaStreamCreate: ") + cudaGetErrorString(e);
        delete ctx;
        return nullptr;
    }
    // keep freed blocks cached in the stream-ordered pool (no trim at sync)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thresh = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh);
    }
    cudaMalloc(&ctx->d_err, sizeof(int));
    cudaMemset(ctx->d_err, 0, sizeof(int));
    cudaMallocHost(&ctx->h_err, sizeof(int));
    cudaMallocHost(&ctx->h_scratch, 64 * sizeof(uint64_t));
    if (!ctx->d_err || !ctx->h_err || !ctx->h_scratch) {
        g_create_err = "ukm_create: allocation of control words failed";
        ukm_destroy(ctx);
        return nullptr;
    }
    return ctx;
}

extern "C" void ukm_destroy(ukm_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto& p : ctx->pending) {
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    for (int q = 0; q < 2; ++q) {
        if (ctx->ev_in[q]) cudaEventDestroy(ctx->ev_in[q]);
        if (ctx->ev_out[q]) cudaEventDestroy(ctx->ev_out[q]);
    }
    if (ctx->ev_misc) cudaEventDestroy(ctx->ev_misc);
    cudaFree(ctx->tax.parent);
    cudaFree(ctx->tax.merged);
    cudaFree(ctx->tax.depth);
    cudaFree(ctx->d_err);
    cudaFreeHost(ctx->h_err);
    cudaFreeHost(ctx->h_scratch);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int ukm_set_stream(ukm_ctx* ctx, void* s) {
    if (!ctx) return UKM_E_ARG;
    UKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)s;
    ctx->own_stream = false;
    return UKM_OK;
}
extern "C" void* ukm_get_stream(ukm_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

extern "C" int ukm_sync(ukm_ctx* ctx) {
    if (!ctx) return UKM_E_ARG;
    UKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return UKM_OK;
}

extern "C" void* ukm_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void ukm_free_pinned(void* p) {
    if (p) cudaFreeHost(p);
}

int ukm_dev_alloc(ukm_ctx* ctx, void** p, size_t bytes) {
    *p = nullptr;
    cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 16, ctx->stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return ukm_fail(ctx, UKM_E_NOMEM, "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    return UKM_OK;
}
void ukm_dev_free(ukm_ctx* ctx, void* p) {
    if (p) cudaFreeAsync(p, ctx->stream);
}

extern "C" void* ukm_alloc_device(ukm_ctx* ctx, size_t bytes) {
    if (!ctx) return nullptr;
    void* p = nullptr;
    if (ukm_dev_alloc(ctx, &p, bytes) != UKM_OK) return nullptr;
    return p;
}
extern "C" int ukm_free_device(ukm_ctx* ctx, void* p) {
    if (!ctx) return UKM_E_ARG;
    ukm_dev_free(ctx, p);
    return UKM_OK;
}

extern "C" int ukm_copy(ukm_ctx* ctx, void* dst, int dst_where, const void* src, int src_where, size_t bytes) {
    if (!ctx || (!dst && bytes) || (!src && bytes)) return ukm_fail(ctx, UKM_E_ARG, "ukm_copy: null pointer");
    (void)dst_where;
    (void)src_where;
    if (bytes == 0) return UKM_OK;
    UKM_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->stream));
    UKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return UKM_OK;
}

// ---- stats -----------------------------------------------------------------------
static cudaEvent_t take_event(ukm_ctx* ctx) {
    if (!ctx->event_pool.empty()) {
        cudaEvent_t e = ctx->event_pool.back();
        ctx->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

ukm_stat_scope::ukm_stat_scope(ukm_ctx* c, const char* n, double algo_bytes) : ctx(c), name(n), bytes(algo_bytes) {
    if (!ctx->stats_on) return;
    a = take_event(ctx);
    b = take_event(ctx);
    cudaEventRecord(a, ctx->stream);
}
ukm_stat_scope::~ukm_stat_scope() {
    if (!a) return;
    cudaEventRecord(b, ctx->stream);
    ctx->pending.push_back({a, b, name, bytes});
}

static void drain_stats(ukm_ctx* ctx) {
    if (ctx->pending.empty()) return;
    cudaStreamSynchronize(ctx->stream);
    for (auto& p : ctx->pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            auto& s = ctx->stats[p.name];
            s.launches++;
            s.ms += ms;
            s.bytes += p.bytes;
        }
        ctx->event_pool.push_back(p.a);
        ctx->event_pool.push_back(p.b);
    }
    ctx->pending.clear();
}

extern "C" uint64_t ukm_launch_count(ukm_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int ukm_stats_enable(ukm_ctx* ctx, int on) {
    if (!ctx) return UKM_E_ARG;
    drain_stats(ctx);
    ctx->stats_on = on != 0;
    return UKM_OK;
}
extern "C" int ukm_stats_reset(ukm_ctx* ctx) {
    if (!ctx) return UKM_E_ARG;
    drain_stats(ctx);
    ctx->stats.clear();
    return UKM_OK;
}
extern "C" int ukm_stats_get(ukm_ctx* ctx, ukm_kernel_stat* out, int cap, int* n) {
    if (!ctx || !n) return UKM_E_ARG;
    drain_stats(ctx);
    int i = 0;
    for (auto& kv : ctx->stats) {
        if (i < cap && out) {
            memset(&out[i], 0, sizeof out[i]);
            strncpy(out[i].name, kv.first.c_str(), sizeof(out[i].name) - 1);
            out[i].launches = kv.second.launches;
            out[i].ms = kv.second.ms;
            out[i].algo_bytes = kv.second.bytes;
        }
        ++i;
    }
    *n = i;
    return UKM_OK;
}

// ---- device error word ---------------------------------------------------------------
int ukm_check_dev_error(ukm_ctx* ctx, const char* what) {
    UKM_CUDA(ctx, cudaMemcpyAsync(ctx->h_err, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    UKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int e = *ctx->h_err;
    if (e == 0) return UKM_OK;
    UKM_CUDA(ctx, cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream));
    switch (e) {
        case UKM_E_NOT_SORTED_UNIQUE:
            return ukm_fail(ctx, e, "%s: an input is not sorted ascending and duplicate-free", what);
        case UKM_E_ILLEGAL_BASE:
            return ukm_fail(ctx, e, "%s: illegal base in sequence (kmers.ErrIllegalBase)", what);
        case UKM_E_INTERNAL:
            return ukm_fail(ctx, e, "%s: kernel watchdog expired (internal error)", what);
        default:
            return ukm_fail(ctx, e, "%s: device reported status %d", what, e);
    }
}

// ---- span staging -----------------------------------------------------------------------
__global__ void fill_u32_kernel(uint32_t* __restrict__ d, uint32_t v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) d[i] = v;
}

int ukm_dev_fill_u32(ukm_ctx* ctx, uint32_t* d, uint32_t v, size_t n) {
    if (n == 0) return UKM_OK;
    if (v == 0) {
        UKM_CUDA(ctx, cudaMemsetAsync(d, 0, n * sizeof(uint32_t), ctx->stream));
        return UKM_OK;
    }
    fill_u32_kernel<<<ukm_grid_for(n, 256 * 8, ctx->sm_count), 256, 0, ctx->stream>>>(d, v, n);
    UKM_LAUNCHED(ctx);
    return UKM_OK;
}

int ukm_stage_in(ukm_ctx* ctx, ukm_tmp& tmp, const ukm_span* in, bool want_taxids, ukm_dspan* out) {
    out->n = in->n;
    out->keys = nullptr;
    out->taxids = nullptr;
    if (in->n && !in->keys) return ukm_fail(ctx, UKM_E_ARG, "span has n=%zu but keys == NULL", in->n);
    if (in->where == UKM_DEVICE) {
        out->keys = in->keys;
    } else {
        UKM_TRY(tmp.alloc(&out->keys, in->n + 2));
        if (in->n)
            UKM_CUDA(ctx, cudaMemcpyAsync(out->keys, in->keys, in->n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (want_taxids) {
        if (in->taxids && in->where == UKM_DEVICE) {
            out->taxids = in->taxids;
        } else {
            UKM_TRY(tmp.alloc(&out->taxids, in->n + 2));
            if (in->taxids) {
                if (in->n)
                    UKM_CUDA(ctx, cudaMemcpyAsync(out->taxids, in->taxids, in->n * sizeof(uint32_t), cudaMemcpyHostToDevice,
                                                  ctx->stream));
            } else {
                UKM_TRY(ukm_dev_fill_u32(ctx, out->taxids, in->global_taxid, in->n));
            }
        }
    }
    return UKM_OK;
}

int ukm_deliver(ukm_ctx* ctx, const uint64_t* d_keys, const uint32_t* d_taxids, size_t n, ukm_span* out) {
    if (!out) return ukm_fail(ctx, UKM_E_ARG, "out span is NULL");
    if (n > out->cap) {
        size_t cap = out->cap;
        out->n = n;
        return ukm_fail(ctx, UKM_E_CAPACITY, "output needs %zu elements, capacity is %zu", n, cap);
    }
    if (n && !out->keys) return ukm_fail(ctx, UKM_E_ARG, "out span has keys == NULL");
    cudaMemcpyKind kind = out->where == UKM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (n && d_keys != out->keys)
        UKM_CUDA(ctx, cudaMemcpyAsync(out->keys, d_keys, n * sizeof(uint64_t), kind, ctx->stream));
    if (n && out->taxids && d_taxids && d_taxids != out->taxids)
        UKM_CUDA(ctx, cudaMemcpyAsync(out->taxids, d_taxids, n * sizeof(uint32_t), kind, ctx->stream));
    if (n && out->taxids && !d_taxids) {
        if (out->where == UKM_DEVICE) UKM_CUDA(ctx, cudaMemsetAsync(out->taxids, 0, n * sizeof(uint32_t), ctx->stream));
        else memset(out->taxids, 0, n * sizeof(uint32_t));
    }
    UKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    out->n = n;
    return UKM_OK;
}
