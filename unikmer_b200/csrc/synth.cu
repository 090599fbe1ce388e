// synth.cu -- counter-based synthetic inputs of SURVEY.md 8(d), generated on the device so the
// benchmark never ships gigabytes over PCIe.  Bit-identical to oracle.c's generators
// (orc_random_keys / orc_member_file / orc_synth_bases).  Bench/test helpers, not a product path.
#include "common.cuh"
#include "select.cuh"

namespace {

__global__ void random_keys_kernel(uint64_t i0, size_t n, uint64_t seed, uint64_t* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = sm64_dev(seed + i0 + i) >> 2;
}

__global__ void synth_bases_kernel(uint64_t r, uint64_t i0, size_t n, uint64_t S, uint8_t* __restrict__ out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; t < n; t += stride) {
        uint64_t i = i0 + t;
        out[t] = (uint8_t)"ACGT"[(sm64_dev(S + (r << 32) + (i >> 5)) >> (2 * (i & 31))) & 3u];
    }
}

struct MemberGen {
    uint64_t j0, W, S, T;
    int f;
    __device__ __forceinline__ uint64_t operator()(size_t i, bool* keep) const {
        uint64_t j = j0 + i;
        *keep = f < 0 ? true : ((sm64_dev(T + j) >> f) & 1u) != 0;
        return j * W + (sm64_dev(S + j) % W);
    }
};

}  // namespace

extern "C" int ukm_synth_random_keys(ukm_ctx* ctx, uint64_t i0, size_t count, uint64_t seed, uint64_t* d_out) {
    if (!ctx) return UKM_E_ARG;
    if (count && !d_out) return ukm_fail(ctx, UKM_E_ARG, "ukm_synth_random_keys: NULL");
    if (!count) return UKM_OK;
    UKM_TRY(ukm_begin_call(ctx));
    random_keys_kernel<<<ukm_grid_for(count, 256 * 8, ctx->sm_count), 256, 0, ctx->stream>>>(i0, count, seed, d_out);
    UKM_LAUNCHED(ctx);
    UKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return UKM_OK;
}

extern "C" int ukm_synth_member_file(ukm_ctx* ctx, uint64_t j0, size_t count, uint64_t N, uint64_t S, uint64_t T, int f,
                                     uint64_t* d_out, size_t* n_out) {
    if (!ctx) return UKM_E_ARG;
    if (!n_out || (count && !d_out) || N == 0 || f > 63) return ukm_fail(ctx, UKM_E_ARG, "ukm_synth_member_file: bad argument");
    UKM_TRY(ukm_begin_call(ctx));
    MemberGen g{j0, (1ull << 62) / N, S, T, f};
    UKM_TRY(ukm_dev_select(ctx, g, count, d_out, n_out, "synth_member", 0.0));
    return ukm_check_dev_error(ctx, "ukm_synth_member_file");
}

extern "C" int ukm_synth_bases(ukm_ctx* ctx, uint64_t r, uint64_t i0, size_t count, uint64_t S, uint8_t* d_out) {
    if (!ctx) return UKM_E_ARG;
    if (count && !d_out) return ukm_fail(ctx, UKM_E_ARG, "ukm_synth_bases: NULL");
    if (!count) return UKM_OK;
    UKM_TRY(ukm_begin_call(ctx));
    synth_bases_kernel<<<ukm_grid_for(count, 256 * 8, ctx->sm_count), 256, 0, ctx->stream>>>(r, i0, count, S, d_out);
    UKM_LAUNCHED(ctx);
    UKM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return UKM_OK;
}
