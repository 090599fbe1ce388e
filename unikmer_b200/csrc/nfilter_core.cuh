// nfilter_core.cuh -- the per-lane searches of the single-pass inter / diff filter (nfilter.cu), written so
// that the same functions run inside the CUDA kernel (on 32-bit shared-memory addresses) and, compiled by g++, inside the
// host model of the CPU test-suite (tests/host/nfilter_model.cpp: the "shared memory" is a byte array there).
//
// What they stand for: the membership tests of inter.go:228-257 / diff.go:395-431 -- "is this key of the running set in
// file f" -- on a tile's segment of file f, as searches with a warp-uniform, branch-free probe sequence.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define NF_HD __device__ __forceinline__
// Shared-memory addresses are 32-bit (ld.shared with a register address, no generic-address translation).
NF_HD uint64_t nf_lds(uint32_t a) {
    uint64_t v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a));
    return v;
}
NF_HD int nf_clz(int x) { return __clz(x); }
NF_HD uint32_t nf_minu(uint32_t a, uint32_t b) { return min(a, b); }
#else
#include <string.h>
#define NF_HD inline
extern const unsigned char* nf_host_smem;  // the model's shared memory; addresses are byte offsets into it
inline uint64_t nf_lds(uint32_t a) {
    uint64_t v;
#ifdef NF_HOST_CHECK
    NF_HOST_CHECK(a);  // the model's hook: every probe address must lie inside the segment it was given
#endif
    memcpy(&v, nf_host_smem + a, 8);
    return v;
}
inline int nf_clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline uint32_t nf_minu(uint32_t a, uint32_t b) { return a < b ? a : b; }
#endif

NF_HD int nf_lg(int n) { return 32 - nf_clz((n > 1 ? n : 1) - 1); }  // smallest lg with 2^lg >= n

// Searches with a warp-uniform, fully unrolled probe sequence: steps 2^(LG-1) .. 1 from the first element, every probe
// address clamped to the last element (one VIADDMNMX), so any length 1 <= n <= 2^LG works with the same straight-line
// code and lanes with different lengths never diverge: LDS / compare / predicated move, 5 instructions per probe.
// `last` = shared address of the last element (of the first when n = 0: the caller masks the result).
template <int LG>
NF_HD uint32_t nf_last_le(uint32_t seg, uint32_t last, uint64_t x) {  // the last element <= x (seg if there is none)
    uint32_t pp = seg;
#pragma unroll
    for (int k = LG - 1; k >= 0; --k) {
        const uint32_t a = nf_minu(pp + (8u << k), last);
        if (nf_lds(a) <= x) pp = a;
    }
    return pp;
}
template <int LG>
NF_HD uint32_t nf_last_lt(uint32_t seg, uint32_t last, uint64_t x) {  // the last element < x (seg if there is none)
    uint32_t pp = seg;
#pragma unroll
    for (int k = LG - 1; k >= 0; --k) {
        const uint32_t a = nf_minu(pp + (8u << k), last);
        if (nf_lds(a) < x) pp = a;
    }
    return pp;
}
NF_HD uint32_t nf_last_le_loop(uint32_t seg, uint32_t last, uint64_t x, int lg) {
    uint32_t pp = seg;
    for (uint32_t st = 8u << lg >> 1; st >= 8u; st >>= 1) {
        const uint32_t a = nf_minu(pp + st, last);
        if (nf_lds(a) <= x) pp = a;
    }
    return pp;
}
NF_HD uint32_t nf_last_lt_loop(uint32_t seg, uint32_t last, uint64_t x, int lg) {
    uint32_t pp = seg;
    for (uint32_t st = 8u << lg >> 1; st >= 8u; st >>= 1) {
        const uint32_t a = nf_minu(pp + st, last);
        if (nf_lds(a) < x) pp = a;
    }
    return pp;
}
// does x occur in the n sorted elements at shared address seg; lg = warp-uniform, 2^lg >= n of every lane
NF_HD bool nf_find(uint32_t seg, int n, uint64_t x, int lg) {
    const uint32_t last = seg + (unsigned)(n > 0 ? n - 1 : 0) * 8u;
    uint32_t pp;
    if (lg <= 6) pp = nf_last_le<6>(seg, last, x);
    else if (lg == 7) pp = nf_last_le<7>(seg, last, x);
    else pp = nf_last_le_loop(seg, last, x, lg);
    return n > 0 && nf_lds(pp) == x;
}
// two keys, one segment: two independent load chains
NF_HD void nf_find2(uint32_t seg, int n, uint64_t x0, uint64_t x1, int lg, bool* f0, bool* f1) {
    const uint32_t last = seg + (unsigned)(n > 0 ? n - 1 : 0) * 8u;
    uint32_t p0 = seg, p1 = seg;
    if (lg <= 7) {
#pragma unroll
        for (int k = 6; k >= 0; --k) {
            const uint32_t a0 = nf_minu(p0 + (8u << k), last), a1 = nf_minu(p1 + (8u << k), last);
            const uint64_t v0 = nf_lds(a0), v1 = nf_lds(a1);
            if (v0 <= x0) p0 = a0;
            if (v1 <= x1) p1 = a1;
        }
    } else {
        for (uint32_t st = 8u << lg >> 1; st >= 8u; st >>= 1) {
            const uint32_t a0 = nf_minu(p0 + st, last), a1 = nf_minu(p1 + st, last);
            const uint64_t v0 = nf_lds(a0), v1 = nf_lds(a1);
            if (v0 <= x0) p0 = a0;
            if (v1 <= x1) p1 = a1;
        }
    }
    *f0 = n > 0 && nf_lds(p0) == x0;
    *f1 = n > 0 && nf_lds(p1) == x1;
}
// number of elements below x
NF_HD int nf_rank(uint32_t seg, int n, uint64_t x, int lg) {
    const uint32_t last = seg + (unsigned)(n > 0 ? n - 1 : 0) * 8u;
    uint32_t pp;
    if (lg <= 10) pp = nf_last_lt<10>(seg, last, x);
    else pp = nf_last_lt_loop(seg, last, x, lg);
    return n > 0 ? (int)((pp - seg) >> 3) + (nf_lds(pp) < x ? 1 : 0) : 0;
}
