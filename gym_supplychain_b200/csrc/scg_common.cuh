// scg_common.cuh -- shared device helpers: Philox4x32-10 counter RNG, warp/atomic utilities.
//
// Stochastic inputs are generated on the fly from counter-based Philox streams keyed by
// (seed, global env index, episode, period) instead of the reference's per-episode tables
// drawn from one MT19937 stream (supplychain_env.py:641-672, demands_generator.py:33-89):
// same distributions, different stream (parity vs the reference is by injection, scg_inject).
#pragma once
#ifdef __CUDACC_RTC__
// run-time compilation (NVRTC, scg_jit.cu): no host headers; the CUDA built-ins are predeclared
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef unsigned long long size_t_rtc_unused;
#else
#include <cuda_runtime.h>
#include <stdint.h>
#endif

namespace scg {

// ---------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  counter = 128 bit, key = 64 bit.
// ---------------------------------------------------------------------------------
struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
#else
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// Stream tags: which input a block of random words feeds.
enum : uint32_t { STREAM_DEMAND = 0, STREAM_LEADTIME = 1, STREAM_ACTION = 2, STREAM_DEMAND_NOISE = 3,
                  STREAM_BG_DEMAND = 4, STREAM_BG_DELAY = 5 };

struct RngKey { uint32_t k0, k1; };

// One 128-bit block for (env, episode, row=period, stream, block index).
__host__ __device__ __forceinline__ Philox4 rng_block(RngKey key, uint64_t env, uint32_t episode, uint32_t row,
                                                      uint32_t stream, uint32_t block) {
    uint32_t c0 = (uint32_t)env;
    uint32_t c1 = ((uint32_t)(env >> 32) & 0xFFFFu) | ((block & 0xFFFu) << 16) | (stream << 28);
    return philox4x32_10(c0, c1, episode, row, key.k0, key.k1);
}

__host__ __device__ __forceinline__ uint32_t pick_word(const Philox4& b, int i) {
    return i == 0 ? b.x : (i == 1 ? b.y : (i == 2 ? b.z : b.w));
}

// uniform integer in [lo, lo+span) from one 32-bit word (multiply-shift; bias < span / 2^32)
__host__ __device__ __forceinline__ int32_t uniform_int(uint32_t word, int32_t lo, uint32_t span) {
#ifdef __CUDA_ARCH__
    return lo + (int32_t)__umulhi(word, span);
#else
    return lo + (int32_t)(((uint64_t)word * span) >> 32);
#endif
}

// (0,1] and [0,1) doubles from 32-bit words
__host__ __device__ __forceinline__ double u01_open_low(uint32_t w) { return ((double)w + 1.0) * (1.0 / 4294967296.0); }
__host__ __device__ __forceinline__ double u01(uint32_t w) { return (double)w * (1.0 / 4294967296.0); }

// standard normal by Box-Muller from two words (double precision: rounding of demand must not
// depend on the arithmetic mode of the handle)
__device__ __forceinline__ double normal_from(uint32_t w0, uint32_t w1) {
    double r = sqrt(-2.0 * log(u01_open_low(w0)));
    double s, c;
    sincospi(2.0 * u01(w1), &s, &c);
    return r * c;
}

// 1 + Poisson(lambda) clipped to [1, maxv] by inverse CDF (supplychain_env.py:670-672)
__device__ __forceinline__ int32_t leadtime_from(uint32_t w, double lambda, int32_t maxv) {
    double u = u01(w), p = exp(-lambda), F = p;
    int32_t k = 0;
    while (u >= F && k + 1 < maxv) { ++k; p *= lambda / (double)k; F += p; }
    return 1 + k;
}

// ---------------------------------------------------------------------------------
// warp / atomic helpers for the episode-return statistics
// ---------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ void atomic_min_double(double* addr, double v) {
    unsigned long long* a = (unsigned long long*)addr;
    unsigned long long old = *a;
    while (__longlong_as_double((long long)old) > v) {
        unsigned long long assumed = old;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}
__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
    unsigned long long* a = (unsigned long long*)addr;
    unsigned long long old = *a;
    while (__longlong_as_double((long long)old) < v) {
        unsigned long long assumed = old;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}

// Per-thread accumulator of finished-episode returns, flushed once per kernel.
struct ReturnStats {
    double cnt, sum, sq, mn, mx;
    __device__ __forceinline__ void init() { cnt = 0.0; sum = 0.0; sq = 0.0; mn = 1e300; mx = -1e300; }
    __device__ __forceinline__ void add(double r) { cnt += 1.0; sum += r; sq += r * r; mn = fmin(mn, r); mx = fmax(mx, r); }
    // all 32 lanes of the warp must call this
    __device__ __forceinline__ void flush(double* g) {
        double c = warp_sum(cnt);
        if (c == 0.0) return;              // warp-uniform
        double s = warp_sum(sum), q = warp_sum(sq), lo = warp_min(mn), hi = warp_max(mx);
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(g + 0, c); atomicAdd(g + 1, s); atomicAdd(g + 2, q);
            atomic_min_double(g + 3, lo); atomic_max_double(g + 4, hi);
        }
    }
};

}  // namespace scg
