// Shared helpers for the ddnerf_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ddnerf_b200.h"

#define DDNERF_EXPORT __attribute__((visibility("default")))

namespace ddnerf {

void set_error(const char* fmt, ...);
void count_launches(int n);          // bookkeeping for ddnerf_launch_count()

#define DDNERF_CHECK_ARG(cond, ...)              \
    do {                                         \
        if (!(cond)) {                           \
            ::ddnerf::set_error(__VA_ARGS__);    \
            return 1;                            \
        }                                        \
    } while (0)

#define DDNERF_CHECK_LAUNCH(name)                                              \
    do {                                                                       \
        cudaError_t e__ = cudaGetLastError();                                  \
        if (e__ != cudaSuccess) {                                              \
            ::ddnerf::set_error("%s: %s", name, cudaGetErrorString(e__));      \
            return 2;                                                          \
        }                                                                      \
    } while (0)

// after a successful launch of n kernels
#define DDNERF_LAUNCHED(name, n)        \
    do {                                \
        DDNERF_CHECK_LAUNCH(name);      \
        ::ddnerf::count_launches(n);    \
    } while (0)

constexpr unsigned FULL = 0xffffffffu;

// Reductions / scans over a power-of-two lane group of width G (G <= 32) inside a warp.
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o, G);
    return v;
}
template <int G>
__device__ __forceinline__ float group_incl_sum(float v, int gl) {
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
        float n = __shfl_up_sync(FULL, v, o, G);
        if (gl >= o) v += n;
    }
    return v;
}
// Inclusive sum in double.  Used for CDFs: a parallel fp32 scan associates differently per lane, so
// prefixes that should be equal (increments below 1 ulp, i.e. empty space) can differ by an ulp or even
// decrease; accumulating in double and rounding once keeps the CDF monotone with exact ties -- and is
// what the reference's torch.cumsum does on the CPU (it accumulates float tensors in double).
template <int G>
__device__ __forceinline__ double group_incl_sum_d(double v, int gl) {
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
        double n = __shfl_up_sync(FULL, v, o, G);
        if (gl >= o) v += n;
    }
    return v;
}
template <int G>
__device__ __forceinline__ float group_incl_prod(float v, int gl) {
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
        float n = __shfl_up_sync(FULL, v, o, G);
        if (gl >= o) v *= n;
    }
    return v;
}
// inclusive suffix sum (lane gl gets sum of lanes >= gl)
template <int G>
__device__ __forceinline__ float group_suffix_sum(float v, int gl) {
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
        float n = __shfl_down_sync(FULL, v, o, G);
        if (gl + o < G) v += n;
    }
    return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
// torch.nn.functional.softplus (beta=1, threshold=20)
__device__ __forceinline__ float softplusf_(float x) { return x > 20.0f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float normal_cdff_(float x) {
    return 0.5f * (1.0f + erff(x / 1.41421354f));      // math_utils.py:193-200, sqrt(2) in fp32
}

// ---- SFU helpers (MUFU ex2 / lg2 / rcp, <= 2 ulp) ------------------------------------------------
__device__ __forceinline__ float ex2_(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
constexpr float L2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
// a / b for normal-range b: reciprocal + one residual correction = the correctly rounded quotient except in
// rare double-rounding cases; no special-case slow path (0/0 and x/0 give NaN, callers handle it).
__device__ __forceinline__ float div_fast(float a, float b) {
    float r = rcp_(b);
    float q = a * r;
    return fmaf(fmaf(-q, b, a), r, q);
}

// Branch-free binary search over a shared-memory table of REC-byte records (positions 1 .. 2*STEP-1): one
// shared load, one compare and one predicated add per step -- the running position is a byte address and the
// step an immediate offset of the load.  M independent searches advance in lock step so that their shared-
// memory latencies overlap (one search alone is a chain of log2(P) dependent loads).  STRICT: count keys < x,
// else keys <= x.  at[m] enters as the shared address of record 0's key and leaves as the address of record
// (count)'s key.
template <int STEP, bool STRICT, int M, int REC>
struct SmemSearch {
    __device__ static __forceinline__ void run(unsigned (&at)[M], const float (&x)[M]) {
        float v[M];
#pragma unroll
        for (int m = 0; m < M; ++m)
            asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v[m]) : "r"(at[m]), "n"(STEP * REC) : "memory");
#pragma unroll
        for (int m = 0; m < M; ++m)
            if (STRICT ? (v[m] < x[m]) : (v[m] <= x[m])) at[m] += STEP * REC;
        SmemSearch<STEP / 2, STRICT, M, REC>::run(at, x);
    }
};
template <bool STRICT, int M, int REC>
struct SmemSearch<0, STRICT, M, REC> {
    __device__ static __forceinline__ void run(unsigned (&)[M], const float (&)[M]) {}
};

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace ddnerf
