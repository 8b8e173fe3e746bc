// K3: samplers -- stratified first-cycle samples and the two coarse-to-fine inverse-CDF
// resamplers, one warp per ray.
//
// Replaces models/samplers.py:30-62 (sample_first_cycle), :64-121 (sample_pdf) and :124-215
// (sample_pdf_with_mu_sigma) of the reference.  The reference materialises an [N,S+1,n] boolean
// mask plus four float temporaries of that shape to locate each sample's interval; here a warp
// stages its ray's weights/bins in shared memory, builds the smoothed CDF with a warp scan and
// binary-searches it per sample: O(S + n log S) work and bins+weights+u in, samples out of HBM.
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace ddnerf {
namespace {

// ------------------------------------------------------------------------------------------
// sample_first_cycle: elementwise
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float linspace01(int idx, int steps, float step) {   // torch.linspace(0,1,steps)[idx]
    return idx < steps / 2 ? step * (float)idx : 1.0f - step * (float)(steps - idx - 1);
}

// Element-parallel over the FLAT [N, S+1] array: a thread owns four or eight consecutive fence-posts (16-byte loads of the
// random numbers, all in flight before the first use; 16-byte stores), finds its ray with one multiply-shift division
// (Granlund-Montgomery: the divisor S+1 is a launch constant) and carries the three neighbouring linspace values along the
// row, re-seeding them when its elements cross into the next ray.  ~35 instructions per fence-post; the round-1 kernel
// (one warp per four rays, neighbours by shuffle) spent ~150: its 32-lane chunks left 20 % of the lanes idle at
// S+1 = 2^m + 1 and its rows of 4 (S+1) bytes were never sector-aligned.  Same arithmetic, same rounding.
struct FastDiv { uint32_t magic; int sh1, sh2; };
__device__ __forceinline__ uint32_t fast_div(uint32_t n, const FastDiv& d) {
    const uint32_t t = __umulhi(d.magic, n);
    return (t + ((n - t) >> d.sh1)) >> d.sh2;
}

template <bool LINDISP, int V>
__global__ void __launch_bounds__(256) first_cycle_flat_kernel(const float* __restrict__ near, const float* __restrict__ far,
                                                                int64_t ray_stride, const float* __restrict__ t_rand,
                                                                float* __restrict__ out, uint32_t total, int S,
                                                                float step, FastDiv dv, int vec_ok) {
    constexpr bool lindisp = LINDISP;
    // V fence-posts per thread: 8 (two 16-byte loads in flight per thread) from 4 M elements up, 4 below (more threads)
    const uint32_t e0 = (blockIdx.x * 256u + threadIdx.x) * (uint32_t)V;
    if (e0 >= total) return;
    const uint32_t steps = (uint32_t)S + 1u;
    const bool full = e0 + (uint32_t)(V - 1) < total;
    float rnd[V];
#pragma unroll
    for (int j = 0; j < V; ++j) rnd[j] = 0.f;
    if (t_rand) {
        if (vec_ok && full) {
#pragma unroll
            for (int q = 0; q < V / 4; ++q) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(t_rand + e0) + q);
                rnd[4 * q] = v.x; rnd[4 * q + 1] = v.y; rnd[4 * q + 2] = v.z; rnd[4 * q + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < V; ++j) if (e0 + j < total) rnd[j] = __ldg(t_rand + e0 + j);
        }
    }
    uint32_t ray = fast_div(e0, dv);
    int i = (int)(e0 - ray * steps);
    float nr, fr, inr, ifr;
    auto load_ray = [&](uint32_t r) {
        nr = __ldg(near + (int64_t)r * ray_stride); fr = __ldg(far + (int64_t)r * ray_stride);
        inr = lindisp ? 1.0f / nr : 0.f; ifr = lindisp ? 1.0f / fr : 0.f;
    };
    auto tv = [&](int k) {                               // samplers.py:32-41 at fence-post k (clamped: edge values are unused)
        k = min(max(k, 0), S);
        const float sfrac = linspace01(k, (int)steps, step);
        return lindisp ? 1.0f / (inr * (1.0f - sfrac) + ifr * sfrac) : nr * (1.0f - sfrac) + fr * sfrac;
    };
    load_ray(ray);
    float prev = tv(i - 1), cur = tv(i), nxt = tv(i + 1);
    float res[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
        float t = cur;
        if (t_rand) {                                    // samplers.py:52-60
            const float lower = i == 0 ? cur : 0.5f * (cur + prev);
            const float upper = i == S ? cur : 0.5f * (nxt + cur);
            t = lower + (upper - lower) * rnd[j];
            if (i == 0) t = nr;
            if (i == S) t = fr;
        }
        res[j] = t;
        if (j < V - 1) {
            if (i < S) { ++i; prev = cur; cur = nxt; nxt = tv(i + 1); }
            else if (e0 + j + 1 < total) { ++ray; i = 0; load_ray(ray); prev = cur = tv(0); nxt = tv(1); }
        }
    }
    if (vec_ok && full) {
#pragma unroll
        for (int q = 0; q < V / 4; ++q)
            reinterpret_cast<float4*>(out + e0)[q] = make_float4(res[4 * q], res[4 * q + 1], res[4 * q + 2], res[4 * q + 3]);
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) if (e0 + j < total) out[e0 + j] = res[j];
    }
}

// any S: element-parallel (no prefetch)
__global__ void first_cycle_generic_kernel(const float* __restrict__ near, const float* __restrict__ far, int64_t ray_stride,
                                           const float* __restrict__ t_rand, float* __restrict__ out, int64_t N, int S,
                                           int lindisp) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N * (S + 1)) return;
    int64_t ray = e / (S + 1);
    int i = (int)(e - ray * (S + 1));
    const float nr = __ldg(near + ray * ray_stride), fr = __ldg(far + ray * ray_stride);
    const float step = 1.0f / (float)S;
    auto tv = [&](int k) {
        float s = linspace01(k, S + 1, step);
        return lindisp ? 1.0f / (1.0f / nr * (1.0f - s) + 1.0f / fr * s) : nr * (1.0f - s) + fr * s;
    };
    float t = tv(i);
    if (t_rand) {                                                       // samplers.py:52-60
        float lower = i == 0 ? t : 0.5f * (t + tv(i - 1));
        float upper = i == S ? t : 0.5f * (tv(i + 1) + t);
        t = lower + (upper - lower) * __ldg(t_rand + e);
        if (i == 0) t = nr;
        if (i == S) t = fr;
    }
    out[e] = t;
}

// ------------------------------------------------------------------------------------------
// shared: smoothed-weight CDF of one ray, built by one warp in shared memory
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) { return group_sum<32>(v); }

// wpad[S+2], what[S], cdf[S+1] in shared memory.  samplers.py:69-91.
__device__ void build_cdf(const float* __restrict__ w_row, float* wpad, float* what, float* cdf, int S, int lane,
                          int pdf_padding) {
    for (int i = lane; i < S; i += 32) wpad[i + 1] = __ldg(w_row + i);
    __syncwarp();
    if (lane == 0) { wpad[0] = wpad[1]; wpad[S + 1] = wpad[S]; }
    __syncwarp();
    float part = 0.f;
    for (int i = lane; i < S; i += 32) {
        float prev = wpad[i], cur = wpad[i + 1], next = wpad[i + 2], v;
        if (pdf_padding) v = 0.5f * (fmaxf(prev, cur) + fmaxf(cur, next)) + 0.01f;
        else v = 0.8f * cur + 0.1f * prev + 0.1f * next + 0.01f;
        what[i] = v;
        part += v;
    }
    const float total = warp_sum(part);
    __syncwarp();
    double carry = 0.0;
    for (int base = 0; base < S - 1; base += 32) {
        int i = base + lane;
        float v = i < S - 1 ? what[i] / total : 0.f;
        double incl = group_incl_sum_d<32>((double)v, lane) + carry;
        if (i < S - 1) cdf[i + 1] = fminf(1.0f, (float)incl);
        carry = __shfl_sync(FULL, incl, 31);
    }
    if (lane == 0) { cdf[0] = 0.f; cdf[S] = 1.f; }
    __syncwarp();
}

// #{m in [0,len) : v[m] <= u}
__device__ __forceinline__ int count_le(const float* v, int len, float u) {
    int lo = 0, hi = len;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (v[mid] <= u) lo = mid + 1; else hi = mid;
    }
    return lo;
}

struct USpec {          // how u is produced (samplers.py:94-104 / 155-171)
    int det;            // 1: linspace(0, hi, n)
    float hi;           // 1.0 (mip) or 0.9999 (dd)
    float stride;       // (float)(1/n) (mip) or (float)(1/(n-1)) (dd)
    float div;          // (float)(n + 1e-5)
    int clamp0;         // dd: max(u, 0)
};

__device__ __forceinline__ float make_u(const USpec& us, int k, int n, const float* rand_row) {
    if (us.det) {
        float step = us.hi / (float)(n - 1);
        return k < n / 2 ? step * (float)k : us.hi - step * (float)(n - k - 1);
    }
    float u = (float)k * us.stride + __ldg(rand_row + k) / us.div;
    u = fminf(u, 0.9999f);
    if (us.clamp0) u = fmaxf(u, 0.0f);
    return u;
}

// ------------------------------------------------------------------------------------------
// Fast path (S <= 256): G lanes per ray, C consecutive cells per lane (G*C >= S, a power of two).
//   * the smoothed pdf lives in registers; its CDF is a per-lane serial sum plus ONE lane-group scan, both
//     in double (monotone, ties exact -- what torch.cumsum on the CPU produces), instead of one scan per
//     32 cells;
//   * {cdf, bin} pairs are staged once in shared memory (8 B per fence-post);
//   * each sample's interval comes from a branch-free binary search, log2(G*C) predicated steps;
//   * divisions are reciprocal + one residual correction (div_fast), no slow path.
// Round-1 profile of the generic kernels below: 313 thread instructions per sample, issue-active 88 %
// (profiles/r01_ncu_sample_pdf_*.md); they remain the fallback for S > 256.
// ------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ double group_incl_sum_d2(double v, int gl) { return group_incl_sum_d<G>(v, gl); }

template <int G, int C>
struct FastCdf {
    static constexpr int P = G * C;                       // searched positions 1..P-1
    // cb[0..P]: x = cdf, y = bin.  Entries S+1..P hold +inf / 0.
    __device__ static __forceinline__ void build(const float* __restrict__ w_row, const float* __restrict__ bins_row,
                                                 float2* cb, int S, int gl, int pdf_padding) {
        float w[C];
        const int e0 = gl * C;
#pragma unroll
        for (int c = 0; c < C; ++c) w[c] = __ldg(w_row + min(e0 + c, S - 1));
        float bv[C + 1];                                  // fence-posts q = gl + c*G, all requested before the first use
#pragma unroll
        for (int c = 0; c <= C; ++c) { const int q = gl + c * G; bv[c] = q <= S ? __ldg(bins_row + q) : 0.f; }
#pragma unroll
        for (int c = 0; c <= C; ++c) {
            const int q = gl + c * G;
            if (q <= P) {
                cb[q].y = bv[c];
                if (q > S) cb[q].x = __int_as_float(0x7f800000);
            }
        }
        float prev = __shfl_up_sync(FULL, w[C - 1], 1, G);
        float next = __shfl_down_sync(FULL, w[0], 1, G);
        if (gl == 0) prev = w[0];                         // samplers.py:70-72 replicate padding
        if (gl == G - 1) next = w[C - 1];
        float v[C], part = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float p = c ? w[c > 0 ? c - 1 : 0] : prev, nx = c < C - 1 ? w[c < C - 1 ? c + 1 : 0] : next, cur = w[c];
            float x;
            if (pdf_padding) x = 0.5f * (fmaxf(p, cur) + fmaxf(cur, nx)) + 0.01f;
            else x = 0.8f * cur + 0.1f * p + 0.1f * nx + 0.01f;
            v[c] = e0 + c < S ? x : 0.f;
            part += v[c];
        }
        const float total = group_sum<G>(part);
        double incl[C], run = 0.0;
#pragma unroll
        for (int c = 0; c < C; ++c) { run += (double)div_fast(v[c], total); incl[c] = run; }
        double lanes_incl = group_incl_sum_d<G>(run, gl);
        double off = __shfl_up_sync(FULL, lanes_incl, 1, G);
        if (gl == 0) off = 0.0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int m = e0 + c + 1;                     // cdf[m] = min(1, sum_{i<m} pdf_i), m = 1..S-1
            if (m <= S - 1) cb[m].x = fminf(1.0f, (float)(off + incl[c]));
        }
        if (gl == 0) { cb[0].x = 0.f; cb[S].x = 1.f; }
        __syncwarp();
    }
    // j[m] = #{q in [0,S] : cdf[q] <= u[m]} - 1  (u >= 0), M searches in lock step (common.cuh: SmemSearch)
    template <int M>
    __device__ static __forceinline__ void search(const float2* cb, int S, const float (&u)[M], int (&j)[M]) {
        const unsigned base = (unsigned)__cvta_generic_to_shared(cb);
        unsigned at[M];
#pragma unroll
        for (int m = 0; m < M; ++m) at[m] = base;
        SmemSearch<P / 2, false, M, 8>::run(at, u);
#pragma unroll
        for (int m = 0; m < M; ++m) {
            j[m] = (int)((at[m] - base) >> 3);
            if (S == P && u[m] >= 1.0f) j[m] = S;         // position P (= S) is outside the searched range
        }
    }
};

struct UGen {
    float det_step, rdiv; USpec us; int n;
    __device__ UGen(const USpec& u, int n_) : us(u), n(n_) { det_step = us.hi / (float)(n - 1); rdiv = rcp_(us.div); }
    __device__ __forceinline__ float operator()(int k, float rnd) const {
        if (us.det) return k < n / 2 ? det_step * (float)k : us.hi - det_step * (float)(n - k - 1);
        const float q = rnd * rdiv;                       // rnd / div, reciprocal + residual correction (div_fast)
        float u = (float)k * us.stride + fmaf(fmaf(-q, us.div, rnd), rdiv, q);
        u = fminf(u, 0.9999f);
        if (us.clamp0) u = fmaxf(u, 0.0f);
        return u;
    }
};

// K: output chunks per lane (n <= K*G).  Every global load of the ray -- weights, bins, the K random numbers --
// is requested before the first dependent instruction; the kernels are otherwise bound by one DRAM latency
// per chunk.
template <int G, int C, int K>
__global__ void __launch_bounds__(256) sample_pdf_fast_kernel(const float* __restrict__ bins, const float* __restrict__ weights,
                                                               const float* __restrict__ rand, float* __restrict__ out,
                                                               int32_t* __restrict__ idx_out, int64_t N, int S, int n,
                                                               int pdf_padding, USpec us) {
    using F = FastCdf<G, C>;
    __shared__ float2 cbs[256 / G][F::P + 1];
    const int grp = threadIdx.x / G, gl = threadIdx.x % G;
    int64_t ray = (int64_t)blockIdx.x * (256 / G) + grp;
    const bool valid = ray < N;
    if (!valid) ray = N - 1;                                             // keep the warp convergent for the shuffles
    float2* cb = cbs[grp];
    float rnd[K];
#pragma unroll
    for (int c = 0; c < K; ++c) rnd[c] = (rand && c * G + gl < n) ? __ldg(rand + ray * n + c * G + gl) : 0.f;
    F::build(weights + ray * S, bins + ray * (S + 1), cb, S, gl, pdf_padding);
    if (!valid) return;
    const UGen ugen(us, n);
    float* orow = out + ray * n;
    float u[K];
    int j[K];
#pragma unroll
    for (int c = 0; c < K; ++c) u[c] = ugen(c * G + gl, rnd[c]);
    F::search(cb, S, u, j);                                              // all K interval searches in lock step
#pragma unroll
    for (int c = 0; c < K; ++c) {
        const int k = c * G + gl;
        if (k >= n) break;
        const float2 a0 = cb[j[c]], a1 = cb[min(j[c] + 1, S)];
        float t = div_fast(u[c] - a0.x, a1.x - a0.x);
        if (t != t) t = 0.f;                                             // nan_to_num(., 0)
        t = fminf(fmaxf(t, 0.f), 1.f);
        orow[k] = a0.y + t * (a1.y - a0.y);
        if (idx_out) idx_out[ray * n + k] = j[c];
    }
}

// Fused DDNeRF coarse path: part_inside / left_tail are NULL and `sigmas` holds the UNSMOOTHED sigmas; the kernel applies
// gaussian_smooth_factor (a kernel argument, or read from device memory when the step is a replayed CUDA graph) and
// evaluates the two tails per cell with the reference's own fp32 formula, models.py:268-273 / math_utils.py:193-200:
// {part_inside, left_tail, smoothed sigma, mu}.
struct Smooth { float value; const float* dev; };
__device__ __forceinline__ float4 dd_cell(float mu, float sigma, float smooth) {
    const float ss = sigma * smooth;
    const float lt = normal_cdff_((0.0f - mu) / ss);
    const float pin = normal_cdff_((1.0f - mu) / ss) - lt;
    return make_float4(pin, lt, ss, mu);
}

// DDNeRF variant.  The per-cell Gaussian parameters (part_inside, left_tail, sigma, mu) are staged in shared
// memory next to the {cdf, bin} pairs -- gathered from global memory they put one DRAM latency into every
// output chunk.  Dynamic shared memory per ray: P float4 + (P+1) float2 (rounded to 16 bytes).
template <int G, int C>
__host__ __device__ constexpr int dd_ray_smem_bytes() { return G * C * 16 + ((G * C + 1) * 8 + 15) / 16 * 16; }

template <int G, int C, int K>
__global__ void __launch_bounds__(256) sample_pdf_mu_sigma_fast_kernel(
    const float* __restrict__ bins, const float* __restrict__ weights, const float* __restrict__ mus,
    const float* __restrict__ sigmas, const float* __restrict__ part_inside, const float* __restrict__ left_tail,
    const float* __restrict__ rand, float* __restrict__ out, int32_t* __restrict__ idx_out, int64_t N, int S, int n,
    int pdf_padding, float near_cfg, float far_cfg, USpec us, Smooth smooth) {
    using F = FastCdf<G, C>;
    extern __shared__ float4 dd_smem[];
    const int grp = threadIdx.x / G, gl = threadIdx.x % G;
    int64_t ray = (int64_t)blockIdx.x * (256 / G) + grp;
    const bool valid = ray < N;
    if (!valid) ray = N - 1;
    float4* pr = dd_smem + (size_t)grp * (dd_ray_smem_bytes<G, C>() / 16);    // {pin, lt, sigma, mu} per cell
    float2* cb = reinterpret_cast<float2*>(pr + F::P);
    float rnd[K];
#pragma unroll
    for (int c = 0; c < K; ++c) rnd[c] = (rand && c * G + gl < n) ? __ldg(rand + ray * n + c * G + gl) : 0.f;
    {
        const float* mu_r = mus + ray * S;
        const float* sg_r = sigmas + ray * S;
        float4 pv[C];
        if (part_inside) {
            const float* pin_r = part_inside + ray * S;
            const float* lt_r = left_tail + ray * S;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int q = min(gl + c * G, S - 1);
                pv[c] = make_float4(__ldg(pin_r + q), __ldg(lt_r + q), __ldg(sg_r + q), __ldg(mu_r + q));
            }
        } else {
            const float sm = smooth.dev ? __ldg(smooth.dev) : smooth.value;
            float mv[C], sv[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int q = min(gl + c * G, S - 1);
                mv[c] = __ldg(mu_r + q); sv[c] = __ldg(sg_r + q);
            }
#pragma unroll
            for (int c = 0; c < C; ++c) pv[c] = dd_cell(mv[c], sv[c], sm);
        }
        F::build(weights + ray * S, bins + ray * (S + 1), cb, S, gl, pdf_padding);   // ends with __syncwarp
#pragma unroll
        for (int c = 0; c < C; ++c) pr[gl + c * G] = pv[c];
        __syncwarp();
    }
    const UGen ugen(us, n);
    float* orow = out + ray * n;
    float u[K];
    int j[K];
#pragma unroll
    for (int c = 0; c < K; ++c) u[c] = ugen(c * G + gl, rnd[c]);
    if (S == 1) {
#pragma unroll
        for (int c = 0; c < K; ++c) j[c] = 0;
    } else {
        F::search(cb, S, u, j);
    }
    bool ok = true;
    float prev_last = -__int_as_float(0x7f800000);
#pragma unroll
    for (int c = 0; c < K; ++c) {                                        // uniform trip count over the warp
        const int k = c * G + gl;
        if (c * G >= n) break;
        const bool in = k < n;
        float val = __int_as_float(0x7f800000);
        if (in) {
            float z, b0, b1;
            int ind;
            if (S == 1) {                                                // samplers.py:185-190
                ind = 0; b0 = cb[0].y; b1 = cb[1].y;
                z = u[c] * pr[0].x + pr[0].y;
            } else {
                const float2 a0 = cb[j[c]], a1 = cb[min(j[c] + 1, S)];
                b0 = a0.y; b1 = a1.y;
                ind = j[c];                                              // torch.max: first index of the maximum
                while (ind > 0 && cb[ind - 1].y == cb[ind].y) --ind;
                ind = min(ind, S - 1);
                z = div_fast(u[c] - a0.x, a1.x - a0.x) * pr[ind].x + pr[ind].y;
                z = fminf(z, 0.999f);
            }
            const float4 pp = pr[ind];
            z = 1.41421354f * erfinvf(2.0f * z - 1.0f);                  // math_utils.py:202-208
            float t = fminf(fmaxf(z * pp.z + pp.w, 0.f), 0.99999f);
            val = b0 + t * (b1 - b0);
            if (k == 0) val = near_cfg;                                  // samplers.py:210-211
            if (k == n - 1) val = far_cfg;
            if (valid) {
                orow[k] = val;
                if (idx_out) idx_out[ray * n + k] = ind;
            }
        }
        // samplers.py:213 sorts; the samples are monotone by construction unless bins leave [near_cfg, far_cfg]
        float nb = __shfl_down_sync(FULL, val, 1, G);
        if (gl == G - 1) nb = __int_as_float(0x7f800000);
        if (val > nb) ok = false;
        if (gl == 0 && prev_last > val) ok = false;
        prev_last = __shfl_sync(FULL, val, G - 1, G);
    }
    // any lane of the group saw an inversion -> lane 0 of the group insertion-sorts the row in place
    unsigned bad = __ballot_sync(FULL, !ok);
    const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u)) << ((threadIdx.x & 31) / G * G);
    if ((bad & gmask) && valid) {
        __syncwarp(gmask);
        if (gl == 0) {
            for (int a = 1; a < n; ++a) {
                float x = orow[a];
                int b = a - 1;
                while (b >= 0 && orow[b] > x) { orow[b + 1] = orow[b]; --b; }
                orow[b + 1] = x;
            }
        }
    }
}

__global__ void sample_pdf_kernel(const float* __restrict__ bins, const float* __restrict__ weights,
                                  const float* __restrict__ rand, float* __restrict__ out, int32_t* __restrict__ idx_out,
                                  int64_t N, int S, int n, int pdf_padding, USpec us, int per_warp_floats) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (ray >= N) return;                                                // whole warp exits together
    float* wpad = smem + (size_t)warp * per_warp_floats;
    float* what = wpad + (S + 2);
    float* cdf = what + S;
    float* sb = cdf + (S + 1);
    for (int i = lane; i <= S; i += 32) sb[i] = __ldg(bins + ray * (S + 1) + i);
    build_cdf(weights + ray * S, wpad, what, cdf, S, lane, pdf_padding);
    for (int k = lane; k < n; k += 32) {
        float u = make_u(us, k, n, rand ? rand + ray * n : nullptr);
        int j = max(count_le(cdf, S + 1, u) - 1, 0);
        int j1 = min(j + 1, S);
        float c0 = cdf[j], c1 = cdf[j1], b0 = sb[j], b1 = sb[j1];
        float t = (u - c0) / (c1 - c0);
        if (t != t) t = 0.f;                                             // nan_to_num(., 0)
        t = fminf(fmaxf(t, 0.f), 1.f);
        out[ray * n + k] = b0 + t * (b1 - b0);
        if (idx_out) idx_out[ray * n + k] = j;
    }
}

// in-place ascending bitonic sort of v[0..len_pow2) by one warp
__device__ void warp_bitonic_sort(float* v, int len_pow2, int lane) {
    for (int k = 2; k <= len_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < len_pow2; i += 32) {
                int p = i ^ j;
                if (p > i) {
                    float a = v[i], b = v[p];
                    bool up = (i & k) == 0;
                    if ((a > b) == up) { v[i] = b; v[p] = a; }
                }
            }
            __syncwarp();
        }
    }
}

__global__ void sample_pdf_mu_sigma_kernel(const float* __restrict__ bins, const float* __restrict__ weights,
                                           const float* __restrict__ mus, const float* __restrict__ sigmas,
                                           const float* __restrict__ part_inside, const float* __restrict__ left_tail,
                                           const float* __restrict__ rand, float* __restrict__ out,
                                           int32_t* __restrict__ idx_out, int64_t N, int S, int n, int n_pow2,
                                           int pdf_padding, float near_cfg, float far_cfg, USpec us,
                                           int per_warp_floats, Smooth smooth) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (ray >= N) return;
    const bool fused = part_inside == nullptr;
    const float smf = fused ? (smooth.dev ? __ldg(smooth.dev) : smooth.value) : 1.0f;
    float* wpad = smem + (size_t)warp * per_warp_floats;
    float* what = wpad + (S + 2);
    float* cdf = what + S;
    float* sb = cdf + (S + 1);
    float* so = sb + (S + 1);                                            // n_pow2 output slots
    for (int i = lane; i <= S; i += 32) sb[i] = __ldg(bins + ray * (S + 1) + i);
    build_cdf(weights + ray * S, wpad, what, cdf, S, lane, pdf_padding);
    const float* mu_r = mus + ray * S;
    const float* sg_r = sigmas + ray * S;
    const float* pin_r = fused ? nullptr : part_inside + ray * S;
    const float* lt_r = fused ? nullptr : left_tail + ray * S;
    auto cell = [&](int q) {                                             // {part_inside, left_tail, sigma, mu} of cell q
        return fused ? dd_cell(__ldg(mu_r + q), __ldg(sg_r + q), smf)
                     : make_float4(__ldg(pin_r + q), __ldg(lt_r + q), __ldg(sg_r + q), __ldg(mu_r + q));
    };
    for (int k = lane; k < n_pow2; k += 32) {
        float val = __int_as_float(0x7f800000);                          // +inf padding for the sort
        if (k < n) {
            float u = make_u(us, k, n, rand ? rand + ray * n : nullptr);
            float z, b0, b1;
            int ind;
            float4 pp;
            if (S == 1) {                                                // samplers.py:185-190
                ind = 0; b0 = sb[0]; b1 = sb[1];
                pp = cell(0);
                z = u * pp.x + pp.y;
            } else {
                int j = max(count_le(cdf, S + 1, u) - 1, 0);
                int j1 = min(j + 1, S);
                b0 = sb[j]; b1 = sb[j1];
                float c0 = cdf[j], c1 = cdf[j1];
                ind = j;                                                 // torch.max: first index of the maximum
                while (ind > 0 && sb[ind - 1] == sb[ind]) --ind;
                ind = min(ind, S - 1);
                pp = cell(ind);
                z = ((u - c0) / (c1 - c0)) * pp.x + pp.y;
                z = fminf(z, 0.999f);
            }
            z = 1.41421354f * erfinvf(2.0f * z - 1.0f);                  // math_utils.py:202-208
            float t = fminf(fmaxf(z * pp.z + pp.w, 0.f), 0.99999f);
            val = b0 + t * (b1 - b0);
            if (k == 0) val = near_cfg;                                  // samplers.py:210-211
            if (k == n - 1) val = far_cfg;
            if (idx_out) idx_out[ray * n + k] = ind;
        }
        so[k] = val;
    }
    __syncwarp();
    // the samples are monotone by construction unless bins violate [near_cfg, far_cfg]; sort only then
    bool ok = true;
    for (int k = lane; k < n - 1; k += 32) ok = ok && !(so[k] > so[k + 1]);
    if (!__all_sync(FULL, ok)) warp_bitonic_sort(so, n_pow2, lane);
    __syncwarp();
    for (int k = lane; k < n; k += 32) out[ray * n + k] = so[k];
}

__global__ void find_interval_kernel(const float* __restrict__ cdf, const float* __restrict__ u, int32_t* __restrict__ idx,
                                     int64_t N, int S, int n) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N * n) return;
    int64_t ray = e / n;
    idx[e] = max(count_le(cdf + ray * (S + 1), S + 1, __ldg(u + e)) - 1, 0);
}

int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// (G lanes per ray, C cells per lane, K output chunks per lane) of the fast path; false -> generic kernels
template <typename F>
bool dispatch_fast(int S, int n, F&& f) {
#define DDNERF_FAST(G, C)                                                                                                   \
    if (S <= G * C) {                                                                                                       \
        if (n <= 5 * G) { f(std::integral_constant<int, G>{}, std::integral_constant<int, C>{}, std::integral_constant<int, 5>{}); return true; } \
        if (n <= 9 * G) { f(std::integral_constant<int, G>{}, std::integral_constant<int, C>{}, std::integral_constant<int, 9>{}); return true; } \
        return false;                                                                                                       \
    }
    // Four cells per lane: measured faster than eight (profiles/r01b_*): these kernels are bound by latency, and
    // the extra resident warps of the smaller per-lane state outweigh the shorter scans of fewer, fatter lanes.
    DDNERF_FAST(4, 1) DDNERF_FAST(4, 2) DDNERF_FAST(4, 4) DDNERF_FAST(8, 4) DDNERF_FAST(16, 4) DDNERF_FAST(32, 4) DDNERF_FAST(32, 8)
#undef DDNERF_FAST
    return false;
}

}  // namespace
}  // namespace ddnerf

using namespace ddnerf;

extern "C" DDNERF_EXPORT int ddnerf_sample_first_cycle(const float* near, const float* far, int64_t ray_stride, const float* t_rand,
                                         float* t_out, int64_t N, int S, int lindisp, void* stream) {
    DDNERF_CHECK_ARG(N == 0 || (near && far && t_out), "sample_first_cycle: null pointer");
    DDNERF_CHECK_ARG(S >= 1, "sample_first_cycle: S=%d < 1", S);
    if (N == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t total = N * (int64_t)(S + 1);
    if (total < (int64_t)0xffff0000u && S < 65536) {
        const uint32_t d = (uint32_t)S + 1u;             // multiply-shift division by S+1 (exact for every 32-bit dividend)
        int l = 0;
        while ((1ull << l) < d) ++l;
        FastDiv dv{(uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1ull), l < 1 ? l : 1, l > 1 ? l - 1 : 0};
        const int vec_ok = (reinterpret_cast<uintptr_t>(t_out) % 32 == 0) && (!t_rand || reinterpret_cast<uintptr_t>(t_rand) % 32 == 0);
        const float step = 1.0f / (float)S;
        auto launch = [&](auto lin, auto vv) {
            constexpr bool L = decltype(lin)::value;
            constexpr int V = decltype(vv)::value;
            first_cycle_flat_kernel<L, V><<<ceil_div(ceil_div(total, V), 256), 256, 0, st>>>(near, far, ray_stride, t_rand, t_out,
                                                                                          (uint32_t)total, S, step, dv, vec_ok);
        };
        const bool big = total >= (1 << 22);
        if (lindisp) { if (big) launch(std::true_type{}, std::integral_constant<int, 8>{}); else launch(std::true_type{}, std::integral_constant<int, 4>{}); }
        else { if (big) launch(std::false_type{}, std::integral_constant<int, 8>{}); else launch(std::false_type{}, std::integral_constant<int, 4>{}); }
    }
    else first_cycle_generic_kernel<<<ceil_div(N * (S + 1), 256), 256, 0, st>>>(near, far, ray_stride, t_rand, t_out, N, S, lindisp);
    DDNERF_LAUNCHED("sample_first_cycle", 1);
    return 0;
}

static int warps_per_block(int per_warp_floats) {
    int w = (int)(48 * 1024 / ((size_t)per_warp_floats * sizeof(float)));
    return w > 8 ? 8 : w;
}

extern "C" DDNERF_EXPORT int ddnerf_sample_pdf(const float* bins, const float* weights, const float* rand, float* out,
                                 int32_t* idx_out, int64_t N, int S, int n, int pdf_padding, void* stream) {
    DDNERF_CHECK_ARG(N == 0 || (bins && weights && out), "sample_pdf: null pointer");
    DDNERF_CHECK_ARG(S >= 1 && S <= 2048 && n >= 2, "sample_pdf: S=%d n=%d unsupported", S, n);
    if (N == 0) return 0;
    double s = 1.0 / n;
    USpec us{rand == nullptr, 1.0f, (float)s, (float)((1.0 / s) + 1e-5), 0};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    bool fast = dispatch_fast(S, n, [&](auto g, auto c, auto kk) {
        constexpr int G = decltype(g)::value, C = decltype(c)::value, K = decltype(kk)::value;
        sample_pdf_fast_kernel<G, C, K><<<ceil_div(N, 256 / G), 256, 0, st>>>(bins, weights, rand, out, idx_out, N, S, n,
                                                                          pdf_padding, us);
    });
    if (!fast) {
        int per_warp = (S + 2) + S + (S + 1) + (S + 1);
        int wpb = warps_per_block(per_warp);
        DDNERF_CHECK_ARG(wpb >= 1, "sample_pdf: S=%d needs too much shared memory", S);
        sample_pdf_kernel<<<ceil_div(N, wpb), wpb * 32, (size_t)wpb * per_warp * sizeof(float), st>>>(
            bins, weights, rand, out, idx_out, N, S, n, pdf_padding, us, per_warp);
    }
    DDNERF_LAUNCHED("sample_pdf", 1);
    return 0;
}

static int sample_dd_impl(const float* bins, const float* weights, const float* mus, const float* sigmas,
                          const float* part_inside, const float* left_tail, Smooth smooth, const float* rand, float* out,
                          int32_t* idx_out, int64_t N, int S, int n, int pdf_padding, float near_cfg, float far_cfg,
                          void* stream) {
    DDNERF_CHECK_ARG(S >= 1 && S <= 2048 && n >= 2 && n <= 4096, "sample_pdf_mu_sigma: S=%d n=%d unsupported", S, n);
    if (N == 0) return 0;
    double s = 1.0 / (n - 1);
    USpec us{rand == nullptr, 0.9999f, (float)s, (float)(n + 1e-5), 1};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    bool fast = dispatch_fast(S, n, [&](auto g, auto c, auto kk) {
        constexpr int G = decltype(g)::value, C = decltype(c)::value, K = decltype(kk)::value;
        auto kern = sample_pdf_mu_sigma_fast_kernel<G, C, K>;
        constexpr int bytes = (256 / G) * dd_ray_smem_bytes<G, C>();
        if (bytes > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        kern<<<ceil_div(N, 256 / G), 256, bytes, st>>>(bins, weights, mus, sigmas, part_inside, left_tail, rand, out,
                                                       idx_out, N, S, n, pdf_padding, near_cfg, far_cfg, us, smooth);
    });
    if (!fast) {
        int np2 = next_pow2(n);
        int per_warp = (S + 2) + S + (S + 1) + (S + 1) + np2;
        int wpb = warps_per_block(per_warp);
        if (wpb < 1) {                      // the largest shapes (S = 2048, n up to 4096: 49 KB per ray) pass the 48 KB default
            wpb = 1;
            const size_t bytes = (size_t)per_warp * sizeof(float);
            DDNERF_CHECK_ARG(bytes <= 200 * 1024, "sample_pdf_mu_sigma: S=%d n=%d needs too much shared memory", S, n);
            cudaFuncSetAttribute(sample_pdf_mu_sigma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        }
        sample_pdf_mu_sigma_kernel<<<ceil_div(N, wpb), wpb * 32, (size_t)wpb * per_warp * sizeof(float), st>>>(
            bins, weights, mus, sigmas, part_inside, left_tail, rand, out, idx_out, N, S, n, np2, pdf_padding, near_cfg,
            far_cfg, us, per_warp, smooth);
    }
    DDNERF_LAUNCHED("sample_pdf_mu_sigma", 1);
    return 0;
}

extern "C" DDNERF_EXPORT int ddnerf_sample_pdf_mu_sigma(const float* bins, const float* weights, const float* mus, const float* sigmas,
                                          const float* part_inside, const float* left_tail, const float* rand, float* out,
                                          int32_t* idx_out, int64_t N, int S, int n, int pdf_padding, float near_cfg,
                                          float far_cfg, void* stream) {
    DDNERF_CHECK_ARG(N == 0 || (bins && weights && mus && sigmas && part_inside && left_tail && out),
                     "sample_pdf_mu_sigma: null pointer");
    return sample_dd_impl(bins, weights, mus, sigmas, part_inside, left_tail, Smooth{1.0f, nullptr}, rand, out, idx_out, N, S,
                          n, pdf_padding, near_cfg, far_cfg, stream);
}

extern "C" DDNERF_EXPORT int ddnerf_sample_pdf_mu_sigma_fused(const float* bins, const float* weights, const float* mus,
                                                const float* sigmas, float smooth, const float* smooth_dev,
                                                const float* rand, float* out, int32_t* idx_out, int64_t N, int S, int n,
                                                int pdf_padding, float near_cfg, float far_cfg, void* stream) {
    DDNERF_CHECK_ARG(N == 0 || (bins && weights && mus && sigmas && out), "sample_pdf_mu_sigma_fused: null pointer");
    DDNERF_CHECK_ARG(smooth_dev || smooth > 0.f, "sample_pdf_mu_sigma_fused: gaussian_smooth_factor=%g must be positive", smooth);
    return sample_dd_impl(bins, weights, mus, sigmas, nullptr, nullptr, Smooth{smooth, smooth_dev}, rand, out, idx_out, N, S, n,
                          pdf_padding, near_cfg, far_cfg, stream);
}

extern "C" DDNERF_EXPORT int ddnerf_find_interval(const float* cdf, const float* u, int32_t* idx_out, int64_t N, int S, int n,
                                    void* stream) {
    DDNERF_CHECK_ARG(N == 0 || n == 0 || (cdf && u && idx_out), "find_interval: null pointer");
    if (N == 0 || n == 0) return 0;
    find_interval_kernel<<<ceil_div(N * n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(cdf, u, idx_out, N, S, n);
    DDNERF_LAUNCHED("find_interval", 1);
    return 0;
}
