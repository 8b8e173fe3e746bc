// K3: samplers -- stratified first-cycle samples and the two coarse-to-fine inverse-CDF
// resamplers, one warp per ray.
//
// Replaces models/samplers.py:30-62 (sample_first_cycle), :64-121 (sample_pdf) and :124-215
// (sample_pdf_with_mu_sigma) of the reference.  The reference materialises an [N,S+1,n] boolean
// mask plus four float temporaries of that shape to locate each sample's interval; here a warp
// stages its ray's weights/bins in shared memory, builds the smoothed CDF with a warp scan and
// binary-searches it per sample: O(S + n log S) work and bins+weights+u in, samples out of HBM.
#include <math.h>

#include "common.cuh"

namespace ddnerf {
namespace {

// ------------------------------------------------------------------------------------------
// sample_first_cycle: elementwise
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float linspace01(int idx, int steps) {     // torch.linspace(0,1,steps)[idx]
    float step = 1.0f / (float)(steps - 1);
    return idx < steps / 2 ? step * (float)idx : 1.0f - step * (float)(steps - idx - 1);
}

__global__ void first_cycle_kernel(const float* __restrict__ near, const float* __restrict__ far, int64_t ray_stride,
                                   const float* __restrict__ t_rand, float* __restrict__ out, int64_t N, int S,
                                   int lindisp) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N * (S + 1)) return;
    int64_t ray = e / (S + 1);
    int i = (int)(e - ray * (S + 1));
    const float nr = __ldg(near + ray * ray_stride), fr = __ldg(far + ray * ray_stride);
    auto tv = [&](int k) {
        float s = linspace01(k, S + 1);
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
                                           int per_warp_floats) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (ray >= N) return;
    float* wpad = smem + (size_t)warp * per_warp_floats;
    float* what = wpad + (S + 2);
    float* cdf = what + S;
    float* sb = cdf + (S + 1);
    float* so = sb + (S + 1);                                            // n_pow2 output slots
    for (int i = lane; i <= S; i += 32) sb[i] = __ldg(bins + ray * (S + 1) + i);
    build_cdf(weights + ray * S, wpad, what, cdf, S, lane, pdf_padding);
    const float* mu_r = mus + ray * S;
    const float* sg_r = sigmas + ray * S;
    const float* pin_r = part_inside + ray * S;
    const float* lt_r = left_tail + ray * S;
    for (int k = lane; k < n_pow2; k += 32) {
        float val = __int_as_float(0x7f800000);                          // +inf padding for the sort
        if (k < n) {
            float u = make_u(us, k, n, rand ? rand + ray * n : nullptr);
            float z, b0, b1;
            int ind;
            if (S == 1) {                                                // samplers.py:185-190
                ind = 0; b0 = sb[0]; b1 = sb[1];
                z = u * __ldg(pin_r) + __ldg(lt_r);
            } else {
                int j = max(count_le(cdf, S + 1, u) - 1, 0);
                int j1 = min(j + 1, S);
                b0 = sb[j]; b1 = sb[j1];
                float c0 = cdf[j], c1 = cdf[j1];
                ind = j;                                                 // torch.max: first index of the maximum
                while (ind > 0 && sb[ind - 1] == sb[ind]) --ind;
                ind = min(ind, S - 1);
                z = ((u - c0) / (c1 - c0)) * __ldg(pin_r + ind) + __ldg(lt_r + ind);
                z = fminf(z, 0.999f);
            }
            z = 1.41421354f * erfinvf(2.0f * z - 1.0f);                  // math_utils.py:202-208
            float t = fminf(fmaxf(z * __ldg(sg_r + ind) + __ldg(mu_r + ind), 0.f), 0.99999f);
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

}  // namespace
}  // namespace ddnerf

using namespace ddnerf;

extern "C" DDNERF_EXPORT int ddnerf_sample_first_cycle(const float* near, const float* far, int64_t ray_stride, const float* t_rand,
                                         float* t_out, int64_t N, int S, int lindisp, void* stream) {
    DDNERF_CHECK_ARG(near && far && t_out, "sample_first_cycle: null pointer");
    DDNERF_CHECK_ARG(S >= 1, "sample_first_cycle: S=%d < 1", S);
    if (N == 0) return 0;
    int64_t total = N * (S + 1);
    first_cycle_kernel<<<ceil_div(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(near, far, ray_stride, t_rand,
                                                                                            t_out, N, S, lindisp);
    DDNERF_LAUNCHED("sample_first_cycle", 1);
    return 0;
}

static int warps_per_block(int per_warp_floats) {
    int w = (int)(48 * 1024 / ((size_t)per_warp_floats * sizeof(float)));
    return w > 8 ? 8 : w;
}

extern "C" DDNERF_EXPORT int ddnerf_sample_pdf(const float* bins, const float* weights, const float* rand, float* out,
                                 int32_t* idx_out, int64_t N, int S, int n, int pdf_padding, void* stream) {
    DDNERF_CHECK_ARG(bins && weights && out, "sample_pdf: null pointer");
    DDNERF_CHECK_ARG(S >= 1 && S <= 2048 && n >= 2, "sample_pdf: S=%d n=%d unsupported", S, n);
    if (N == 0) return 0;
    double s = 1.0 / n;
    USpec us{rand == nullptr, 1.0f, (float)s, (float)((1.0 / s) + 1e-5), 0};
    int per_warp = (S + 2) + S + (S + 1) + (S + 1);
    int wpb = warps_per_block(per_warp);
    DDNERF_CHECK_ARG(wpb >= 1, "sample_pdf: S=%d needs too much shared memory", S);
    sample_pdf_kernel<<<ceil_div(N, wpb), wpb * 32, (size_t)wpb * per_warp * sizeof(float),
                        static_cast<cudaStream_t>(stream)>>>(bins, weights, rand, out, idx_out, N, S, n, pdf_padding, us,
                                                             per_warp);
    DDNERF_LAUNCHED("sample_pdf", 1);
    return 0;
}

extern "C" DDNERF_EXPORT int ddnerf_sample_pdf_mu_sigma(const float* bins, const float* weights, const float* mus, const float* sigmas,
                                          const float* part_inside, const float* left_tail, const float* rand, float* out,
                                          int32_t* idx_out, int64_t N, int S, int n, int pdf_padding, float near_cfg,
                                          float far_cfg, void* stream) {
    DDNERF_CHECK_ARG(bins && weights && mus && sigmas && part_inside && left_tail && out,
                     "sample_pdf_mu_sigma: null pointer");
    DDNERF_CHECK_ARG(S >= 1 && S <= 2048 && n >= 2 && n <= 4096, "sample_pdf_mu_sigma: S=%d n=%d unsupported", S, n);
    if (N == 0) return 0;
    double s = 1.0 / (n - 1);
    USpec us{rand == nullptr, 0.9999f, (float)s, (float)(n + 1e-5), 1};
    int np2 = next_pow2(n);
    int per_warp = (S + 2) + S + (S + 1) + (S + 1) + np2;
    int wpb = warps_per_block(per_warp);
    DDNERF_CHECK_ARG(wpb >= 1, "sample_pdf_mu_sigma: S=%d n=%d needs too much shared memory", S, n);
    sample_pdf_mu_sigma_kernel<<<ceil_div(N, wpb), wpb * 32, (size_t)wpb * per_warp * sizeof(float),
                                 static_cast<cudaStream_t>(stream)>>>(bins, weights, mus, sigmas, part_inside, left_tail,
                                                                      rand, out, idx_out, N, S, n, np2, pdf_padding,
                                                                      near_cfg, far_cfg, us, per_warp);
    DDNERF_LAUNCHED("sample_pdf_mu_sigma", 1);
    return 0;
}

extern "C" DDNERF_EXPORT int ddnerf_find_interval(const float* cdf, const float* u, int32_t* idx_out, int64_t N, int S, int n,
                                    void* stream) {
    DDNERF_CHECK_ARG(cdf && u && idx_out, "find_interval: null pointer");
    if (N == 0 || n == 0) return 0;
    find_interval_kernel<<<ceil_div(N * n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(cdf, u, idx_out, N, S, n);
    DDNERF_LAUNCHED("find_interval", 1);
    return 0;
}
