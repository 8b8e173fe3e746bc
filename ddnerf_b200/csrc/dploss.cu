// K5: depth-distribution loss, forward and analytic backward, one warp per ray.
//
// Replaces models/dd_utils.py:6-78 (estimate_dp_loss) of the reference: KL( fine histogram ||
// coarse piecewise-Gaussian CDF evaluated at the fine edges ), reduction 'mean'.  The reference
// builds an [N,S0+1,S1+1] mask to find each fine edge's coarse cell; here a warp stages its ray's
// coarse pdf / CDF / fence-posts in shared memory, binary-searches per edge and scatters the
// backward contributions with shared-memory atomics.  ~32 B/sample in, 12 B/sample of grads out.
//
// Deviation (documented in DESIGN.md): when the blender row filter (dd_utils.py:12-28) drops rays,
// the reference forgets to filter left_tails_0 and mis-aligns it; this kernel indexes every
// per-ray tensor consistently.  With no dropped rays (always, given the 1e-10 the compositor adds
// to the last weight) the two agree.
#include "common.cuh"

namespace ddnerf {
namespace {

constexpr float EPS = 1e-12f;

struct DpArgs {
    const float *t1, *t0, *w1, *w0, *mus0, *sig0, *lt0, *pin0;
    int blender; int64_t N; int S0, S1;
};

__device__ __forceinline__ float warp_sum32(float v) { return group_sum<32>(v); }

// #{m in [0,len): v[m] < x}
__device__ __forceinline__ int count_lt(const float* v, int len, float x) {
    int lo = 0, hi = len;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (v[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

struct RayState { float Z0, Z1, Zq; bool relevant; };

// Shared layout per warp (floats): p0[S0] cum[S0+1] cdf[S0+1] t0s[S0+1] E[S1+1]  (+ backward arrays)
struct Smem {
    float *p0, *cum, *cdf, *t0s, *E;
    __device__ Smem(float* base, int S0, int S1) {
        p0 = base; cum = p0 + S0; cdf = cum + (S0 + 1); t0s = cdf + (S0 + 1); E = t0s + (S0 + 1);
    }
    static __host__ __device__ int floats(int S0, int S1) { return S0 + 3 * (S0 + 1) + (S1 + 1); }
};

struct Edge { int idx; float x, sr, width, F, eraw; };

__device__ __forceinline__ Edge eval_edge(const DpArgs& a, const Smem& sm, int64_t ray, float t1k) {
    Edge e;
    int j = max(count_lt(sm.t0s, a.S0 + 1, t1k) - 1, 0);      // dd_utils.py:43 mask = t1 > t0 (strict)
    j = min(j, a.S0 - 1);
    while (j > 0 && sm.cdf[j - 1] == sm.cdf[j]) --j;          // torch.max: first index of the maximum
    e.idx = j;
    e.width = sm.t0s[j + 1] - sm.t0s[j];
    float mur = sm.t0s[j] + __ldg(a.mus0 + ray * a.S0 + j) * e.width;
    e.sr = __ldg(a.sig0 + ray * a.S0 + j) * e.width;
    e.x = (t1k - mur) / e.sr;
    e.F = (normal_cdff_(e.x) - __ldg(a.lt0 + ray * a.S0 + j)) / __ldg(a.pin0 + ray * a.S0 + j);
    e.eraw = sm.cdf[j] + e.F * sm.p0[j];
    return e;
}

// Everything both passes need: normalised coarse pdf, CDF, clamped edge estimates E_k, Z's.
__device__ RayState ray_forward(const DpArgs& a, const Smem& sm, int64_t ray, int lane) {
    RayState rs;
    const int S0 = a.S0, S1 = a.S1;
    float part = 0.f;
    for (int i = lane; i < S0; i += 32) { float v = __ldg(a.w0 + ray * S0 + i) + EPS; sm.p0[i] = v; part += v; }
    for (int i = lane; i <= S0; i += 32) sm.t0s[i] = __ldg(a.t0 + ray * (S0 + 1) + i);
    rs.Z0 = warp_sum32(part);
    float p1 = 0.f, raw1 = 0.f;
    for (int k = lane; k < S1; k += 32) { float v = __ldg(a.w1 + ray * S1 + k); raw1 += v; p1 += v + EPS; }
    rs.Z1 = warp_sum32(p1);
    rs.relevant = !a.blender || warp_sum32(raw1) > 1e-10f;     // dd_utils.py:16
    __syncwarp();
    for (int i = lane; i < S0; i += 32) sm.p0[i] = sm.p0[i] / rs.Z0;
    __syncwarp();
    double carry = 0.0;
    for (int base = 0; base < S0 - 1; base += 32) {            // cdf[m] = min(1, sum_{i<m} p0_i), m=1..S0-1
        int i = base + lane;
        float v = i < S0 - 1 ? sm.p0[i] : 0.f;
        double incl_d = group_incl_sum_d<32>((double)v, lane) + carry;
        float incl = (float)incl_d;
        if (i < S0 - 1) { sm.cum[i + 1] = incl; sm.cdf[i + 1] = fminf(1.0f, incl); }
        carry = __shfl_sync(FULL, incl_d, 31);
    }
    if (lane == 0) { sm.cdf[0] = 0.f; sm.cum[0] = 0.f; sm.cdf[S0] = 1.f; sm.cum[S0] = 2.f; }
    __syncwarp();
    for (int k = lane; k <= S1; k += 32) {
        Edge e = eval_edge(a, sm, ray, __ldg(a.t1 + ray * (S1 + 1) + k));
        sm.E[k] = e.eraw > 1.0f ? 1.0f : e.eraw;               // dd_utils.py:66
    }
    __syncwarp();
    float zq = 0.f;
    for (int k = lane; k < S1; k += 32) { float q = sm.E[k + 1] - sm.E[k]; zq += (q < 0.f ? 0.f : q) + EPS; }
    rs.Zq = warp_sum32(zq);
    return rs;
}

__global__ void dp_loss_fwd_kernel(DpArgs a, float* __restrict__ loss_out, float* __restrict__ scratch, int per_warp) {
    extern __shared__ float smem[];
    __shared__ float blk[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (threadIdx.x < 2) blk[threadIdx.x] = 0.f;
    __syncthreads();
    if (ray < a.N) {
        Smem sm(smem + (size_t)warp * per_warp, a.S0, a.S1);
        RayState rs = ray_forward(a, sm, ray, lane);
        float l = 0.f;
        for (int k = lane; k < a.S1; k += 32) {
            float q = sm.E[k + 1] - sm.E[k];
            float qn = ((q < 0.f ? 0.f : q) + EPS) / rs.Zq;
            float p1 = (__ldg(a.w1 + ray * a.S1 + k) + EPS) / rs.Z1;
            l += p1 * (logf(p1) - logf(qn));                   // kl_div pointwise: target*(log target - input)
        }
        l = warp_sum32(l);
        if (lane == 0 && rs.relevant) { atomicAdd(&blk[0], l); atomicAdd(&blk[1], 1.0f); }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(scratch + 0, blk[0]);
        atomicAdd(scratch + 1, blk[1]);
        __threadfence();
        unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(scratch) + 2, 1u);
        if (ticket == gridDim.x - 1) {
            __threadfence();
            float sum = atomicAdd(scratch + 0, 0.f), cnt = atomicAdd(scratch + 1, 0.f);
            *loss_out = cnt > 0.f ? sum / (cnt * (float)a.S1) : 0.f;   // reduction='mean' over kept rays x S1
        }
    }
}

__global__ void dp_loss_bwd_kernel(DpArgs a, const float* __restrict__ g_loss, const float* __restrict__ scratch,
                                   float* __restrict__ g_w0, float* __restrict__ g_mus0, float* __restrict__ g_sig0,
                                   int per_warp) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (ray >= a.N) return;
    const int S0 = a.S0, S1 = a.S1;
    float* base = smem + (size_t)warp * per_warp;
    Smem sm(base, S0, S1);
    float* gcdf = base + Smem::floats(S0, S1);                 // [S0+1]
    float* gp0 = gcdf + (S0 + 1);                              // [S0]
    float* gmu = gp0 + S0;
    float* gsg = gmu + S0;
    RayState rs = ray_forward(a, sm, ray, lane);
    const float cnt = __ldg(scratch + 1);
    if (!rs.relevant || cnt <= 0.f) {
        for (int i = lane; i < S0; i += 32) { g_w0[ray * S0 + i] = 0.f; g_mus0[ray * S0 + i] = 0.f; g_sig0[ray * S0 + i] = 0.f; }
        return;
    }
    const float scale = __ldg(g_loss) / (cnt * (float)S1);
    for (int i = lane; i < S0; i += 32) { gcdf[i] = 0.f; gp0[i] = 0.f; gmu[i] = 0.f; gsg[i] = 0.f; }
    if (lane == 0) gcdf[S0] = 0.f;
    __syncwarp();
    auto gq = [&](int k) -> float {                            // dL/dq_k, zero where the clamp q<0 -> 0 is active
        float q = sm.E[k + 1] - sm.E[k];
        if (q < 0.f) return 0.f;
        float qn = (q + EPS) / rs.Zq;
        float p1 = (__ldg(a.w1 + ray * S1 + k) + EPS) / rs.Z1;
        return scale / rs.Zq * (1.0f - p1 / qn);
    };
    for (int k = lane; k <= S1; k += 32) {
        float t1k = __ldg(a.t1 + ray * (S1 + 1) + k);
        Edge e = eval_edge(a, sm, ray, t1k);
        if (e.eraw > 1.0f) continue;                           // clamp est_cdf > 1 -> 1 blocks the gradient
        float gE = (k >= 1 ? gq(k - 1) : 0.f) - (k <= S1 - 1 ? gq(k) : 0.f);
        if (gE == 0.f) continue;
        const int j = e.idx;
        atomicAdd(gcdf + j, gE);
        atomicAdd(gp0 + j, gE * e.F);
        float pin = __ldg(a.pin0 + ray * S0 + j);
        float gx = gE * sm.p0[j] / pin * (0.3989422804f * expf(-0.5f * e.x * e.x));
        atomicAdd(gmu + j, -gx / e.sr * e.width);
        atomicAdd(gsg + j, -gx * e.x / e.sr * e.width);
    }
    __syncwarp();
    // cdf[m] = min(1, cum[m]): route g_cdf[m] to p0[0..m-1] for m = 1..S0-1 (suffix sum), with
    // torch.minimum's tie rule (half the gradient when cum == 1)
    float carry = 0.f;
    const int nchunks = (S0 + 31) / 32;
    for (int c = nchunks - 1; c >= 0; --c) {
        int i = c * 32 + lane;                                 // p0 index; receives sum_{m=i+1}^{S0-1} h_m
        float h = 0.f;
        int m = i + 1;
        if (m <= S0 - 1) { float cu = sm.cum[m]; h = gcdf[m] * (cu < 1.0f ? 1.0f : (cu == 1.0f ? 0.5f : 0.f)); }
        float suf = group_suffix_sum<32>(h, lane) + carry;
        if (i < S0) gp0[i] += suf;
        carry = __shfl_sync(FULL, suf, 0);
    }
    __syncwarp();
    float dot = 0.f;
    for (int i = lane; i < S0; i += 32) dot += gp0[i] * sm.p0[i];
    dot = warp_sum32(dot);
    for (int i = lane; i < S0; i += 32) {
        g_w0[ray * S0 + i] = (gp0[i] - dot) / rs.Z0;           // p0 = (w0+eps)/sum(w0+eps)
        g_mus0[ray * S0 + i] = gmu[i];
        g_sig0[ray * S0 + i] = gsg[i];
    }
}

int warps_for(size_t per_warp_bytes) {
    int w = (int)(48 * 1024 / per_warp_bytes);
    return w > 8 ? 8 : w;
}

}  // namespace
}  // namespace ddnerf

using namespace ddnerf;

extern "C" DDNERF_EXPORT int ddnerf_dp_loss_forward(const float* t1, const float* t0, const float* w1, const float* w0,
                                      const float* mus0, const float* sigmas0, const float* lt0, const float* pin0,
                                      int blender, float* loss_out, float* scratch, int64_t N, int S0, int S1,
                                      void* stream) {
    DDNERF_CHECK_ARG(t1 && t0 && w1 && w0 && mus0 && sigmas0 && lt0 && pin0 && loss_out && scratch,
                     "dp_loss_forward: null pointer");
    DDNERF_CHECK_ARG(S0 >= 1 && S1 >= 1 && S0 <= 1024 && S1 <= 1024, "dp_loss_forward: S0=%d S1=%d unsupported", S0, S1);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(scratch, 0, 4 * sizeof(float), st);
    if (N == 0) { cudaMemsetAsync(loss_out, 0, sizeof(float), st); return 0; }
    DpArgs a{t1, t0, w1, w0, mus0, sigmas0, lt0, pin0, blender, N, S0, S1};
    int per_warp = Smem::floats(S0, S1);
    int wpb = warps_for(per_warp * sizeof(float));
    DDNERF_CHECK_ARG(wpb >= 1, "dp_loss_forward: shapes need too much shared memory");
    dp_loss_fwd_kernel<<<ceil_div(N, wpb), wpb * 32, (size_t)wpb * per_warp * sizeof(float), st>>>(a, loss_out, scratch,
                                                                                                   per_warp);
    DDNERF_LAUNCHED("dp_loss_forward", 1);
    return 0;
}

extern "C" DDNERF_EXPORT int ddnerf_dp_loss_backward(const float* t1, const float* t0, const float* w1, const float* w0,
                                       const float* mus0, const float* sigmas0, const float* lt0, const float* pin0,
                                       int blender, const float* g_loss, const float* scratch, float* g_w0,
                                       float* g_mus0, float* g_sigmas0, int64_t N, int S0, int S1, void* stream) {
    DDNERF_CHECK_ARG(t1 && t0 && w1 && w0 && mus0 && sigmas0 && lt0 && pin0 && g_loss && scratch && g_w0 && g_mus0 &&
                         g_sigmas0, "dp_loss_backward: null pointer");
    DDNERF_CHECK_ARG(S0 >= 1 && S1 >= 1 && S0 <= 1024 && S1 <= 1024, "dp_loss_backward: S0=%d S1=%d unsupported", S0, S1);
    if (N == 0) return 0;
    DpArgs a{t1, t0, w1, w0, mus0, sigmas0, lt0, pin0, blender, N, S0, S1};
    int per_warp = Smem::floats(S0, S1) + (S0 + 1) + 3 * S0;
    int wpb = warps_for(per_warp * sizeof(float));
    DDNERF_CHECK_ARG(wpb >= 1, "dp_loss_backward: shapes need too much shared memory");
    dp_loss_bwd_kernel<<<ceil_div(N, wpb), wpb * 32, (size_t)wpb * per_warp * sizeof(float),
                         static_cast<cudaStream_t>(stream)>>>(a, g_loss, scratch, g_w0, g_mus0, g_sigmas0, per_warp);
    DDNERF_LAUNCHED("dp_loss_backward", 1);
    return 0;
}
