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
#include <type_traits>

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

// ------------------------------------------------------------------------------------------
// Fast path (S0 <= 256): G lanes per ray, C consecutive coarse cells per lane (G*C >= S0, a power of two).
// Same scheme as the fast resamplers (sampler.cu): blocked double-precision CDF with one lane-group scan,
// {cdf, t0} pairs in shared memory, branch-free binary search per fine edge, reciprocal-based divisions,
// one lg2 per KL term.  The backward keeps every edge's (cell, F, x) from its forward phase in shared
// memory instead of searching and evaluating erf twice, and walks the edges in per-lane consecutive runs
// so that the scatter into a coarse cell is accumulated in registers and flushed once per (lane, cell) --
// shared-memory float atomics are CAS loops on this architecture, and peaked fine histograms put whole
// runs of edges into one cell.  Round-1 profile of the generic kernels above: profiles/r01_ncu_dp_loss_*.md.
// ------------------------------------------------------------------------------------------
constexpr float INF_F = __builtin_huge_valf();
// division that keeps IEEE semantics for a zero divisor (degenerate cells), fast otherwise
__device__ __forceinline__ float div_guard(float a, float b) { return b != 0.f ? div_fast(a, b) : a / b; }

// K: fine-side chunks per lane (S1 + 1 <= K*G); all of a ray's global loads are requested up front.
template <int G, int C, int K, bool BWD>
struct FastDp {
    static constexpr int P = G * C;
    // floats per ray
    static __host__ __device__ int floats(int S1) {
        int e = S1 + 2;                                   // E[S1+1] (+1 pad)
        int f = 2 * (P + 1) + P + e;                      // cb, p0, E
        if (BWD) f += (P + 1) + (P + 1) + 3 * P + 3 * e;  // cum, gcdf, gp0/gmu/gsg, edge j/F/x
        return (f + 1) & ~1;                              // keep float2 alignment
    }
    float2* cb; float *p0, *E, *cum, *gcdf, *gp0, *gmu, *gsg, *eF, *ex; int* ej;
    __device__ FastDp(float* base, int S1) {
        const int e = S1 + 2;
        cb = reinterpret_cast<float2*>(base); p0 = base + 2 * (P + 1); E = p0 + P;
        if (BWD) {
            cum = E + e; gcdf = cum + (P + 1); gp0 = gcdf + (P + 1); gmu = gp0 + P; gsg = gmu + P;
            ej = reinterpret_cast<int*>(gsg + P); eF = gsg + P + e; ex = eF + e;
        }
    }

    // normalised coarse pdf, CDF, clamped edge estimates E_k, the three normalisers
    __device__ __forceinline__ RayState forward(const DpArgs& a, int64_t ray, int gl, float (&w1v)[K]) {
        RayState rs;
        const int S0 = a.S0, S1 = a.S1;
        const float* w0r = a.w0 + ray * S0;
        const float* t1r = a.t1 + ray * (S1 + 1);
        const float* w1r = a.w1 + ray * S1;
        const int e0 = gl * C;
        float t1v[K];
#pragma unroll
        for (int c = 0; c < K; ++c) {
            const int k = c * G + gl;
            t1v[c] = k <= S1 ? __ldg(t1r + k) : 0.f;
            w1v[c] = k < S1 ? __ldg(w1r + k) : 0.f;
        }
        float v[C], part = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float w = __ldg(w0r + min(e0 + c, S0 - 1));
            v[c] = e0 + c < S0 ? w + EPS : 0.f;
            part += v[c];
        }
        const float* t0r = a.t0 + ray * (S0 + 1);
        for (int q = gl; q <= P; q += G) {
            cb[q].y = q <= S0 ? __ldg(t0r + q) : INF_F;
            if (q > S0) cb[q].x = 2.f;
        }
        float p1 = 0.f, raw1 = 0.f;
#pragma unroll
        for (int c = 0; c < K; ++c) if (c * G + gl < S1) { raw1 += w1v[c]; p1 += w1v[c] + EPS; }
        rs.Z0 = group_sum<G>(part);
        rs.Z1 = group_sum<G>(p1);
        rs.relevant = !a.blender || group_sum<G>(raw1) > 1e-10f;      // dd_utils.py:16
        double incl[C], run = 0.0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float pc = div_fast(v[c], rs.Z0);
            if (e0 + c < S0) p0[e0 + c] = pc;
            run += (double)pc; incl[c] = run;
        }
        double lanes_incl = group_incl_sum_d<G>(run, gl);
        double off = __shfl_up_sync(FULL, lanes_incl, 1, G);
        if (gl == 0) off = 0.0;
#pragma unroll
        for (int c = 0; c < C; ++c) {                     // cdf[m] = min(1, sum_{i<m} p0_i), m = 1..S0-1
            const int m = e0 + c + 1;
            if (m <= S0 - 1) {
                float cu = (float)(off + incl[c]);
                cb[m].x = fminf(1.0f, cu);
                if (BWD) cum[m] = cu;
            }
        }
        if (gl == 0) { cb[0].x = 0.f; cb[S0].x = 1.f; if (BWD) { cum[0] = 0.f; cum[S0] = 2.f; } }
        __syncwarp();
        const float* mur = a.mus0 + ray * S0; const float* sgr = a.sig0 + ray * S0;
        const float* ltr = a.lt0 + ray * S0; const float* pir = a.pin0 + ray * S0;
#pragma unroll
        for (int c = 0; c < K; ++c) {
            const int k = c * G + gl;
            if (k > S1) break;
            const float t = t1v[c];
            int pos = 0;                                  // #{q in 1..P-1 : t0[q] < t}  (dd_utils.py:43, strict)
#pragma unroll
            for (int step = P / 2; step >= 1; step /= 2)
                if (cb[pos + step].y < t) pos += step;
            int j = min(pos, S0 - 1);
            while (j > 0 && cb[j - 1].x == cb[j].x) --j;  // torch.max: first index of the maximum
            const float2 a0 = cb[j];
            const float width = cb[j + 1].y - a0.y;
            const float mr = a0.y + __ldg(mur + j) * width;
            const float sr = __ldg(sgr + j) * width;
            const float pin = __ldg(pir + j);
            const float x = div_guard(t - mr, sr);
            const float num = normal_cdff_(x) - __ldg(ltr + j);
            const float F = div_guard(num, pin);
            const float eraw = a0.x + F * p0[j];
            E[k] = eraw > 1.0f ? 1.0f : eraw;             // dd_utils.py:66
            if (BWD) { ej[k] = eraw > 1.0f ? -1 - j : j; eF[k] = F; ex[k] = x; }
        }
        __syncwarp();
        float zq = 0.f;
        for (int k = gl; k < S1; k += G) { float q = E[k + 1] - E[k]; zq += (q < 0.f ? 0.f : q) + EPS; }
        rs.Zq = group_sum<G>(zq);
        return rs;
    }
};

template <int G, int C, int K>
__global__ void __launch_bounds__(256) dp_loss_fwd_fast_kernel(DpArgs a, float* __restrict__ loss_out,
                                                                float* __restrict__ scratch, int per_ray) {
    extern __shared__ __align__(8) float smem[];
    __shared__ float blk[2];
    const int grp = threadIdx.x / G, gl = threadIdx.x % G;
    int64_t ray = (int64_t)blockIdx.x * (blockDim.x / G) + grp;
    const bool valid = ray < a.N;
    if (!valid) ray = a.N - 1;
    if (threadIdx.x < 2) blk[threadIdx.x] = 0.f;
    __syncthreads();
    FastDp<G, C, K, false> sm(smem + (size_t)grp * per_ray, a.S1);
    float w1v[K];
    RayState rs = sm.forward(a, ray, gl, w1v);
    // sum_k p1 (log p1 - log qn) = (1/Z1) sum_k pe ln(pe/qe) + ln(Zq/Z1),  pe = w1+eps, qe = max(q,0)+eps
    float l = 0.f;
#pragma unroll
    for (int c = 0; c < K; ++c) {
        const int k = c * G + gl;
        if (k >= a.S1) break;
        float q = sm.E[k + 1] - sm.E[k];
        float qe = (q < 0.f ? 0.f : q) + EPS;
        float pe = w1v[c] + EPS;
        l += pe * lg2_(pe * rcp_(qe));
    }
    l = group_sum<G>(l);
    if (gl == 0 && valid && rs.relevant) {
        float kl = l * LN2 / rs.Z1 + logf(rs.Zq / rs.Z1);
        atomicAdd(&blk[0], kl); atomicAdd(&blk[1], 1.0f);
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

template <int G, int C, int K>
__global__ void __launch_bounds__(128) dp_loss_bwd_fast_kernel(DpArgs a, const float* __restrict__ g_loss,
                                                                const float* __restrict__ scratch,
                                                                float* __restrict__ g_w0, float* __restrict__ g_mus0,
                                                                float* __restrict__ g_sig0, int per_ray) {
    extern __shared__ __align__(8) float smem[];
    constexpr int P = G * C;
    const int grp = threadIdx.x / G, gl = threadIdx.x % G;
    int64_t ray = (int64_t)blockIdx.x * (blockDim.x / G) + grp;
    const bool valid = ray < a.N;
    if (!valid) ray = a.N - 1;
    const int S0 = a.S0, S1 = a.S1;
    FastDp<G, C, K, true> sm(smem + (size_t)grp * per_ray, S1);
    float w1v[K];
    RayState rs = sm.forward(a, ray, gl, w1v);
    const float cnt = __ldg(scratch + 1);
    const bool live = rs.relevant && cnt > 0.f;           // uniform over the lane group
    const float scale = live ? __ldg(g_loss) / (cnt * (float)S1) : 0.f;
    for (int i = gl; i <= P; i += G) { sm.gcdf[i] = 0.f; if (i < P) { sm.gp0[i] = 0.f; sm.gmu[i] = 0.f; sm.gsg[i] = 0.f; } }
    __syncwarp();
    const float zr = rs.Zq / rs.Z1, sz = scale / rs.Zq;
    const float* w1r = a.w1 + ray * S1;
    auto gq = [&](int k) -> float {                       // dL/dq_k, zero where the clamp q<0 -> 0 is active
        float q = sm.E[k + 1] - sm.E[k];
        if (q < 0.f) return 0.f;
        float pe = __ldg(w1r + k) + EPS;
        return sz * (1.0f - pe * rcp_(q + EPS) * zr);     // scale/Zq (1 - p1/qn)
    };
    // every lane walks KC consecutive edges; contributions to one coarse cell accumulate in registers
    const int KC = (S1 + 1 + G - 1) / G;
    const float* pir = a.pin0 + ray * S0;
    int cur = -1;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    auto flush = [&]() {
        if (cur >= 0) {
            if (acc0 != 0.f) atomicAdd(sm.gcdf + cur, acc0);
            if (acc1 != 0.f) atomicAdd(sm.gp0 + cur, acc1);
            if (acc2 != 0.f) atomicAdd(sm.gmu + cur, acc2);
            if (acc3 != 0.f) atomicAdd(sm.gsg + cur, acc3);
        }
    };
    if (live) {
        for (int c = 0; c < KC; ++c) {
            const int k = gl * KC + c;
            if (k > S1) break;
            const int jj = sm.ej[k];
            if (jj < 0) continue;                         // clamp est_cdf > 1 -> 1 blocks the gradient
            const float gE = (k >= 1 ? gq(k - 1) : 0.f) - (k <= S1 - 1 ? gq(k) : 0.f);
            if (gE == 0.f) continue;
            if (jj != cur) { flush(); cur = jj; acc0 = acc1 = acc2 = acc3 = 0.f; }
            const float F = sm.eF[k], x = sm.ex[k];
            const float2 a0 = sm.cb[jj];
            const float width = sm.cb[jj + 1].y - a0.y;
            const float sr = __ldg(a.sig0 + ray * S0 + jj) * width;
            const float pin = __ldg(pir + jj);
            const float gx = div_guard(gE * sm.p0[jj], pin) * (0.3989422804f * ex2_(-0.5f * x * x * L2E));
            const float gxs = div_guard(gx, sr) * width;
            acc0 += gE; acc1 += gE * F; acc2 -= gxs; acc3 -= gxs * x;
        }
        flush();
    }
    __syncwarp();
    // cdf[m] = min(1, cum[m]): route g_cdf[m] to p0[0..m-1] for m = 1..S0-1 (suffix sum), with
    // torch.minimum's tie rule (half the gradient when cum == 1).  Lane owns cells e0..e0+C-1.
    const int e0 = gl * C;
    float h[C], tail = 0.f;
#pragma unroll
    for (int c = C - 1; c >= 0; --c) {                    // h[c] = sum over m in (e0+c, e0+C] of routed g_cdf[m]
        const int m = e0 + c + 1;
        float r = 0.f;
        if (m <= S0 - 1) { float cu = sm.cum[m]; r = sm.gcdf[m] * (cu < 1.0f ? 1.0f : (cu == 1.0f ? 0.5f : 0.f)); }
        tail += r;
        h[c] = tail;
    }
    const float suf_incl = group_suffix_sum<G>(tail, gl);
    const float after = suf_incl - tail;                  // routed g_cdf of all later lanes
    float gp[C], dot = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int i = e0 + c;
        gp[c] = i < S0 ? sm.gp0[i] + h[c] + after : 0.f;
        dot += i < S0 ? gp[c] * sm.p0[i] : 0.f;
    }
    dot = group_sum<G>(dot);
    if (valid) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int i = e0 + c;
            if (i < S0) {
                g_w0[ray * S0 + i] = live ? (gp[c] - dot) / rs.Z0 : 0.f;   // p0 = (w0+eps)/sum(w0+eps)
                g_mus0[ray * S0 + i] = sm.gmu[i];
                g_sig0[ray * S0 + i] = sm.gsg[i];
            }
        }
    }
}

template <typename F>
bool dispatch_fast(int S0, int S1, F&& f) {
#define DDNERF_FAST(G, C)                                                                                                  \
    if (S0 <= G * C) {                                                                                                     \
        if (S1 + 1 <= 5 * G) return f(std::integral_constant<int, G>{}, std::integral_constant<int, C>{}, std::integral_constant<int, 5>{}); \
        if (S1 + 1 <= 9 * G) return f(std::integral_constant<int, G>{}, std::integral_constant<int, C>{}, std::integral_constant<int, 9>{}); \
        return false;                                                                                                      \
    }
    // eight cells per lane wherever S allows: short rays share a warp
    DDNERF_FAST(4, 1) DDNERF_FAST(4, 2) DDNERF_FAST(4, 4) DDNERF_FAST(4, 8) DDNERF_FAST(8, 8) DDNERF_FAST(16, 8) DDNERF_FAST(32, 8)
#undef DDNERF_FAST
    return false;
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
    bool fast = dispatch_fast(S0, S1, [&](auto g, auto c, auto kk) {
        constexpr int G = decltype(g)::value, C = decltype(c)::value, K = decltype(kk)::value;
        const int per_ray = FastDp<G, C, K, false>::floats(S1);
        const size_t bytes = (size_t)(256 / G) * per_ray * sizeof(float);
        if (bytes > 200 * 1024) return false;
        auto kern = dp_loss_fwd_fast_kernel<G, C, K>;
        if (bytes > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        kern<<<ceil_div(N, 256 / G), 256, bytes, st>>>(a, loss_out, scratch, per_ray);
        return true;
    });
    if (!fast) {
        int per_warp = Smem::floats(S0, S1);
        int wpb = warps_for(per_warp * sizeof(float));
        DDNERF_CHECK_ARG(wpb >= 1, "dp_loss_forward: shapes need too much shared memory");
        dp_loss_fwd_kernel<<<ceil_div(N, wpb), wpb * 32, (size_t)wpb * per_warp * sizeof(float), st>>>(a, loss_out, scratch,
                                                                                                       per_warp);
    }
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
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    bool fast = dispatch_fast(S0, S1, [&](auto g, auto c, auto kk) {
        constexpr int G = decltype(g)::value, C = decltype(c)::value, K = decltype(kk)::value;
        const int per_ray = FastDp<G, C, K, true>::floats(S1);
        const size_t bytes = (size_t)(128 / G) * per_ray * sizeof(float);
        if (bytes > 200 * 1024) return false;
        auto kern = dp_loss_bwd_fast_kernel<G, C, K>;
        if (bytes > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        kern<<<ceil_div(N, 128 / G), 128, bytes, st>>>(a, g_loss, scratch, g_w0, g_mus0, g_sigmas0, per_ray);
        return true;
    });
    if (!fast) {
        int per_warp = Smem::floats(S0, S1) + (S0 + 1) + 3 * S0;
        int wpb = warps_for(per_warp * sizeof(float));
        DDNERF_CHECK_ARG(wpb >= 1, "dp_loss_backward: shapes need too much shared memory");
        dp_loss_bwd_kernel<<<ceil_div(N, wpb), wpb * 32, (size_t)wpb * per_warp * sizeof(float), st>>>(
            a, g_loss, scratch, g_w0, g_mus0, g_sigmas0, per_warp);
    }
    DDNERF_LAUNCHED("dp_loss_backward", 1);
    return 0;
}
