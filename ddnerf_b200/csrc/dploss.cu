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
#include <stdlib.h>

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

// left_tail / part_inside of coarse cell (models.py:254-258, detached constants of the loss): read from the caller's
// tensors, or -- fused DDNeRF coarse path, both pointers NULL -- evaluated per cell with the reference's fp32 formula.
__device__ __forceinline__ void cell_tails(const DpArgs& a, int64_t off, float mu, float sigma, float& lt, float& pin) {
    if (a.lt0) { lt = __ldg(a.lt0 + off); pin = __ldg(a.pin0 + off); return; }
    lt = normal_cdff_((0.0f - mu) / sigma);
    pin = normal_cdff_((1.0f - mu) / sigma) - lt;
}

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

// The loss as DDNerfModel.predict consumes it (models.py:287-289): kl * scale + mus_reg + sig_reg with scale = the number
// of fine cells and the two regularisers in regs[2], regs[3] (what ddnerf_composite_dd_forward wrote).  regs == NULL and
// scale == 1: the plain kl_div of dd_utils.py.  Folding it here removes three elementwise launches from the forward and
// the slice / mul / add chain (eight launches) from the backward of a training step.
struct DpTotal { float scale; const float* regs; float* g_regs; };
__device__ __forceinline__ float dp_total(const DpTotal& tt, float kl) {
    float v = kl * tt.scale;
    if (tt.regs) { v = v + __ldg(tt.regs + 2); v = v + __ldg(tt.regs + 3); }
    return v;
}
// cotangent of regs = {mus_loss, sig_loss, mus_reg, sig_reg}: written by one thread of the backward launch
__device__ __forceinline__ void dp_total_g_regs(const DpTotal& tt, const float* g_loss) {
    if (tt.g_regs && blockIdx.x == 0 && threadIdx.x == 0) {
        const float g = __ldg(g_loss);
        tt.g_regs[0] = 0.f; tt.g_regs[1] = 0.f; tt.g_regs[2] = g; tt.g_regs[3] = g;
    }
}

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
    const float mu_j = __ldg(a.mus0 + ray * a.S0 + j), sg_j = __ldg(a.sig0 + ray * a.S0 + j);
    float mur = sm.t0s[j] + mu_j * e.width;
    e.sr = sg_j * e.width;
    e.x = (t1k - mur) / e.sr;
    float lt_j, pin_j;
    cell_tails(a, ray * a.S0 + j, mu_j, sg_j, lt_j, pin_j);
    e.F = (normal_cdff_(e.x) - lt_j) / pin_j;
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

__global__ void dp_loss_fwd_kernel(DpArgs a, DpTotal tt, float* __restrict__ loss_out, float* __restrict__ scratch, int per_warp) {
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
            *loss_out = dp_total(tt, cnt > 0.f ? sum / (cnt * (float)a.S1) : 0.f);   // reduction='mean' over kept rays x S1
        }
    }
}

__global__ void dp_loss_bwd_kernel(DpArgs a, DpTotal tt, const float* __restrict__ g_loss, const float* __restrict__ scratch,
                                   float* __restrict__ g_w0, float* __restrict__ g_mus0, float* __restrict__ g_sig0,
                                   int per_warp) {
    extern __shared__ float smem[];
    dp_total_g_regs(tt, g_loss);
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
    const float scale = __ldg(g_loss) * tt.scale / (cnt * (float)S1);
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
        float lt_unused, pin;
        cell_tails(a, ray * S0 + j, __ldg(a.mus0 + ray * S0 + j), __ldg(a.sig0 + ray * S0 + j), lt_unused, pin);
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
// Fast path (S0 <= 256): G lanes per ray, C coarse cells per lane (G*C >= S0, a power of two), K fine-side
// chunks per lane (S1 + 1 <= K*G).
//   * every global load of the ray is requested before the first dependent instruction;
//   * each coarse cell becomes one 32-byte shared-memory record {t0, cdf, a, b | lt, 1/pin, p0, 1/sigma} with
//     a = t0 + mu*width and b = 1/(sigma*width), so a fine edge costs one search, two 16-byte shared loads,
//     x = (t-a)*b, one erf and two FMAs -- the divisions are per cell, not per edge, and nothing is gathered
//     from global memory inside the edge loop (a gather there puts a DRAM latency into every chunk);
//   * the coarse CDF is a per-lane serial sum plus ONE lane-group scan in double (monotone, ties exact);
//   * the K interval searches of a lane run in lock step (common.cuh: SmemSearch);
//   * one lg2 per KL term:  sum_k p1 (log p1 - log qn) = (1/Z1) sum_k pe ln(pe/qe) + ln(Zq/Z1);
//   * backward: every edge's (cell, F, x) stays in registers from the forward phase; the scatter into the
//     coarse cells is a segmented lane-group scan over the (sorted) cells of a chunk's edges, so only the last
//     lane of each run touches shared memory -- no atomics (float atomics on shared memory are CAS loops).
// Round-1 profile of the generic kernels above: profiles/r01_ncu_dp_loss_*.md.
// ------------------------------------------------------------------------------------------
constexpr float INF_F = __builtin_huge_valf();

// 1/x with one Newton step on the SFU reciprocal (correctly rounded except in rare cases); IEEE for x = 0
__device__ __forceinline__ float rcp_nr(float x) {
    if (x == 0.f) return 1.0f / x;
    float r = rcp_(x);
    return fmaf(fmaf(-x, r, 1.0f), r, r);
}

// Standard normal CDF through erfc(z) = t exp(-z^2 + P(t)), t = 1/(1 + z/2) (Chebyshev fit, fractional error
// < 1.2e-7 for every z >= 0, Numerical Recipes 6.2): 10 FMAs and two SFU operations instead of the ~45
// instructions of erff, and -- unlike 0.5(1 + erf) -- it keeps its RELATIVE accuracy in the lower tail.
__device__ __forceinline__ float normal_cdf_fast(float x) {
    const float z = fabsf(x) * 0.70710678f;
    const float t = rcp_(fmaf(0.5f, z, 1.0f));
    float p = fmaf(t, 0.17087277f, -0.82215223f);
    p = fmaf(t, p, 1.48851587f); p = fmaf(t, p, -1.13520398f); p = fmaf(t, p, 0.27886807f);
    p = fmaf(t, p, -0.18628806f); p = fmaf(t, p, 0.09678418f); p = fmaf(t, p, 0.37409196f);
    p = fmaf(t, p, 1.00002368f); p = fmaf(t, p, -1.26551223f);
    const float half_erfc = 0.5f * t * ex2_(fmaf(-z, z, p) * L2E);
    return x <= 0.f ? half_erfc : 1.0f - half_erfc;      // NaN falls through to 1 - NaN = NaN
}

__device__ __forceinline__ int skew(int i) { return i + (i >> 5); }     // stride-C writes of the blocked lanes: no bank conflicts

template <int G, int C, int K, bool BWD>
struct CellDp {
    static constexpr int P = G * C;
    static constexpr int NSK = ((P + (P >> 5) + 2) + 3) & ~3;            // skewed array length (multiple of 4)
    static __host__ __device__ constexpr int e_floats(int S1) { return (S1 + 2 + 3) & ~3; }
    static __host__ __device__ constexpr int floats(int S1) {
        return 4 * (P + 1) + P + 2 * NSK + e_floats(S1) + (BWD ? P + 4 * P : 0);
    }
    float4* rec;     // [P+1] {t0, a, b, lt}: search key first; a = t0 + mu*width, b = 1/(sigma*width)
    float* ipin;     // [P]   1/part_inside
    float* cdf;      // [NSK] skewed: coarse CDF at the cell's left edge
    float* p0;       // [NSK] skewed: normalised coarse pdf
    float* E;        // [S1+2] clamped CDF estimates at the fine edges
    float* isig;     // [P]   1/sigma                              (backward only)
    float* acc;      // [4][P] sum gE, gE F, gE phi, gE phi x     (backward only)
    __device__ CellDp(float* base, int S1) {
        rec = reinterpret_cast<float4*>(base);
        ipin = base + 4 * (P + 1);
        cdf = ipin + P; p0 = cdf + NSK; E = p0 + NSK;
        isig = E + e_floats(S1); acc = isig + P;
    }
    // per-lane state of the fine side
    struct Fine { float w1[K]; int j[K]; float F[K], x[K]; };   // j < 0: estimate clamped to 1 (no gradient)
    float tie[C];                                                // d min(1,cum_m)/d cum_m for the lane's m = e0+c+1

    __device__ __forceinline__ RayState forward(const DpArgs& a, int64_t ray, int gl, Fine& fs) {
        RayState rs;
        const int S0 = a.S0, S1 = a.S1;
        const float* t1r = a.t1 + ray * (S1 + 1);
        const float* w1r = a.w1 + ray * S1;
        const float* t0r = a.t0 + ray * (S0 + 1);
        const int64_t r0 = ray * S0;
        const int e0 = gl * C;
        // ---- all global loads ----
        float t1v[K];
#pragma unroll
        for (int c = 0; c < K; ++c) {
            const int k = c * G + gl;
            t1v[c] = k <= S1 ? __ldg(t1r + k) : 0.f;
            fs.w1[c] = k < S1 ? __ldg(w1r + k) : 0.f;
        }
        float w0v[C];
#pragma unroll
        for (int c = 0; c < C; ++c) w0v[c] = __ldg(a.w0 + r0 + min(e0 + c, S0 - 1));       // blocked: lane owns e0..e0+C-1
        float t0v[C + 1], t0n[C], muv[C], sgv[C], ltv[C], piv[C];                            // interleaved: q = gl + c*G
#pragma unroll
        for (int c = 0; c <= C; ++c) { const int q = gl + c * G; t0v[c] = q <= S0 ? __ldg(t0r + q) : INF_F; }
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int q = min(gl + c * G, S0 - 1);
            t0n[c] = __ldg(t0r + q + 1);
            muv[c] = __ldg(a.mus0 + r0 + q); sgv[c] = __ldg(a.sig0 + r0 + q);
            if (a.lt0) { ltv[c] = __ldg(a.lt0 + r0 + q); piv[c] = __ldg(a.pin0 + r0 + q); }
        }
        if (!a.lt0) {
#pragma unroll
            for (int c = 0; c < C; ++c) cell_tails(a, 0, muv[c], sgv[c], ltv[c], piv[c]);
        }
        // ---- fine-side normalisers ----
        float p1 = 0.f, raw1 = 0.f;
#pragma unroll
        for (int c = 0; c < K; ++c) if (c * G + gl < S1) { raw1 += fs.w1[c]; p1 += fs.w1[c] + EPS; }
        // ---- coarse pdf and CDF (blocked) ----
        float v[C], part = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) { v[c] = e0 + c < S0 ? w0v[c] + EPS : 0.f; part += v[c]; }
        rs.Z0 = group_sum<G>(part);
        rs.Z1 = group_sum<G>(p1);
        rs.relevant = !a.blender || group_sum<G>(raw1) > 1e-10f;      // dd_utils.py:16
        double incl[C], run = 0.0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float pc = div_fast(v[c], rs.Z0);
            p0[skew(e0 + c)] = pc;
            run += (double)pc; incl[c] = run;
        }
        const double lanes_incl = group_incl_sum_d<G>(run, gl);
        double off = __shfl_up_sync(FULL, lanes_incl, 1, G);
        if (gl == 0) off = 0.0;
#pragma unroll
        for (int c = 0; c < C; ++c) {                     // cdf[m] = min(1, sum_{i<m} p0_i), m = 1..S0-1
            const int m = e0 + c + 1;
            tie[c] = 0.f;
            if (m <= S0 - 1) {
                const float cu = (float)(off + incl[c]);
                cdf[skew(m)] = fminf(1.0f, cu);
                tie[c] = cu < 1.0f ? 1.0f : (cu == 1.0f ? 0.5f : 0.f);   // torch.minimum's tie rule
            }
        }
        if (gl == 0) cdf[0] = 0.f;
        // ---- per-cell constants (interleaved) ----
#pragma unroll
        for (int c = 0; c <= C; ++c) {
            const int q = gl + c * G;
            if (q <= P) {
                float4 r = make_float4(t0v[c], 0.f, 0.f, 0.f);
                if (c < C && q < S0) {
                    const float width = t0n[c] - t0v[c];
                    r.y = t0v[c] + muv[c] * width;                    // a: the Gaussian's mean in ray space
                    r.z = rcp_nr(sgv[c] * width);                     // b: 1 / its standard deviation
                    r.w = ltv[c];
                    ipin[q] = rcp_nr(piv[c]);
                    if (BWD) isig[q] = rcp_nr(sgv[c]);
                }
                rec[q] = r;
            }
        }
        __syncwarp();
        // ---- fine edges: coarse cell by binary search (dd_utils.py:43, strict), estimated CDF ----
        const unsigned sbase = (unsigned)__cvta_generic_to_shared(rec);
        unsigned at[K];
#pragma unroll
        for (int c = 0; c < K; ++c) at[c] = sbase;
        SmemSearch<P / 2, true, K, 16>::run(at, t1v);
#pragma unroll
        for (int c = 0; c < K; ++c) {
            const int k = c * G + gl;
            if (k > S1) { fs.j[c] = -1; fs.F[c] = 0.f; fs.x[c] = 0.f; continue; }
            int j = min((int)((at[c] - sbase) >> 4), S0 - 1);
            const float cj = cdf[skew(j)];
            while (j > 0 && cdf[skew(j - 1)] == cj) --j;   // torch.max: first index of the maximum
            const float4 r = rec[j];
            const float x = (t1v[c] - r.y) * r.z;
            const float F = (normal_cdf_fast(x) - r.w) * ipin[j];
            const float eraw = cj + F * p0[skew(j)];
            E[k] = eraw > 1.0f ? 1.0f : eraw;             // dd_utils.py:66
            fs.j[c] = eraw > 1.0f ? ~j : j; fs.F[c] = F; fs.x[c] = x;
        }
        __syncwarp();
        float zq = 0.f;
#pragma unroll
        for (int c = 0; c < K; ++c) {
            const int k = c * G + gl;
            if (k < S1) { float q = E[k + 1] - E[k]; zq += (q < 0.f ? 0.f : q) + EPS; }
        }
        rs.Zq = group_sum<G>(zq);
        return rs;
    }
};

// Forward: one KL value and one relevance flag per ray into scratch[4 + ray], scratch[4 + N + ray]; the mean
// is taken by dp_loss_finish_kernel in a fixed order (bit-reproducible, no same-address atomics).
template <int G, int C, int K>
__global__ void __launch_bounds__(128) dp_loss_fwd_fast_kernel(DpArgs a, float* __restrict__ scratch, int per_ray) {
    extern __shared__ __align__(16) float smem[];
    const int grp = threadIdx.x / G, gl = threadIdx.x % G;
    int64_t ray = (int64_t)blockIdx.x * (blockDim.x / G) + grp;
    const bool valid = ray < a.N;
    if (!valid) ray = a.N - 1;
    using D = CellDp<G, C, K, false>;
    D sm(smem + (size_t)grp * per_ray, a.S1);
    typename D::Fine fs;
    RayState rs = sm.forward(a, ray, gl, fs);
    float l = 0.f;
#pragma unroll
    for (int c = 0; c < K; ++c) {
        const int k = c * G + gl;
        if (k >= a.S1) break;
        float q = sm.E[k + 1] - sm.E[k];
        float qe = (q < 0.f ? 0.f : q) + EPS;
        float pe = fs.w1[c] + EPS;
        l += pe * lg2_(pe * rcp_(qe));
    }
    l = group_sum<G>(l);
    if (gl == 0 && valid) {
        scratch[4 + ray] = rs.relevant ? l * LN2 / rs.Z1 + logf(rs.Zq / rs.Z1) : 0.f;
        scratch[4 + a.N + ray] = rs.relevant ? 1.f : 0.f;
    }
}

__global__ void __launch_bounds__(1024) dp_loss_finish_kernel(float* __restrict__ scratch, float* __restrict__ loss_out,
                                                               int64_t N, int S1, DpTotal tt) {
    __shared__ float ssum[32], scnt[32];
    float sum = 0.f, cnt = 0.f;
    for (int64_t i = threadIdx.x; i < N; i += 1024) { sum += scratch[4 + i]; cnt += scratch[4 + N + i]; }
    sum = group_sum<32>(sum); cnt = group_sum<32>(cnt);
    if ((threadIdx.x & 31) == 0) { ssum[threadIdx.x >> 5] = sum; scnt[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x < 32) {
        sum = group_sum<32>(ssum[threadIdx.x]); cnt = group_sum<32>(scnt[threadIdx.x]);
        if (threadIdx.x == 0) {
            scratch[0] = sum; scratch[1] = cnt;
            *loss_out = dp_total(tt, cnt > 0.f ? sum / (cnt * (float)S1) : 0.f);     // reduction='mean' over kept rays x S1
        }
    }
}

template <int G, int C, int K>
__global__ void __launch_bounds__(128) dp_loss_bwd_fast_kernel(DpArgs a, DpTotal tt, const float* __restrict__ g_loss,
                                                                const float* __restrict__ scratch,
                                                                float* __restrict__ g_w0, float* __restrict__ g_mus0,
                                                                float* __restrict__ g_sig0, int per_ray) {
    extern __shared__ __align__(16) float smem[];
    constexpr int P = G * C;
    dp_total_g_regs(tt, g_loss);
    const int grp = threadIdx.x / G, gl = threadIdx.x % G;
    int64_t ray = (int64_t)blockIdx.x * (blockDim.x / G) + grp;
    const bool valid = ray < a.N;
    if (!valid) ray = a.N - 1;
    const int S0 = a.S0, S1 = a.S1;
    using D = CellDp<G, C, K, true>;
    D sm(smem + (size_t)grp * per_ray, S1);
    typename D::Fine fs;
    for (int i = gl; i < 4 * P; i += G) sm.acc[i] = 0.f;  // [4][P]: sum gE, gE F, gE phi, gE phi x per coarse cell
    RayState rs = sm.forward(a, ray, gl, fs);             // (its __syncwarp()s order the zero fill)
    const float cnt = __ldg(scratch + 1);
    const bool live = rs.relevant && cnt > 0.f;           // uniform over the lane group
    const float scale = live ? __ldg(g_loss) * tt.scale / (cnt * (float)S1) : 0.f;
    const float zr = rs.Zq / rs.Z1, sz = scale / rs.Zq;
    float carry = 0.f;                                    // gq of the previous chunk's last edge
#pragma unroll
    for (int c = 0; c < K; ++c) {
        const int k = c * G + gl;
        if (c * G > S1) break;                            // uniform
        // dL/dq_k = scale/Zq (1 - p1/qn), zero where the clamp q<0 -> 0 is active
        float gqk = 0.f;
        if (k < S1) {
            const float q = sm.E[k + 1] - sm.E[k];
            if (!(q < 0.f)) gqk = sz * (1.0f - (fs.w1[c] + EPS) * rcp_(q + EPS) * zr);
        }
        float gprev = __shfl_up_sync(FULL, gqk, 1, G);
        if (gl == 0) gprev = carry;
        carry = __shfl_sync(FULL, gqk, G - 1, G);
        const float gE = gprev - gqk;                     // E_k enters q_{k-1} with +1 and q_k with -1
        const int j = fs.j[c];
        const bool active = live && j >= 0 && gE != 0.f;
        const float x = fs.x[c];
        const float v2 = gE * (0.3989422804f * ex2_(-0.5f * x * x * L2E));
        // Scatter into the coarse cells.  The cells of consecutive edges are non-decreasing, so the edges of one
        // cell are a run of adjacent lanes: a segmented inclusive scan (the lane d below has the same cell iff the
        // whole stretch has) leaves each run's total in its last lane, which alone updates shared memory -- no
        // atomics (float atomics on shared memory are CAS loops) and a fixed summation order.
        const int jr = k > S1 ? 0x7fffffff : (j < 0 ? ~j : j);           // the edge's cell, clamped or not
        float s0 = active ? gE : 0.f, s1 = active ? gE * fs.F[c] : 0.f, s2 = active ? v2 : 0.f, s3 = active ? v2 * x : 0.f;
#pragma unroll
        for (int d = 1; d < G; d <<= 1) {
            const int jn = __shfl_up_sync(FULL, jr, d, G);
            const float n0 = __shfl_up_sync(FULL, s0, d, G), n1 = __shfl_up_sync(FULL, s1, d, G);
            const float n2 = __shfl_up_sync(FULL, s2, d, G), n3 = __shfl_up_sync(FULL, s3, d, G);
            if (gl >= d && jn == jr) { s0 += n0; s1 += n1; s2 += n2; s3 += n3; }
        }
        const int jnext = __shfl_down_sync(FULL, jr, 1, G);
        if ((gl == G - 1 || jnext != jr) && jr < S0 && (s0 != 0.f || s1 != 0.f || s2 != 0.f || s3 != 0.f)) {
            float* p = sm.acc + jr;
            p[0] += s0; p[P] += s1; p[2 * P] += s2; p[3 * P] += s3;
        }
        __syncwarp();
    }
    // cdf[m] = min(1, cum[m]): route g_cdf[m] to p0[0..m-1] for m = 1..S0-1 (suffix sum).  Lane owns cells e0..e0+C-1.
    const int e0 = gl * C;
    float h[C], tail = 0.f;
#pragma unroll
    for (int c = C - 1; c >= 0; --c) {                    // h[c] = sum over m in (e0+c, e0+C] of routed g_cdf[m]
        const int m = e0 + c + 1;
        if (m <= S0 - 1) tail += sm.acc[m] * sm.tie[c];
        h[c] = tail;
    }
    const float suf_incl = group_suffix_sum<G>(tail, gl);
    const float after = suf_incl - tail;                  // routed g_cdf of all later lanes
    float gp[C], pc[C], dot = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int i = e0 + c;
        pc[c] = sm.p0[skew(i)];
        gp[c] = i < S0 ? sm.acc[P + i] + h[c] + after : 0.f;
        dot += i < S0 ? gp[c] * pc[c] : 0.f;
    }
    dot = group_sum<G>(dot);
    if (valid) {
        const float iz0 = 1.0f / rs.Z0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int i = e0 + c;
            if (i < S0) {
                const float kc = -pc[c] * sm.ipin[i] * sm.isig[i];        // d x / d mu, d x / d sigma carry -p0/(pin sigma)
                g_w0[ray * S0 + i] = live ? (gp[c] - dot) * iz0 : 0.f;    // p0 = (w0+eps)/sum(w0+eps)
                g_mus0[ray * S0 + i] = live ? kc * sm.acc[2 * P + i] : 0.f;
                g_sig0[ray * S0 + i] = live ? kc * sm.acc[3 * P + i] : 0.f;
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
    // Four cells per lane: measured faster than eight (profiles/r01b_*): these kernels are bound by latency, and
    // the extra resident warps of the smaller per-lane state outweigh the shorter scans of fewer, fatter lanes.
    DDNERF_FAST(4, 1) DDNERF_FAST(4, 2) DDNERF_FAST(4, 4) DDNERF_FAST(8, 4) DDNERF_FAST(16, 4) DDNERF_FAST(32, 4) DDNERF_FAST(32, 8)
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

static int dp_loss_forward_impl(const float* t1, const float* t0, const float* w1, const float* w0, const float* mus0,
                                const float* sigmas0, const float* lt0, const float* pin0, int blender, DpTotal tt,
                                float* loss_out, float* scratch, int64_t N, int S0, int S1, void* stream) {
    DDNERF_CHECK_ARG(loss_out && scratch && (N == 0 || (t1 && t0 && w1 && w0 && mus0 && sigmas0)), "dp_loss_forward: null pointer");
    DDNERF_CHECK_ARG((lt0 == nullptr) == (pin0 == nullptr), "dp_loss_forward: lt0 and pin0 go together (both NULL: computed in the kernel)");
    DDNERF_CHECK_ARG(S0 >= 1 && S1 >= 1 && S0 <= 1024 && S1 <= 1024, "dp_loss_forward: S0=%d S1=%d unsupported", S0, S1);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (N == 0) {                                           // no rays: kl = 0 (and the header the backward reads)
        cudaMemsetAsync(scratch, 0, 4 * sizeof(float), st);
        dp_loss_finish_kernel<<<1, 1024, 0, st>>>(scratch, loss_out, 0, S1, tt);
        DDNERF_LAUNCHED("dp_loss_forward", 1);
        return 0;
    }
    DpArgs a{t1, t0, w1, w0, mus0, sigmas0, lt0, pin0, blender, N, S0, S1};
    bool fast = dispatch_fast(S0, S1, [&](auto g, auto c, auto kk) {
        constexpr int G = decltype(g)::value, C = decltype(c)::value, K = decltype(kk)::value;
        const int per_ray = CellDp<G, C, K, false>::floats(S1);
        const size_t bytes = (size_t)(128 / G) * per_ray * sizeof(float);
        if (bytes > 200 * 1024) return false;
        auto kern = dp_loss_fwd_fast_kernel<G, C, K>;
        if (bytes > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        kern<<<ceil_div(N, 128 / G), 128, bytes, st>>>(a, scratch, per_ray);
        dp_loss_finish_kernel<<<1, 1024, 0, st>>>(scratch, loss_out, N, S1, tt);   // (writes the header: no memset needed)
        ddnerf::count_launches(1);
        return true;
    });
    if (!fast) {
        int per_warp = Smem::floats(S0, S1);
        int wpb = warps_for(per_warp * sizeof(float));
        DDNERF_CHECK_ARG(wpb >= 1, "dp_loss_forward: shapes need too much shared memory");
        cudaMemsetAsync(scratch, 0, 4 * sizeof(float), st);                        // the generic kernel accumulates with atomics
        dp_loss_fwd_kernel<<<ceil_div(N, wpb), wpb * 32, (size_t)wpb * per_warp * sizeof(float), st>>>(a, tt, loss_out, scratch,
                                                                                                       per_warp);
    }
    DDNERF_LAUNCHED("dp_loss_forward", 1);
    return 0;
}

static int dp_loss_backward_impl(const float* t1, const float* t0, const float* w1, const float* w0, const float* mus0,
                                 const float* sigmas0, const float* lt0, const float* pin0, int blender, DpTotal tt,
                                 const float* g_loss, const float* scratch, float* g_w0, float* g_mus0, float* g_sigmas0,
                                 int64_t N, int S0, int S1, void* stream) {
    DDNERF_CHECK_ARG(g_loss && scratch && (N == 0 || (t1 && t0 && w1 && w0 && mus0 && sigmas0 && g_w0 && g_mus0 && g_sigmas0)),
                     "dp_loss_backward: null pointer");
    DDNERF_CHECK_ARG((lt0 == nullptr) == (pin0 == nullptr), "dp_loss_backward: lt0 and pin0 go together");
    DDNERF_CHECK_ARG(S0 >= 1 && S1 >= 1 && S0 <= 1024 && S1 <= 1024, "dp_loss_backward: S0=%d S1=%d unsupported", S0, S1);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (N == 0) {
        if (tt.g_regs) {                                    // the regularisers still receive the cotangent
            cudaMemsetAsync(tt.g_regs, 0, 2 * sizeof(float), st);
            cudaMemcpyAsync(tt.g_regs + 2, g_loss, sizeof(float), cudaMemcpyDeviceToDevice, st);
            cudaMemcpyAsync(tt.g_regs + 3, g_loss, sizeof(float), cudaMemcpyDeviceToDevice, st);
        }
        return 0;
    }
    DpArgs a{t1, t0, w1, w0, mus0, sigmas0, lt0, pin0, blender, N, S0, S1};
    bool fast = dispatch_fast(S0, S1, [&](auto g, auto c, auto kk) {
        constexpr int G = decltype(g)::value, C = decltype(c)::value, K = decltype(kk)::value;
        const int per_ray = CellDp<G, C, K, true>::floats(S1);
        const size_t bytes = (size_t)(128 / G) * per_ray * sizeof(float);
        if (bytes > 200 * 1024) return false;
        auto kern = dp_loss_bwd_fast_kernel<G, C, K>;
        if (bytes > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        kern<<<ceil_div(N, 128 / G), 128, bytes, st>>>(a, tt, g_loss, scratch, g_w0, g_mus0, g_sigmas0, per_ray);
        return true;
    });
    if (!fast) {
        int per_warp = Smem::floats(S0, S1) + (S0 + 1) + 3 * S0;
        int wpb = warps_for(per_warp * sizeof(float));
        DDNERF_CHECK_ARG(wpb >= 1, "dp_loss_backward: shapes need too much shared memory");
        dp_loss_bwd_kernel<<<ceil_div(N, wpb), wpb * 32, (size_t)wpb * per_warp * sizeof(float), st>>>(
            a, tt, g_loss, scratch, g_w0, g_mus0, g_sigmas0, per_warp);
    }
    DDNERF_LAUNCHED("dp_loss_backward", 1);
    return 0;
}

extern "C" DDNERF_EXPORT int ddnerf_dp_loss_forward(const float* t1, const float* t0, const float* w1, const float* w0,
                                      const float* mus0, const float* sigmas0, const float* lt0, const float* pin0,
                                      int blender, float* loss_out, float* scratch, int64_t N, int S0, int S1,
                                      void* stream) {
    return dp_loss_forward_impl(t1, t0, w1, w0, mus0, sigmas0, lt0, pin0, blender, DpTotal{1.0f, nullptr, nullptr}, loss_out,
                                scratch, N, S0, S1, stream);
}

extern "C" DDNERF_EXPORT int ddnerf_dp_loss_backward(const float* t1, const float* t0, const float* w1, const float* w0,
                                       const float* mus0, const float* sigmas0, const float* lt0, const float* pin0,
                                       int blender, const float* g_loss, const float* scratch, float* g_w0,
                                       float* g_mus0, float* g_sigmas0, int64_t N, int S0, int S1, void* stream) {
    return dp_loss_backward_impl(t1, t0, w1, w0, mus0, sigmas0, lt0, pin0, blender, DpTotal{1.0f, nullptr, nullptr}, g_loss,
                                 scratch, g_w0, g_mus0, g_sigmas0, N, S0, S1, stream);
}

extern "C" DDNERF_EXPORT int ddnerf_dp_loss_total_forward(const float* t1, const float* t0, const float* w1, const float* w0,
                                            const float* mus0, const float* sigmas0, const float* lt0, const float* pin0,
                                            int blender, float scale, const float* regs, float* loss_out, float* scratch,
                                            int64_t N, int S0, int S1, void* stream) {
    DDNERF_CHECK_ARG(regs, "dp_loss_total_forward: regs is NULL");
    return dp_loss_forward_impl(t1, t0, w1, w0, mus0, sigmas0, lt0, pin0, blender, DpTotal{scale, regs, nullptr}, loss_out,
                                scratch, N, S0, S1, stream);
}

extern "C" DDNERF_EXPORT int ddnerf_dp_loss_total_backward(const float* t1, const float* t0, const float* w1, const float* w0,
                                             const float* mus0, const float* sigmas0, const float* lt0, const float* pin0,
                                             int blender, float scale, const float* g_loss, const float* scratch,
                                             float* g_w0, float* g_mus0, float* g_sigmas0, float* g_regs, int64_t N, int S0,
                                             int S1, void* stream) {
    DDNERF_CHECK_ARG(g_regs, "dp_loss_total_backward: g_regs is NULL");
    return dp_loss_backward_impl(t1, t0, w1, w0, mus0, sigmas0, lt0, pin0, blender, DpTotal{scale, nullptr, g_regs}, g_loss,
                                 scratch, g_w0, g_mus0, g_sigmas0, N, S0, S1, stream);
}
