// K4 / K4b: alpha compositing forward and analytic backward, one lane-group per ray.
//
// Replaces general_utils/volume_rendering_utils.py:6-84 (+ cumprod_exclusive,
// nerf_helpers.py:43-64) of the reference.  HBM-bound by design: every sample is read once (raw
// rgb/density, fence-posts, noise, mu) and its weight written once; the transmittance product and all
// per-ray sums are lane-group scans/reductions in registers.  A ray is owned by G = 4..32 lanes and
// walked in NCH chunks of G samples, so adjacent lanes read adjacent samples (coalesced 16 B/lane when
// raw is [N,S,4]); short rays share a warp, which spreads the per-ray prologue/epilogue.
//
// Round-1 profile (profiles/r01_ncu_composite_*.md): the first version was instruction-issue bound
// (274 thread instructions per sample forward, issue-active 88 %), most of them in full-precision
// expf/log1pf/division sequences.  This version (a) issues every load of a ray before any math, (b) uses
// the SFU (ex2/lg2/rcp.approx, <= 2 ulp) with argument handling that keeps the relative error of
// softplus, sigmoid and alpha at a few 1e-6 -- two orders below the 1e-4 the parity tests hold --,
// (c) in the backward keeps the per-sample forward quantities in registers between the forward scan and
// the suffix scan instead of reloading and recomputing them, and (d) resolves the optional inputs
// (mu, exact-fit S, raw layout) at compile time so the per-sample code carries no dead branches.
#include <type_traits>

#include "common.cuh"

namespace ddnerf {
namespace {

struct CompositeArgs {
    const float* raw; int raw_stride;
    const float* t; const float* rd; int64_t rd_stride;
    const float* noise; float noise_std;
    const float* mus;
    int white, blender;
    int64_t N; int S;
};

__device__ __forceinline__ float torch_max_(float a, float b) {   // torch.max propagates NaN
    return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b);
}

// sigmoid(x) = 1 / (1 + e^-x); saturates to exactly 0 / 1 like the reference (1/(1+inf) = 0)
__device__ __forceinline__ float sigmoid_sfu(float x) { return rcp_(1.0f + ex2_(-x * L2E)); }

// softplus(x) (beta 1, threshold 20) and its derivative sigmoid(x) from one exponential.
// log1p(e): four-term series below e = 1/16 (relative error < 4e-6), lg2(1+e) above (absolute error
// 2^-22 on a value >= 0.06).
template <bool DERIV>
__device__ __forceinline__ void softplus_sfu(float x, float& sp, float& dsp) {
    float e = ex2_(x * L2E);
    float u = 1.0f + e;
    float series = e * fmaf(e, fmaf(e, fmaf(e, -0.25f, 0.33333334f), -0.5f), 1.0f);
    float lg = lg2_(u) * LN2;
    sp = e < 0.0625f ? series : lg;
    if (DERIV) dsp = e * rcp_(u);
    if (x > 20.0f) { sp = x; if (DERIV) dsp = 1.0f; }
}

// quotient for the per-ray epilogues: fast when the divisor is comfortably normal, IEEE otherwise
__device__ __forceinline__ float div_ray(float a, float b) { return fabsf(b) > 1e-30f ? div_fast(a, b) : a / b; }

// ---- lane-group primitives.  Every shuffle names the FULL warp: a computed sub-warp member mask makes the
// compiler wrap each shuffle in WARPSYNC/BSSY/BSYNC convergence code (5 extra instructions per shuffle), so
// lane groups whose ray is out of range stay alive on a clamped ray index and only their stores are gated. ----
template <int G>
__device__ __forceinline__ unsigned group_mask() { return FULL; }
template <int G>
__device__ __forceinline__ float gsum(float v, unsigned m) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(m, v, o, G);
    return v;
}
template <int G>
__device__ __forceinline__ float gscan_prod(float v, int gl, unsigned m) {       // inclusive product
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
        float n = __shfl_up_sync(m, v, o, G);
        if (gl >= o) v *= n;
    }
    return v;
}
template <int G>
__device__ __forceinline__ float gscan_suffix(float v, int gl, unsigned m) {     // inclusive suffix sum
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
        float n = __shfl_down_sync(m, v, o, G);
        if (gl + o < G) v += n;
    }
    return v;
}

// RAWV: 4 = one 16-byte load per sample ([N,S,4] contiguous), 2 = two 8-byte loads (even channel stride,
// e.g. the 6-channel output of the DDNeRF coarse network), 1 = scalar loads.
template <int G, int NCH, int RAWV, bool MU, bool EXACT>
struct Loaded {
    float4 rv[NCH]; float t0[NCH], t1[NCH], nz[NCH], mu[NCH];
    __device__ __forceinline__ void load(const CompositeArgs& a, int64_t ray, int gl) {
        const int S = EXACT ? G * NCH : a.S;
        const float* tp = a.t + ray * (S + 1);
        const int64_t row0 = ray * S;
        const bool has_noise = a.noise != nullptr;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const int i = ch * G + gl;
            rv[ch] = make_float4(0.f, 0.f, 0.f, 0.f); t0[ch] = t1[ch] = 0.f; nz[ch] = 0.f; mu[ch] = 0.f;
            if (EXACT || i < S) {
                if (RAWV == 4) {
                    rv[ch] = __ldg(reinterpret_cast<const float4*>(a.raw) + row0 + i);
                } else if (RAWV == 2) {
                    const float2* rp = reinterpret_cast<const float2*>(a.raw + (row0 + i) * a.raw_stride);
                    float2 lo = __ldg(rp), hi = __ldg(rp + 1);
                    rv[ch] = make_float4(lo.x, lo.y, hi.x, hi.y);
                } else {
                    const float* rp = a.raw + (row0 + i) * a.raw_stride;
                    rv[ch] = make_float4(__ldg(rp), __ldg(rp + 1), __ldg(rp + 2), __ldg(rp + 3));
                }
                t0[ch] = __ldg(tp + i); t1[ch] = __ldg(tp + i + 1);
                if (has_noise) nz[ch] = __ldg(a.noise + row0 + i);
                if (MU) mu[ch] = __ldg(a.mus + row0 + i);
            }
        }
    }
};

__device__ __forceinline__ float ray_norm(const CompositeArgs& a, int64_t ray) {
    const float* d = a.rd + ray * a.rd_stride;
    float x = __ldg(d), y = __ldg(d + 1), z = __ldg(d + 2);
    return sqrtf(x * x + y * y + z * z);
}

template <int G, int NCH, int RAWV, bool MU, bool EXACT>
__global__ void __launch_bounds__(256) composite_fwd_kernel(CompositeArgs a, float* __restrict__ rgb_map,
                                                             float* __restrict__ disp, float* __restrict__ acc,
                                                             float* __restrict__ weights, float* __restrict__ depth,
                                                             float* __restrict__ cdisp, float* __restrict__ rgb) {
    const int gl = threadIdx.x % G;
    int64_t ray = (int64_t)blockIdx.x * (blockDim.x / G) + threadIdx.x / G;
    const bool valid = ray < a.N;                              // uniform over the lane group
    if (!valid) ray = a.N - 1;
    const unsigned m = group_mask<G>();
    const int S = EXACT ? G * NCH : a.S;
    Loaded<G, NCH, RAWV, MU, EXACT> L;
    L.load(a, ray, gl);
    const float norm_d = ray_norm(a, ray);
    const int64_t row0 = ray * S;
    // rgb_map = 1.002 sum(w sigmoid) - 0.001 sum(w), with w BEFORE the blender epsilon
    float T_carry = 1.f, s_w0 = 0.f, s_w = 0.f, s_wm = 0.f, s_ws = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        const int i = ch * G + gl;
        const bool in = EXACT || i < S;
        const float dist = L.t1[ch] - L.t0[ch];
        const float mid = (L.t1[ch] + L.t0[ch]) * 0.5f;
        const float dens = L.rv[ch].w + L.nz[ch] * a.noise_std;
        float sa, unused;
        softplus_sfu<false>(dens - 1.0f, sa, unused);
        const float alpha = 1.0f - ex2_(-(sa * (dist * norm_d)) * L2E);
        const float tm = in ? (1.0f - alpha) + 1e-10f : 1.0f;
        float incl = gscan_prod<G>(tm, gl, m);
        float excl = __shfl_up_sync(m, incl, 1, G);
        if (gl == 0) excl = 1.f;
        const float T = T_carry * excl;
        if (NCH > 1) T_carry *= __shfl_sync(m, incl, G - 1, G);
        const float w = in ? alpha * T : 0.f;
        const float sr = sigmoid_sfu(L.rv[ch].x), sg = sigmoid_sfu(L.rv[ch].y), sb = sigmoid_sfu(L.rv[ch].z);
        cr += w * sr; cg += w * sg; cb += w * sb; s_w0 += w;
        float wp = w;
        if (a.blender && i == S - 1) wp = w + 1e-10f;          // volume_rendering_utils.py:52-56
        if (in && valid) {
            weights[row0 + i] = wp;
            if (rgb) {
                float* o = rgb + (row0 + i) * 3;
                o[0] = sr * 1.002f - 0.001f; o[1] = sg * 1.002f - 0.001f; o[2] = sb * 1.002f - 0.001f;
            }
        }
        s_w += wp; s_wm += wp * mid;
        if (MU) s_ws += wp * (L.t0[ch] + L.mu[ch] * dist);
    }
    s_w = gsum<G>(s_w, m); s_wm = gsum<G>(s_wm, m); s_w0 = gsum<G>(s_w0, m);
    if (MU) s_ws = gsum<G>(s_ws, m);
    cr = gsum<G>(cr, m); cg = gsum<G>(cg, m); cb = gsum<G>(cb, m);
    if (gl == 0 && valid) {
        const float W = s_w;
        cr = cr * 1.002f - 0.001f * s_w0; cg = cg * 1.002f - 0.001f * s_w0; cb = cb * 1.002f - 0.001f * s_w0;
        const float d0 = a.blender ? div_ray(s_wm, W) : s_wm;
        const float dsp = div_fast(1.0f, torch_max_(1e-10f, div_ray(d0, W)));
        if (a.white) { cr += 1.0f - W; cg += 1.0f - W; cb += 1.0f - W; }
        rgb_map[ray * 3] = cr; rgb_map[ray * 3 + 1] = cg; rgb_map[ray * 3 + 2] = cb;
        acc[ray] = W; disp[ray] = dsp;
        if (MU) {
            const float cd = a.blender ? div_ray(s_ws, W) : s_ws;
            cdisp[ray] = div_fast(1.0f, torch_max_(1e-10f, div_ray(cd, W)));
            depth[ray] = cd;                                   // :83 depth_map := corrected
        } else {
            depth[ray] = d0;
        }
    }
}

struct CompositeGrads {
    const float* g_rgb_map; const float* g_disp; const float* g_acc; const float* g_weights;
    const float* g_depth; const float* g_cdisp;
};

template <int G, int NCH, int RAWV, bool MU, bool EXACT>
__global__ void __launch_bounds__(256) composite_bwd_kernel(CompositeArgs a, CompositeGrads gr,
                                                             float* __restrict__ g_raw, float* __restrict__ g_mus) {
    const int gl = threadIdx.x % G;
    int64_t ray = (int64_t)blockIdx.x * (blockDim.x / G) + threadIdx.x / G;
    const bool valid = ray < a.N;
    if (!valid) ray = a.N - 1;
    const unsigned m = group_mask<G>();
    const int S = EXACT ? G * NCH : a.S;
    const int64_t row0 = ray * S;
    // per-sample forward quantities, kept in registers between the two scans
    float sr[NCH], sg[NCH], sb[NCH], al[NCH], Tr[NCH], ds[NCH], t0[NCH], dist[NCH], mu[NCH], gwt[NCH];
    float s_w = 0.f, s_wm = 0.f, s_ws = 0.f;
    const bool has_gw = gr.g_weights != nullptr;
    float norm_d;
    {
        Loaded<G, NCH, RAWV, MU, EXACT> L;
        L.load(a, ray, gl);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {                     // cotangent of the weights: one more streamed input
            const int i = ch * G + gl;
            gwt[ch] = (has_gw && (EXACT || i < S)) ? __ldg(gr.g_weights + row0 + i) : 0.f;
        }
        norm_d = ray_norm(a, ray);
        // phase 1: the forward scan (alpha_i, T_i) and the per-ray sums
        float T_carry = 1.f;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const int i = ch * G + gl;
            const bool in = EXACT || i < S;
            t0[ch] = L.t0[ch]; mu[ch] = L.mu[ch];
            dist[ch] = L.t1[ch] - L.t0[ch];
            const float dens = L.rv[ch].w + L.nz[ch] * a.noise_std;
            float sa;
            softplus_sfu<true>(dens - 1.0f, sa, ds[ch]);
            al[ch] = 1.0f - ex2_(-(sa * (dist[ch] * norm_d)) * L2E);
            const float tm = in ? (1.0f - al[ch]) + 1e-10f : 1.0f;
            float incl = gscan_prod<G>(tm, gl, m);
            float excl = __shfl_up_sync(m, incl, 1, G);
            if (gl == 0) excl = 1.f;
            Tr[ch] = T_carry * excl;
            if (NCH > 1) T_carry *= __shfl_sync(m, incl, G - 1, G);
            sr[ch] = sigmoid_sfu(L.rv[ch].x); sg[ch] = sigmoid_sfu(L.rv[ch].y); sb[ch] = sigmoid_sfu(L.rv[ch].z);
            float wp = in ? al[ch] * Tr[ch] : 0.f;
            if (a.blender && i == S - 1) wp += 1e-10f;
            s_w += wp; s_wm += wp * (t0[ch] + 0.5f * dist[ch]);
            if (MU) s_ws += wp * (t0[ch] + mu[ch] * dist[ch]);
        }
    }
    s_w = gsum<G>(s_w, m); s_wm = gsum<G>(s_wm, m);
    if (MU) s_ws = gsum<G>(s_ws, m);
    // per-ray cotangents folded to (Ga on acc, Gd on depth0, Gc on corrected depth)
    float Gr = 0.f, Gg = 0.f, Gb = 0.f, Ga = 0.f, Gd = 0.f, Gc = 0.f;
    const float W = s_w;
    const float d0 = a.blender ? div_ray(s_wm, W) : s_wm;
    const float cd = MU ? (a.blender ? div_ray(s_ws, W) : s_ws) : 0.f;
    if (gr.g_rgb_map) { Gr = __ldg(gr.g_rgb_map + ray * 3); Gg = __ldg(gr.g_rgb_map + ray * 3 + 1); Gb = __ldg(gr.g_rgb_map + ray * 3 + 2); }
    if (gr.g_acc) Ga = __ldg(gr.g_acc + ray);
    if (a.white) Ga -= Gr + Gg + Gb;
    {
        const float gdep = gr.g_depth ? __ldg(gr.g_depth + ray) : 0.f;
        if (MU) Gc = gdep; else Gd = gdep;
    }
    if (gr.g_disp) {                                           // NaN (0/0, empty non-blender ray) propagates like torch
        float x = d0 / W;
        if (x > 1e-10f || x != x) { float k = -__ldg(gr.g_disp + ray) / (x * x); Gd += k / W; Ga += -k * d0 / (W * W); }
    }
    if (MU && gr.g_cdisp) {
        float x = cd / W;
        if (x > 1e-10f || x != x) { float k = -__ldg(gr.g_cdisp + ray) / (x * x); Gc += k / W; Ga += -k * cd / (W * W); }
    }
    // g_i = Gr' sr + Gg' sg + Gb' sb + G0 + g_w_i + depth terms, with the 1.002 s - 0.001 of the colours folded in
    const float Gr1 = Gr * 1.002f, Gg1 = Gg * 1.002f, Gb1 = Gb * 1.002f;
    const float G0 = Ga - 0.001f * (Gr + Gg + Gb);
    // depth terms per sample: blender  Gd (mid - d0)/W + Gc (spos - cd)/W ; else  Gd mid + Gc spos
    const float kd = a.blender ? Gd / W : Gd, kc = a.blender ? Gc / W : Gc;
    const float od = a.blender ? d0 : 0.f, oc = a.blender ? cd : 0.f;
    const bool depth_terms = (Gd != 0.f) || (Gc != 0.f) || (Gd != Gd) || (Gc != Gc);
    // phase 2: walk the chunks backwards with the suffix sum A_i = sum_{k>i} g_k w_k
    float A_carry = 0.f;
#pragma unroll
    for (int ch = NCH - 1; ch >= 0; --ch) {
        const int i = ch * G + gl;
        const bool in = EXACT || i < S;
        const float T = Tr[ch], alpha = al[ch];
        const float w = alpha * T;
        float g = Gr1 * sr[ch] + Gg1 * sg[ch] + Gb1 * sb[ch] + G0 + gwt[ch];
        if (depth_terms) {
            g += kd * ((t0[ch] + 0.5f * dist[ch]) - od);
            if (MU) g += kc * ((t0[ch] + mu[ch] * dist[ch]) - oc);
        }
        const float gw = in ? g * w : 0.f;
        float suf = gscan_suffix<G>(gw, gl, m);
        const float A = A_carry + suf - gw;
        if (NCH > 1) A_carry += __shfl_sync(m, suf, 0, G);
        if (in && valid) {
            const float one_m_alpha = 1.0f - alpha;            // = exp(-sigma_a delta) as the forward rounded it
            const float tm = one_m_alpha + 1e-10f;
            const float g_alpha = g * T - A * rcp_(tm);
            float4 o;
            o.x = Gr1 * w * (sr[ch] - sr[ch] * sr[ch]);
            o.y = Gg1 * w * (sg[ch] - sg[ch] * sg[ch]);
            o.z = Gb1 * w * (sb[ch] - sb[ch] * sb[ch]);
            o.w = g_alpha * (dist[ch] * norm_d) * one_m_alpha * ds[ch];
            reinterpret_cast<float4*>(g_raw)[row0 + i] = o;
            if (MU && g_mus) {
                float wp = (a.blender && i == S - 1) ? w + 1e-10f : w;
                float p = a.blender ? wp / W : wp;
                g_mus[row0 + i] = Gc * p * dist[ch];           // d cdepth / d mu_i = p_i * (t_{i+1}-t_i)
            }
        }
    }
}


// ---- DDNeRF coarse pass: the compositor with the depth-distribution head folded in (models.py:242-273) -----------------
// raw6 [N,S,6] = (r, g, b, density, raw_mu, raw_sigma) is read in place (three 8-byte loads per sample).  On top of the
// compositor (with mu = sigmoid(raw_mu) feeding the corrected depth) the forward writes mus [N,S], sigmas = sigmoid(raw_sigma)
// + 0.001 [N,S] and regs[4] = {mus_loss, sig_loss, mus_reg, sig_reg} = {sum raw_mu^2 / N, sum raw_sigma^2 / N, coef x each}
// (models.py:245-252); the tails Phi((0 - mu) / sigma), Phi((1 - mu) / sigma) of :254-258 and their smoothed versions
// (:268-273) are per-cell work of the two consumers (resampler, dp-loss) and never touch HBM.  The backward takes the
// cotangents of the maps, of the weights, of mus / sigmas (dp-loss) and of regs, and writes the cotangent of raw6 directly.
// The regulariser sums are reduced in a fixed order (per block, then by the last block to finish): bit-reproducible.
template <int G, int NCH, bool EXACT>
struct LoadedDD {
    float4 rv[NCH]; float t0[NCH], t1[NCH], nz[NCH], rm[NCH], rs[NCH];
    __device__ __forceinline__ void load(const CompositeArgs& a, int64_t ray, int gl) {
        const int S = EXACT ? G * NCH : a.S;
        const float* tp = a.t + ray * (S + 1);
        const int64_t row0 = ray * S;
        const bool has_noise = a.noise != nullptr;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const int i = ch * G + gl;
            rv[ch] = make_float4(0.f, 0.f, 0.f, 0.f); t0[ch] = t1[ch] = 0.f; nz[ch] = 0.f; rm[ch] = rs[ch] = 0.f;
            if (EXACT || i < S) {
                const float2* rp = reinterpret_cast<const float2*>(a.raw + (row0 + i) * 6);
                const float2 lo = __ldg(rp), hi = __ldg(rp + 1), ms = __ldg(rp + 2);
                rv[ch] = make_float4(lo.x, lo.y, hi.x, hi.y);
                rm[ch] = ms.x; rs[ch] = ms.y;
                t0[ch] = __ldg(tp + i); t1[ch] = __ldg(tp + i + 1);
                if (has_noise) nz[ch] = __ldg(a.noise + row0 + i);
            }
        }
    }
};

struct DDOut { float* mus; float* sigmas; float* regs; float* scratch; float coef; };

template <int G, int NCH, bool EXACT>
__global__ void __launch_bounds__(256) composite_dd_fwd_kernel(CompositeArgs a, DDOut dd, float* __restrict__ rgb_map,
                                                                float* __restrict__ disp, float* __restrict__ acc,
                                                                float* __restrict__ weights, float* __restrict__ depth,
                                                                float* __restrict__ cdisp) {
    const int gl = threadIdx.x % G;
    int64_t ray = (int64_t)blockIdx.x * (blockDim.x / G) + threadIdx.x / G;
    const bool valid = ray < a.N;
    if (!valid) ray = a.N - 1;
    const unsigned m = group_mask<G>();
    const int S = EXACT ? G * NCH : a.S;
    LoadedDD<G, NCH, EXACT> L;
    L.load(a, ray, gl);
    const float norm_d = ray_norm(a, ray);
    const int64_t row0 = ray * S;
    float T_carry = 1.f, s_w0 = 0.f, s_w = 0.f, s_wm = 0.f, s_ws = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, m2 = 0.f, s2 = 0.f;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        const int i = ch * G + gl;
        const bool in = EXACT || i < S;
        const float dist = L.t1[ch] - L.t0[ch];
        const float mid = (L.t1[ch] + L.t0[ch]) * 0.5f;
        const float dens = L.rv[ch].w + L.nz[ch] * a.noise_std;
        float sa, unused;
        softplus_sfu<false>(dens - 1.0f, sa, unused);
        const float alpha = 1.0f - ex2_(-(sa * (dist * norm_d)) * L2E);
        const float tm = in ? (1.0f - alpha) + 1e-10f : 1.0f;
        float incl = gscan_prod<G>(tm, gl, m);
        float excl = __shfl_up_sync(m, incl, 1, G);
        if (gl == 0) excl = 1.f;
        const float T = T_carry * excl;
        if (NCH > 1) T_carry *= __shfl_sync(m, incl, G - 1, G);
        const float w = in ? alpha * T : 0.f;
        const float sr = sigmoid_sfu(L.rv[ch].x), sg = sigmoid_sfu(L.rv[ch].y), sb = sigmoid_sfu(L.rv[ch].z);
        const float mu = sigmoid_sfu(L.rm[ch]);                          // models.py:245
        const float sig = sigmoid_sfu(L.rs[ch]) + 0.001f;                // models.py:246
        cr += w * sr; cg += w * sg; cb += w * sb; s_w0 += w;
        float wp = w;
        if (a.blender && i == S - 1) wp = w + 1e-10f;
        if (in && valid) {
            weights[row0 + i] = wp;
            dd.mus[row0 + i] = mu;
            dd.sigmas[row0 + i] = sig;
            m2 += L.rm[ch] * L.rm[ch];                                   // models.py:248-249
            s2 += L.rs[ch] * L.rs[ch];
        }
        s_w += wp; s_wm += wp * mid;
        s_ws += wp * (L.t0[ch] + mu * dist);
    }
    s_w = gsum<G>(s_w, m); s_wm = gsum<G>(s_wm, m); s_w0 = gsum<G>(s_w0, m); s_ws = gsum<G>(s_ws, m);
    cr = gsum<G>(cr, m); cg = gsum<G>(cg, m); cb = gsum<G>(cb, m);
    if (gl == 0 && valid) {
        const float W = s_w;
        cr = cr * 1.002f - 0.001f * s_w0; cg = cg * 1.002f - 0.001f * s_w0; cb = cb * 1.002f - 0.001f * s_w0;
        const float d0 = a.blender ? div_ray(s_wm, W) : s_wm;
        const float dsp = div_fast(1.0f, torch_max_(1e-10f, div_ray(d0, W)));
        if (a.white) { cr += 1.0f - W; cg += 1.0f - W; cb += 1.0f - W; }
        rgb_map[ray * 3] = cr; rgb_map[ray * 3 + 1] = cg; rgb_map[ray * 3 + 2] = cb;
        acc[ray] = W; disp[ray] = dsp;
        const float cd = a.blender ? div_ray(s_ws, W) : s_ws;
        cdisp[ray] = div_fast(1.0f, torch_max_(1e-10f, div_ray(cd, W)));
        depth[ray] = cd;
    }
    // ---- regulariser sums: warp -> block -> scratch; the last block to arrive adds the block partials in index order ----
    __shared__ float red[2][8];
    __shared__ unsigned is_last;
    m2 = group_sum<32>(m2); s2 = group_sum<32>(s2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    if (lane == 0) { red[0][warp] = m2; red[1][warp] = s2; }
    __syncthreads();
    unsigned* counter = reinterpret_cast<unsigned*>(dd.scratch + 2 * (size_t)gridDim.x);
    if (threadIdx.x == 0) {
        float b0 = 0.f, b1 = 0.f;
        for (int w = 0; w < nwarps; ++w) { b0 += red[0][w]; b1 += red[1][w]; }
        dd.scratch[2 * (size_t)blockIdx.x] = b0;
        dd.scratch[2 * (size_t)blockIdx.x + 1] = b1;
        __threadfence();
        is_last = atomicAdd(counter, 1u) == gridDim.x - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        float p0 = 0.f, p1 = 0.f;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
            p0 += __ldcg(dd.scratch + 2 * (size_t)b);
            p1 += __ldcg(dd.scratch + 2 * (size_t)b + 1);
        }
        p0 = group_sum<32>(p0); p1 = group_sum<32>(p1);
        __syncthreads();
        if (lane == 0) { red[0][warp] = p0; red[1][warp] = p1; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float t0 = 0.f, t1 = 0.f;
            for (int w = 0; w < nwarps; ++w) { t0 += red[0][w]; t1 += red[1][w]; }
            const float inv = 1.0f / (float)a.N;
            const float ml = t0 * inv, sl = t1 * inv;
            dd.regs[0] = ml; dd.regs[1] = sl; dd.regs[2] = dd.coef * ml; dd.regs[3] = dd.coef * sl;
            *counter = 0u;
        }
    }
}

struct DDGrads { const float* g_mus; const float* g_sigmas; const float* g_regs; float coef; };

template <int G, int NCH, bool EXACT>
__global__ void __launch_bounds__(256) composite_dd_bwd_kernel(CompositeArgs a, CompositeGrads gr, DDGrads dg,
                                                                float* __restrict__ g_raw6) {
    const int gl = threadIdx.x % G;
    int64_t ray = (int64_t)blockIdx.x * (blockDim.x / G) + threadIdx.x / G;
    const bool valid = ray < a.N;
    if (!valid) ray = a.N - 1;
    const unsigned m = group_mask<G>();
    const int S = EXACT ? G * NCH : a.S;
    const int64_t row0 = ray * S;
    float sr[NCH], sg[NCH], sb[NCH], al[NCH], Tr[NCH], ds[NCH], t0[NCH], dist[NCH], mu[NCH], gwt[NCH];
    float gm[NCH], gs[NCH];                                  // cotangent of raw_mu / raw_sigma without the corrected-depth term
    float s_w = 0.f, s_wm = 0.f, s_ws = 0.f;
    const bool has_gw = gr.g_weights != nullptr;
    float norm_d;
    {
        LoadedDD<G, NCH, EXACT> L;
        L.load(a, ray, gl);
        float kmu = 0.f, ksg = 0.f;                          // d regs / d raw = (g[0] + coef g[2]) 2 raw_mu / N, likewise sigma
        if (dg.g_regs) {
            const float inv2 = 2.0f / (float)a.N;
            kmu = (__ldg(dg.g_regs) + dg.coef * __ldg(dg.g_regs + 2)) * inv2;
            ksg = (__ldg(dg.g_regs + 1) + dg.coef * __ldg(dg.g_regs + 3)) * inv2;
        }
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const int i = ch * G + gl;
            const bool in = EXACT || i < S;
            gwt[ch] = (has_gw && in) ? __ldg(gr.g_weights + row0 + i) : 0.f;
            const float gmu_ext = (dg.g_mus && in) ? __ldg(dg.g_mus + row0 + i) : 0.f;
            const float gsg_ext = (dg.g_sigmas && in) ? __ldg(dg.g_sigmas + row0 + i) : 0.f;
            mu[ch] = sigmoid_sfu(L.rm[ch]);
            const float sgm = sigmoid_sfu(L.rs[ch]);
            gm[ch] = gmu_ext * (mu[ch] - mu[ch] * mu[ch]) + kmu * L.rm[ch];
            gs[ch] = gsg_ext * (sgm - sgm * sgm) + ksg * L.rs[ch];
        }
        norm_d = ray_norm(a, ray);
        float T_carry = 1.f;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const int i = ch * G + gl;
            const bool in = EXACT || i < S;
            t0[ch] = L.t0[ch];
            dist[ch] = L.t1[ch] - L.t0[ch];
            const float dens = L.rv[ch].w + L.nz[ch] * a.noise_std;
            float sa;
            softplus_sfu<true>(dens - 1.0f, sa, ds[ch]);
            al[ch] = 1.0f - ex2_(-(sa * (dist[ch] * norm_d)) * L2E);
            const float tm = in ? (1.0f - al[ch]) + 1e-10f : 1.0f;
            float incl = gscan_prod<G>(tm, gl, m);
            float excl = __shfl_up_sync(m, incl, 1, G);
            if (gl == 0) excl = 1.f;
            Tr[ch] = T_carry * excl;
            if (NCH > 1) T_carry *= __shfl_sync(m, incl, G - 1, G);
            sr[ch] = sigmoid_sfu(L.rv[ch].x); sg[ch] = sigmoid_sfu(L.rv[ch].y); sb[ch] = sigmoid_sfu(L.rv[ch].z);
            float wp = in ? al[ch] * Tr[ch] : 0.f;
            if (a.blender && i == S - 1) wp += 1e-10f;
            s_w += wp; s_wm += wp * (t0[ch] + 0.5f * dist[ch]);
            s_ws += wp * (t0[ch] + mu[ch] * dist[ch]);
        }
    }
    s_w = gsum<G>(s_w, m); s_wm = gsum<G>(s_wm, m); s_ws = gsum<G>(s_ws, m);
    float Gr = 0.f, Gg = 0.f, Gb = 0.f, Ga = 0.f, Gd = 0.f, Gc = 0.f;
    const float W = s_w;
    const float d0 = a.blender ? div_ray(s_wm, W) : s_wm;
    const float cd = a.blender ? div_ray(s_ws, W) : s_ws;
    if (gr.g_rgb_map) { Gr = __ldg(gr.g_rgb_map + ray * 3); Gg = __ldg(gr.g_rgb_map + ray * 3 + 1); Gb = __ldg(gr.g_rgb_map + ray * 3 + 2); }
    if (gr.g_acc) Ga = __ldg(gr.g_acc + ray);
    if (a.white) Ga -= Gr + Gg + Gb;
    if (gr.g_depth) Gc = __ldg(gr.g_depth + ray);            // the returned depth map is the corrected one
    if (gr.g_disp) {
        float x = d0 / W;
        if (x > 1e-10f || x != x) { float k = -__ldg(gr.g_disp + ray) / (x * x); Gd += k / W; Ga += -k * d0 / (W * W); }
    }
    if (gr.g_cdisp) {
        float x = cd / W;
        if (x > 1e-10f || x != x) { float k = -__ldg(gr.g_cdisp + ray) / (x * x); Gc += k / W; Ga += -k * cd / (W * W); }
    }
    const float Gr1 = Gr * 1.002f, Gg1 = Gg * 1.002f, Gb1 = Gb * 1.002f;
    const float G0 = Ga - 0.001f * (Gr + Gg + Gb);
    const float kd = a.blender ? Gd / W : Gd, kc = a.blender ? Gc / W : Gc;
    const float od = a.blender ? d0 : 0.f, oc = a.blender ? cd : 0.f;
    const bool depth_terms = (Gd != 0.f) || (Gc != 0.f) || (Gd != Gd) || (Gc != Gc);
    float A_carry = 0.f;
#pragma unroll
    for (int ch = NCH - 1; ch >= 0; --ch) {
        const int i = ch * G + gl;
        const bool in = EXACT || i < S;
        const float T = Tr[ch], alpha = al[ch];
        const float w = alpha * T;
        float g = Gr1 * sr[ch] + Gg1 * sg[ch] + Gb1 * sb[ch] + G0 + gwt[ch];
        if (depth_terms) {
            g += kd * ((t0[ch] + 0.5f * dist[ch]) - od);
            g += kc * ((t0[ch] + mu[ch] * dist[ch]) - oc);
        }
        const float gw = in ? g * w : 0.f;
        float suf = gscan_suffix<G>(gw, gl, m);
        const float A = A_carry + suf - gw;
        if (NCH > 1) A_carry += __shfl_sync(m, suf, 0, G);
        if (in && valid) {
            const float one_m_alpha = 1.0f - alpha;
            const float tm = one_m_alpha + 1e-10f;
            const float g_alpha = g * T - A * rcp_(tm);
            float g_rm = gm[ch];
            if (depth_terms) {
                const float wp = (a.blender && i == S - 1) ? w + 1e-10f : w;
                const float p = a.blender ? wp / W : wp;
                g_rm += Gc * p * dist[ch] * (mu[ch] - mu[ch] * mu[ch]);
            }
            float2* o = reinterpret_cast<float2*>(g_raw6 + (row0 + i) * 6);
            o[0] = make_float2(Gr1 * w * (sr[ch] - sr[ch] * sr[ch]), Gg1 * w * (sg[ch] - sg[ch] * sg[ch]));
            o[1] = make_float2(Gb1 * w * (sb[ch] - sb[ch] * sb[ch]), g_alpha * (dist[ch] * norm_d) * one_m_alpha * ds[ch]);
            o[2] = make_float2(g_rm, gs[ch]);
        }
    }
}

// (G, NCH) for S samples, measured (profiles/r01b_stage_roofline.md): four chunks per lane, short rays sharing a
// warp; the register-heavy backward prefers eight chunks from S = 128 up (two rays per warp amortise its
// per-ray epilogue; the forward is indifferent there).
template <typename F>
int dispatch_shape(int S, bool backward, F&& f) {
#define DDNERF_CASE(G, NCH)                                                                               \
    if (S <= G * NCH) {                                                                                   \
        if (S == G * NCH) return f(std::integral_constant<int, G>{}, std::integral_constant<int, NCH>{}, std::true_type{}); \
        return f(std::integral_constant<int, G>{}, std::integral_constant<int, NCH>{}, std::false_type{});                  \
    }
    DDNERF_CASE(4, 1) DDNERF_CASE(4, 2) DDNERF_CASE(4, 4) DDNERF_CASE(8, 4) DDNERF_CASE(16, 4)
    if (backward) { if (S > 64) { DDNERF_CASE(16, 8) } } else { DDNERF_CASE(32, 4) }
    DDNERF_CASE(32, 8) DDNERF_CASE(32, 16)
#undef DDNERF_CASE
    return -1;
}

template <typename F>
int dispatch_layout(int rawv, bool mu, F&& f) {
#define DDNERF_L(V)                                                                 \
    if (rawv == V) {                                                                \
        if (mu) return f(std::integral_constant<int, V>{}, std::true_type{});       \
        return f(std::integral_constant<int, V>{}, std::false_type{});              \
    }
    DDNERF_L(4) DDNERF_L(2) DDNERF_L(1)
#undef DDNERF_L
    return -1;
}

int raw_vector_width(const float* raw, int raw_stride) {
    const uintptr_t p = reinterpret_cast<uintptr_t>(raw);
    if (raw_stride == 4 && p % 16 == 0) return 4;
    if (raw_stride % 2 == 0 && p % 8 == 0) return 2;
    return 1;
}

}  // namespace
}  // namespace ddnerf

using namespace ddnerf;

extern "C" DDNERF_EXPORT int ddnerf_composite_forward(const float* raw, int raw_stride, const float* t, const float* rd,
                                        int64_t rd_stride, const float* noise, float noise_std, const float* mus,
                                        int white_background, int blender, float* rgb_map, float* disp, float* acc,
                                        float* weights, float* depth, float* cdisp, float* rgb, int64_t N, int S,
                                        void* stream) {
    DDNERF_CHECK_ARG(N == 0 || (raw && t && rd && rgb_map && disp && acc && weights && depth), "composite_forward: null pointer");
    DDNERF_CHECK_ARG(N == 0 || !mus || cdisp, "composite_forward: mus given but cdisp is null");
    DDNERF_CHECK_ARG(S >= 1 && S <= 512, "composite_forward: S=%d outside [1,512]", S);
    DDNERF_CHECK_ARG(raw_stride >= 4, "composite_forward: raw_stride=%d < 4", raw_stride);
    if (N == 0) return 0;
    const bool use_noise = noise_std > 0.f && noise;
    CompositeArgs a{raw, raw_stride, t, rd, rd_stride, use_noise ? noise : nullptr, use_noise ? noise_std : 0.f, mus,
                    white_background, blender, N, S};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = dispatch_shape(S, false, [&](auto g, auto nch, auto ex) {
        return dispatch_layout(raw_vector_width(raw, raw_stride), mus != nullptr, [&](auto v, auto mu) {
            constexpr int G = decltype(g)::value, NCH = decltype(nch)::value, V = decltype(v)::value;
            constexpr bool MU = decltype(mu)::value, EX = decltype(ex)::value;
            const int threads = NCH >= 16 ? 128 : 256, rays_per_block = threads / G;
            composite_fwd_kernel<G, NCH, V, MU, EX><<<ceil_div(N, rays_per_block), threads, 0, st>>>(
                a, rgb_map, disp, acc, weights, depth, cdisp, rgb);
            return 0;
        });
    });
    DDNERF_CHECK_ARG(rc == 0, "composite_forward: unsupported S=%d", S);
    DDNERF_LAUNCHED("composite_forward", 1);
    return 0;
}

extern "C" DDNERF_EXPORT int ddnerf_composite_backward(const float* raw, int raw_stride, const float* t, const float* rd,
                                         int64_t rd_stride, const float* noise, float noise_std, const float* mus,
                                         int white_background, int blender, const float* g_rgb_map, const float* g_disp,
                                         const float* g_acc, const float* g_weights, const float* g_depth,
                                         const float* g_cdisp, float* g_raw, float* g_mus, int64_t N, int S,
                                         void* stream) {
    DDNERF_CHECK_ARG(N == 0 || (raw && t && rd && g_raw), "composite_backward: null pointer");
    DDNERF_CHECK_ARG(S >= 1 && S <= 512, "composite_backward: S=%d outside [1,512]", S);
    DDNERF_CHECK_ARG(reinterpret_cast<uintptr_t>(g_raw) % 16 == 0, "composite_backward: g_raw not 16-byte aligned");
    if (N == 0) return 0;
    const bool use_noise = noise_std > 0.f && noise;
    CompositeArgs a{raw, raw_stride, t, rd, rd_stride, use_noise ? noise : nullptr, use_noise ? noise_std : 0.f, mus,
                    white_background, blender, N, S};
    CompositeGrads gr{g_rgb_map, g_disp, g_acc, g_weights, g_depth, g_cdisp};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = dispatch_shape(S, true, [&](auto g, auto nch, auto ex) {
        return dispatch_layout(raw_vector_width(raw, raw_stride), mus != nullptr, [&](auto v, auto mu) {
            constexpr int G = decltype(g)::value, NCH = decltype(nch)::value, V = decltype(v)::value;
            constexpr bool MU = decltype(mu)::value, EX = decltype(ex)::value;
            const int threads = NCH >= 8 ? 128 : 256, rays_per_block = threads / G;
            composite_bwd_kernel<G, NCH, V, MU, EX><<<ceil_div(N, rays_per_block), threads, 0, st>>>(a, gr, g_raw, g_mus);
            return 0;
        });
    });
    DDNERF_CHECK_ARG(rc == 0, "composite_backward: unsupported S=%d", S);
    DDNERF_LAUNCHED("composite_backward", 1);
    return 0;
}

/* Upper bound of the scratch floats ddnerf_composite_dd_forward needs for N rays (block partial sums + a counter). */
extern "C" DDNERF_EXPORT int64_t ddnerf_composite_dd_scratch_floats(int64_t N) { return N + 8; }

extern "C" DDNERF_EXPORT int ddnerf_composite_dd_forward(const float* raw6, const float* t, const float* rd, int64_t rd_stride,
                                           const float* noise, float noise_std, int white_background, int blender,
                                           float dist_reg_coef, float* rgb_map, float* disp, float* acc, float* weights,
                                           float* depth, float* cdisp, float* mus, float* sigmas, float* regs,
                                           float* scratch, int64_t N, int S, void* stream) {
    DDNERF_CHECK_ARG(regs && (N == 0 || (raw6 && t && rd && rgb_map && disp && acc && weights && depth && cdisp && mus && sigmas && scratch)),
                     "composite_dd_forward: null pointer");
    DDNERF_CHECK_ARG(S >= 1 && S <= 512, "composite_dd_forward: S=%d outside [1,512]", S);
    DDNERF_CHECK_ARG(reinterpret_cast<uintptr_t>(raw6) % 8 == 0, "composite_dd_forward: raw6 not 8-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (N == 0) { cudaMemsetAsync(regs, 0, 4 * sizeof(float), st); return 0; }
    const bool use_noise = noise_std > 0.f && noise;
    CompositeArgs a{raw6, 6, t, rd, rd_stride, use_noise ? noise : nullptr, use_noise ? noise_std : 0.f, nullptr,
                    white_background, blender, N, S};
    DDOut dd{mus, sigmas, regs, scratch, dist_reg_coef};
    int rc = dispatch_shape(S, false, [&](auto g, auto nch, auto ex) {
        constexpr int G = decltype(g)::value, NCH = decltype(nch)::value;
        constexpr bool EX = decltype(ex)::value;
        const int threads = NCH >= 16 ? 128 : 256, rays_per_block = threads / G;
        const int grid = ceil_div(N, rays_per_block);
        cudaMemsetAsync(scratch + 2 * (size_t)grid, 0, sizeof(unsigned), st);          // the arrival counter
        composite_dd_fwd_kernel<G, NCH, EX><<<grid, threads, 0, st>>>(a, dd, rgb_map, disp, acc, weights, depth, cdisp);
        return 0;
    });
    DDNERF_CHECK_ARG(rc == 0, "composite_dd_forward: unsupported S=%d", S);
    DDNERF_LAUNCHED("composite_dd_forward", 1);
    return 0;
}

extern "C" DDNERF_EXPORT int ddnerf_composite_dd_backward(const float* raw6, const float* t, const float* rd, int64_t rd_stride,
                                            const float* noise, float noise_std, int white_background, int blender,
                                            float dist_reg_coef, const float* g_rgb_map, const float* g_disp,
                                            const float* g_acc, const float* g_weights, const float* g_depth,
                                            const float* g_cdisp, const float* g_mus, const float* g_sigmas,
                                            const float* g_regs, float* g_raw6, int64_t N, int S, void* stream) {
    DDNERF_CHECK_ARG(N == 0 || (raw6 && t && rd && g_raw6), "composite_dd_backward: null pointer");
    DDNERF_CHECK_ARG(S >= 1 && S <= 512, "composite_dd_backward: S=%d outside [1,512]", S);
    DDNERF_CHECK_ARG(reinterpret_cast<uintptr_t>(raw6) % 8 == 0 && reinterpret_cast<uintptr_t>(g_raw6) % 8 == 0,
                     "composite_dd_backward: raw6 / g_raw6 not 8-byte aligned");
    if (N == 0) return 0;
    const bool use_noise = noise_std > 0.f && noise;
    CompositeArgs a{raw6, 6, t, rd, rd_stride, use_noise ? noise : nullptr, use_noise ? noise_std : 0.f, nullptr,
                    white_background, blender, N, S};
    CompositeGrads gr{g_rgb_map, g_disp, g_acc, g_weights, g_depth, g_cdisp};
    DDGrads dg{g_mus, g_sigmas, g_regs, dist_reg_coef};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = dispatch_shape(S, true, [&](auto g, auto nch, auto ex) {
        constexpr int G = decltype(g)::value, NCH = decltype(nch)::value;
        constexpr bool EX = decltype(ex)::value;
        const int threads = NCH >= 8 ? 128 : 256, rays_per_block = threads / G;
        composite_dd_bwd_kernel<G, NCH, EX><<<ceil_div(N, rays_per_block), threads, 0, st>>>(a, gr, dg, g_raw6);
        return 0;
    });
    DDNERF_CHECK_ARG(rc == 0, "composite_dd_backward: unsupported S=%d", S);
    DDNERF_LAUNCHED("composite_dd_backward", 1);
    return 0;
}
