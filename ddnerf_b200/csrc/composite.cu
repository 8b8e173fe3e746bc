// K4 / K4b: alpha compositing forward and analytic backward, one lane-group per ray.
//
// Replaces general_utils/volume_rendering_utils.py:6-84 (+ cumprod_exclusive,
// nerf_helpers.py:43-64) of the reference.  HBM-bound: every sample is read once (raw rgb/density,
// fence-posts, noise, mu) and its weight written once; the transmittance product and all per-ray
// sums are lane-group scans/reductions in registers.  A ray is owned by G = 8/16/32 lanes
// (smallest power of two >= S, capped at 32) and walked in NCH chunks of G samples, so adjacent
// lanes read adjacent samples (coalesced 16 B/lane when raw is [N,S,4]).
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

struct Sample {
    float r, g, b, dens, delta, dist, mid, spos, alpha, tm, mu;
};

template <bool VEC4>
__device__ __forceinline__ Sample load_sample(const CompositeArgs& a, int64_t ray, int i, bool in, float norm_d) {
    Sample s;
    s.r = s.g = s.b = s.dens = 0.f; s.delta = 0.f; s.dist = 0.f; s.mid = 0.f; s.spos = 0.f; s.alpha = 0.f; s.tm = 1.f; s.mu = 0.f;
    if (!in) return s;
    const float* tp = a.t + ray * (a.S + 1) + i;
    float t0 = __ldg(tp), t1 = __ldg(tp + 1);
    int64_t row = ray * a.S + i;
    if (VEC4) {
        float4 v = __ldg(reinterpret_cast<const float4*>(a.raw) + row);
        s.r = v.x; s.g = v.y; s.b = v.z; s.dens = v.w;
    } else {
        const float* rp = a.raw + row * a.raw_stride;
        s.r = __ldg(rp); s.g = __ldg(rp + 1); s.b = __ldg(rp + 2); s.dens = __ldg(rp + 3);
    }
    if (a.noise) s.dens += __ldg(a.noise + row) * a.noise_std;
    float dist = t1 - t0;
    s.dist = dist;
    s.delta = dist * norm_d;
    s.mid = (t1 + t0) / 2.0f;
    if (a.mus) { s.mu = __ldg(a.mus + row); s.spos = t0 + s.mu * dist; }
    float sa = softplusf_(s.dens - 1.0f);
    s.alpha = 1.0f - expf(-sa * s.delta);
    s.tm = 1.0f - s.alpha + 1e-10f;
    return s;
}

__device__ __forceinline__ float ray_norm(const CompositeArgs& a, int64_t ray, bool valid) {
    if (!valid) return 0.f;
    const float* d = a.rd + ray * a.rd_stride;
    float x = __ldg(d), y = __ldg(d + 1), z = __ldg(d + 2);
    return sqrtf(x * x + y * y + z * z);
}

template <int G, int NCH, bool VEC4>
__global__ void __launch_bounds__(256) composite_fwd_kernel(CompositeArgs a, float* __restrict__ rgb_map,
                                                             float* __restrict__ disp, float* __restrict__ acc,
                                                             float* __restrict__ weights, float* __restrict__ depth,
                                                             float* __restrict__ cdisp, float* __restrict__ rgb) {
    const int gl = threadIdx.x % G;
    const int64_t ray = (int64_t)blockIdx.x * (blockDim.x / G) + threadIdx.x / G;
    const bool valid = ray < a.N;
    const float norm_d = ray_norm(a, ray, valid);
    float T_carry = 1.f, s_w = 0.f, s_wm = 0.f, s_ws = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        const int i = ch * G + gl;
        const bool in = valid && i < a.S;
        Sample s = load_sample<VEC4>(a, ray, i, in, norm_d);
        float incl = group_incl_prod<G>(s.tm, gl);
        float excl = __shfl_up_sync(FULL, incl, 1, G);
        if (gl == 0) excl = 1.f;
        float T = T_carry * excl;
        T_carry *= __shfl_sync(FULL, incl, G - 1, G);
        float w = s.alpha * T;
        float r = sigmoidf_(s.r) * 1.002f - 0.001f;
        float g = sigmoidf_(s.g) * 1.002f - 0.001f;
        float b = sigmoidf_(s.b) * 1.002f - 0.001f;
        cr += w * r; cg += w * g; cb += w * b;                 // rgb_map uses w BEFORE the eps
        float wp = w;
        if (a.blender && i == a.S - 1) wp = w + 1e-10f;        // volume_rendering_utils.py:52-56
        if (in) {
            int64_t row = ray * a.S + i;
            weights[row] = wp;
            if (rgb) { rgb[row * 3] = r; rgb[row * 3 + 1] = g; rgb[row * 3 + 2] = b; }
            s_w += wp; s_wm += wp * s.mid; s_ws += wp * s.spos;
        }
    }
    s_w = group_sum<G>(s_w); s_wm = group_sum<G>(s_wm); s_ws = group_sum<G>(s_ws);
    cr = group_sum<G>(cr); cg = group_sum<G>(cg); cb = group_sum<G>(cb);
    if (valid && gl == 0) {
        float W = s_w;
        float d0 = a.blender ? s_wm / W : s_wm;
        float dsp = 1.0f / torch_max_(1e-10f, d0 / W);
        if (a.white) { cr += 1.0f - W; cg += 1.0f - W; cb += 1.0f - W; }
        rgb_map[ray * 3] = cr; rgb_map[ray * 3 + 1] = cg; rgb_map[ray * 3 + 2] = cb;
        acc[ray] = W; disp[ray] = dsp;
        if (a.mus) {
            float cd = a.blender ? s_ws / W : s_ws;
            cdisp[ray] = 1.0f / torch_max_(1e-10f, cd / W);
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

template <int G, int NCH, bool VEC4>
__global__ void __launch_bounds__(256) composite_bwd_kernel(CompositeArgs a, CompositeGrads gr,
                                                             float* __restrict__ g_raw, float* __restrict__ g_mus) {
    const int gl = threadIdx.x % G;
    const int64_t ray = (int64_t)blockIdx.x * (blockDim.x / G) + threadIdx.x / G;
    const bool valid = ray < a.N;
    const float norm_d = ray_norm(a, ray, valid);
    float alpha[NCH], Tr[NCH];
    float T_carry = 1.f, s_w = 0.f, s_wm = 0.f, s_ws = 0.f;
    // phase 1: recompute the forward scan (alpha_i, T_i) and the per-ray sums
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        const int i = ch * G + gl;
        const bool in = valid && i < a.S;
        Sample s = load_sample<VEC4>(a, ray, i, in, norm_d);
        float incl = group_incl_prod<G>(s.tm, gl);
        float excl = __shfl_up_sync(FULL, incl, 1, G);
        if (gl == 0) excl = 1.f;
        float T = T_carry * excl;
        T_carry *= __shfl_sync(FULL, incl, G - 1, G);
        alpha[ch] = s.alpha; Tr[ch] = T;
        float wp = s.alpha * T;
        if (a.blender && i == a.S - 1) wp += 1e-10f;
        if (in) { s_w += wp; s_wm += wp * s.mid; s_ws += wp * s.spos; }
    }
    s_w = group_sum<G>(s_w); s_wm = group_sum<G>(s_wm); s_ws = group_sum<G>(s_ws);
    // per-ray cotangents folded to (Ga on acc, Gd on depth0, Gc on corrected depth)
    float Gr = 0.f, Gg = 0.f, Gb = 0.f, Ga = 0.f, Gd = 0.f, Gc = 0.f;
    const float W = s_w;
    const float d0 = a.blender ? s_wm / W : s_wm;
    const float cd = a.blender ? s_ws / W : s_ws;
    if (valid) {
        if (gr.g_rgb_map) { Gr = __ldg(gr.g_rgb_map + ray * 3); Gg = __ldg(gr.g_rgb_map + ray * 3 + 1); Gb = __ldg(gr.g_rgb_map + ray * 3 + 2); }
        if (gr.g_acc) Ga = __ldg(gr.g_acc + ray);
        if (a.white) Ga -= Gr + Gg + Gb;
        float gdep = gr.g_depth ? __ldg(gr.g_depth + ray) : 0.f;
        if (a.mus) Gc = gdep; else Gd = gdep;
        if (gr.g_disp) {                                       // NaN (0/0, empty non-blender ray) propagates like torch
            float x = d0 / W;
            if (x > 1e-10f || x != x) { float k = -__ldg(gr.g_disp + ray) / (x * x); Gd += k / W; Ga += -k * d0 / (W * W); }
        }
        if (a.mus && gr.g_cdisp) {
            float x = cd / W;
            if (x > 1e-10f || x != x) { float k = -__ldg(gr.g_cdisp + ray) / (x * x); Gc += k / W; Ga += -k * cd / (W * W); }
        }
    }
    // phase 2: walk the chunks backwards with the suffix sum A_i = sum_{k>i} g_k w_k
    float A_carry = 0.f;
#pragma unroll
    for (int ch = NCH - 1; ch >= 0; --ch) {
        const int i = ch * G + gl;
        const bool in = valid && i < a.S;
        Sample s = load_sample<VEC4>(a, ray, i, in, norm_d);
        const float T = Tr[ch], al = alpha[ch];
        const float w = al * T;
        float sr = sigmoidf_(s.r), sg = sigmoidf_(s.g), sb = sigmoidf_(s.b);
        float g = 0.f;
        if (in) {
            int64_t row = ray * a.S + i;
            g = Gr * (sr * 1.002f - 0.001f) + Gg * (sg * 1.002f - 0.001f) + Gb * (sb * 1.002f - 0.001f) + Ga;
            if (gr.g_weights) g += __ldg(gr.g_weights + row);
            if (a.blender) g += Gd * (s.mid - d0) / W + Gc * (s.spos - cd) / W;
            else g += Gd * s.mid + Gc * s.spos;
        }
        float gw = g * w;
        float suf = group_suffix_sum<G>(gw, gl);
        float A = A_carry + suf - gw;
        A_carry += __shfl_sync(FULL, suf, 0, G);
        if (in) {
            int64_t row = ray * a.S + i;
            float g_alpha = g * T - A / s.tm;
            float x = s.dens - 1.0f;
            float sa = softplusf_(x);
            float dsp = x > 20.0f ? 1.0f : sigmoidf_(x);
            float g_dens = g_alpha * s.delta * expf(-sa * s.delta) * dsp;
            float4 o;
            o.x = Gr * w * 1.002f * sr * (1.0f - sr);
            o.y = Gg * w * 1.002f * sg * (1.0f - sg);
            o.z = Gb * w * 1.002f * sb * (1.0f - sb);
            o.w = g_dens;
            reinterpret_cast<float4*>(g_raw)[row] = o;
            if (g_mus) {
                float wp = (a.blender && i == a.S - 1) ? w + 1e-10f : w;
                float p = a.blender ? wp / W : wp;
                g_mus[row] = Gc * p * s.dist;                 // d cdepth / d mu_i = p_i * (t_{i+1}-t_i)
            }
        }
    }
}

template <typename F>
int dispatch_shape(int S, bool vec4, F&& f) {
#define DDNERF_CASE(G, NCH)                                          \
    if (S <= G * NCH) {                                              \
        if (vec4) return f(std::integral_constant<int, G>{}, std::integral_constant<int, NCH>{}, std::true_type{}); \
        return f(std::integral_constant<int, G>{}, std::integral_constant<int, NCH>{}, std::false_type{});          \
    }
    DDNERF_CASE(8, 1) DDNERF_CASE(16, 1) DDNERF_CASE(32, 1) DDNERF_CASE(32, 2) DDNERF_CASE(32, 4)
    DDNERF_CASE(32, 8) DDNERF_CASE(32, 16)
#undef DDNERF_CASE
    return -1;
}

}  // namespace
}  // namespace ddnerf

using namespace ddnerf;

extern "C" DDNERF_EXPORT int ddnerf_composite_forward(const float* raw, int raw_stride, const float* t, const float* rd,
                                        int64_t rd_stride, const float* noise, float noise_std, const float* mus,
                                        int white_background, int blender, float* rgb_map, float* disp, float* acc,
                                        float* weights, float* depth, float* cdisp, float* rgb, int64_t N, int S,
                                        void* stream) {
    DDNERF_CHECK_ARG(raw && t && rd && rgb_map && disp && acc && weights && depth, "composite_forward: null pointer");
    DDNERF_CHECK_ARG(!mus || cdisp, "composite_forward: mus given but cdisp is null");
    DDNERF_CHECK_ARG(S >= 1 && S <= 512, "composite_forward: S=%d outside [1,512]", S);
    DDNERF_CHECK_ARG(raw_stride >= 4, "composite_forward: raw_stride=%d < 4", raw_stride);
    if (N == 0) return 0;
    CompositeArgs a{raw, raw_stride, t, rd, rd_stride, noise_std > 0.f ? noise : nullptr, noise_std, mus,
                    white_background, blender, N, S};
    bool vec4 = raw_stride == 4 && (reinterpret_cast<uintptr_t>(raw) % 16 == 0);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = dispatch_shape(S, vec4, [&](auto g, auto nch, auto v) {
        constexpr int G = decltype(g)::value, NCH = decltype(nch)::value;
        constexpr bool V = decltype(v)::value;
        const int threads = 256, rays_per_block = threads / G;
        composite_fwd_kernel<G, NCH, V><<<ceil_div(N, rays_per_block), threads, 0, st>>>(a, rgb_map, disp, acc, weights,
                                                                                          depth, cdisp, rgb);
        return 0;
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
    DDNERF_CHECK_ARG(raw && t && rd && g_raw, "composite_backward: null pointer");
    DDNERF_CHECK_ARG(S >= 1 && S <= 512, "composite_backward: S=%d outside [1,512]", S);
    DDNERF_CHECK_ARG(reinterpret_cast<uintptr_t>(g_raw) % 16 == 0, "composite_backward: g_raw not 16-byte aligned");
    if (N == 0) return 0;
    CompositeArgs a{raw, raw_stride, t, rd, rd_stride, noise_std > 0.f ? noise : nullptr, noise_std, mus,
                    white_background, blender, N, S};
    CompositeGrads gr{g_rgb_map, g_disp, g_acc, g_weights, g_depth, g_cdisp};
    bool vec4 = raw_stride == 4 && (reinterpret_cast<uintptr_t>(raw) % 16 == 0);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = dispatch_shape(S, vec4, [&](auto g, auto nch, auto v) {
        constexpr int G = decltype(g)::value, NCH = decltype(nch)::value;
        constexpr bool V = decltype(v)::value;
        const int threads = 256, rays_per_block = threads / G;
        composite_bwd_kernel<G, NCH, V><<<ceil_div(N, rays_per_block), threads, 0, st>>>(a, gr, g_raw,
                                                                                          mus ? g_mus : nullptr);
        return 0;
    });
    DDNERF_CHECK_ARG(rc == 0, "composite_backward: unsupported S=%d", S);
    DDNERF_LAUNCHED("composite_backward", 1);
    return 0;
}
