// Device-side mip-NeRF featurisation shared by the standalone encoder (encode.cu) and the
// MLP kernels that fuse it as their producer stage.
//
// Restates general_utils/math_utils.py:57-88 (conical_frustum_to_gaussian), :91-110
// (cylinder_to_gaussian), :34-46 (lift_gaussian, diagonal), :112-166 (integrated_pos_enc with
// safe_sin) and general_utils/nerf_helpers.py:127-171 (positional_encoding) of the reference.
#pragma once
#include "common.cuh"

namespace ddnerf {

struct RayGeom {           // per-ray constants, rays[N,12] = (o3, d3, radius, near, far, viewdir3)
    float ox, oy, oz, dx, dy, dz, rad, vx, vy, vz;
};

__device__ __forceinline__ RayGeom load_ray(const float* __restrict__ rays, int64_t ray) {
    const float* r = rays + ray * 12;
    RayGeom g;
    g.ox = __ldg(r); g.oy = __ldg(r + 1); g.oz = __ldg(r + 2);
    g.dx = __ldg(r + 3); g.dy = __ldg(r + 4); g.dz = __ldg(r + 5);
    g.rad = __ldg(r + 6);
    g.vx = __ldg(r + 9); g.vy = __ldg(r + 10); g.vz = __ldg(r + 11);
    return g;
}

struct Gauss3 { float mx, my, mz, cx, cy, cz; };

// interval (t0,t1) of a ray -> diagonal Gaussian (mean, cov_diag).
// Written with the round-to-nearest intrinsics, which the compiler never contracts into FMAs: the encoding at level l
// multiplies the mean by 2^l, so ONE ulp of the mean is 2^l ulps of phase (5e-4 rad at l = 10) -- the reference's own
// features are that sensitive to its op-by-op fp32 rounding, and every compile unit (the fp32 encoder built with
// -fmad=false, the bf16 MLP kernels built with contraction on) must reproduce it bit for bit.
__device__ __forceinline__ Gauss3 cast_interval(const RayGeom& g, float t0, float t1, int ray_shape) {
    const auto mul = [](float a, float b) { return __fmul_rn(a, b); };
    const auto add = [](float a, float b) { return __fadd_rn(a, b); };
    const auto sub = [](float a, float b) { return __fsub_rn(a, b); };
    const auto div = [](float a, float b) { return __fdiv_rn(a, b); };
    float t_mean, t_var, r_var;
    const float rad2 = mul(g.rad, g.rad);
    if (ray_shape == 0) {                                   // cone, math_utils.py:76-82
        const float mu = div(add(t0, t1), 2.0f), hw = div(sub(t1, t0), 2.0f);
        const float mu2 = mul(mu, mu), hw2 = mul(hw, hw), hw4 = mul(hw2, hw2);
        const float den = add(mul(3.0f, mu2), hw2);
        t_mean = add(mu, div(mul(mul(2.0f, mu), hw2), den));
        t_var = sub(div(hw2, 3.0f), mul((float)(4.0 / 15.0), div(mul(hw4, sub(mul(12.0f, mu2), hw2)), mul(den, den))));
        r_var = mul(rad2, sub(add(div(mu2, 4.0f), mul((float)(5.0 / 12.0), hw2)), div(mul((float)(4.0 / 15.0), hw4), den)));
    } else {                                                // cylinder, math_utils.py:107-109
        t_mean = div(add(t0, t1), 2.0f);
        r_var = div(rad2, 4.0f);
        const float d = sub(t1, t0);
        t_var = div(mul(d, d), 12.0f);
    }
    const float dx2 = mul(g.dx, g.dx), dy2 = mul(g.dy, g.dy), dz2 = mul(g.dz, g.dz);
    const float dmag = fmaxf(1e-10f, add(add(dx2, dy2), dz2));            // math_utils.py:38
    Gauss3 o;
    o.mx = add(mul(g.dx, t_mean), g.ox); o.my = add(mul(g.dy, t_mean), g.oy); o.mz = add(mul(g.dz, t_mean), g.oz);
    o.cx = add(mul(t_var, dx2), mul(r_var, sub(1.0f, div(dx2, dmag))));
    o.cy = add(mul(t_var, dy2), mul(r_var, sub(1.0f, div(dy2, dmag))));
    o.cz = add(mul(t_var, dz2), mul(r_var, sub(1.0f, div(dz2, dmag))));
    return o;
}

// safe_sin's range reduction, math_utils.py:154-166: x if |x| < T else x % T (floored remainder,
// T = fl32(100*pi)).  floor(x/T) can only be one too large (when x/T rounds up to an integer), in
// which case x - q*T is a tiny negative number and +T is exact; the fma makes x - q*T exact, so the
// result equals torch's fmod-based remainder bit for bit for |x| < 2^18.
__device__ __forceinline__ float safe_arg(float x) {
    const float T = 314.15927124f;
    if (fabsf(x) < T) return x;
    float q = floorf(x / T);
    float r = fmaf(-q, T, x);
    if (r < 0.f) r += T;
    return r;
}

// the 6 IPE features of (degree l, axes xyz): sin block at [l*3 + a], cos block at [48 + l*3 + a]
__device__ __forceinline__ void ipe_degree(const Gauss3& s, int l, float* sin3, float* cos3) {
    const float scale = (float)(1 << l);
    const float sc2 = scale * scale;
    const float half_pi = 1.57079637f;                      // fl32(0.5*pi), math_utils.py:143
    const float m[3] = {s.mx, s.my, s.mz}, c[3] = {s.cx, s.cy, s.cz};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float y = m[a] * scale;
        float e = expf(-0.5f * (c[a] * sc2));
        sin3[a] = e * sinf(safe_arg(y));
        cos3[a] = e * sinf(safe_arg(y + half_pi));
    }
}

// group g in [0,9) of the 27-wide view-direction encoding: [v | sin v, cos v | sin 2v, cos 2v | ...]
__device__ __forceinline__ void dir_group(const RayGeom& r, int g, float* out3) {
    const float v[3] = {r.vx, r.vy, r.vz};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (g == 0) out3[a] = v[a];
        else {
            float x = v[a] * (float)(1 << ((g - 1) >> 1));
            out3[a] = (g & 1) ? sinf(x) : cosf(x);
        }
    }
}

}  // namespace ddnerf
