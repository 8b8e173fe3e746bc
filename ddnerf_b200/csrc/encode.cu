// K2 (standalone): rays + fence-posts -> the 96 IPE and 27 view-direction features per sample.
//
// Replaces cast_rays -> integrated_pos_enc -> positional_encoding -> cat of run_network,
// models/models.py:117-133.  One thread per (sample row, degree l): 16 threads of a row write
// 48 contiguous floats of the sin block and 48 of the cos block; threads l < 9 also write one
// 3-wide group of the view-direction encoding.  Inputs are 48 B/ray + 4 B/sample; the output
// (492 B/sample fp32) exists only for callers that want the feature matrix in HBM -- the MLP
// kernels run the same device functions (encode.cuh) as their producer stage instead.
#include "encode.cuh"

namespace ddnerf {
namespace {

__global__ void __launch_bounds__(256) encode_kernel(const float* __restrict__ rays, const float* __restrict__ t_vals,
                                                      float* __restrict__ enc_out, int64_t ld_enc,
                                                      float* __restrict__ dir_out, int64_t ld_dir, int64_t N, int S,
                                                      int ray_shape) {
    const int l = threadIdx.x & 15;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 4) + (threadIdx.x >> 4);
    if (row >= N * S) return;
    const int64_t ray = row / S;
    const int i = (int)(row - ray * S);
    const RayGeom g = load_ray(rays, ray);
    const float* tp = t_vals + ray * (S + 1) + i;
    const Gauss3 s = cast_interval(g, __ldg(tp), __ldg(tp + 1), ray_shape);
    float sn[3], cs[3];
    ipe_degree(s, l, sn, cs);
    float* e = enc_out + row * ld_enc;
#pragma unroll
    for (int a = 0; a < 3; ++a) { e[l * 3 + a] = sn[a]; e[48 + l * 3 + a] = cs[a]; }
    if (dir_out && l < 9) {
        float d3[3];
        dir_group(g, l, d3);
        float* d = dir_out + row * ld_dir;
#pragma unroll
        for (int a = 0; a < 3; ++a) d[l * 3 + a] = d3[a];
    }
}

// get_rays_batches, models/models.py:144-158: viewdirs = d / ||d||_2, rays = cat(o, d, radius, near, far, viewdirs).
// One thread per ray instead of eight elementwise / reduce / cat launches; fp32 op by op (this file is built without FMA
// contraction): squares summed in x, y, z order, IEEE square root and divisions.
__global__ void __launch_bounds__(256) pack_rays_kernel(const float* __restrict__ ro, const float* __restrict__ rd,
                                                        const float* __restrict__ rad, float near, float far,
                                                        float* __restrict__ rays, int64_t N) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float ox = __ldg(ro + 3 * i), oy = __ldg(ro + 3 * i + 1), oz = __ldg(ro + 3 * i + 2);
    const float dx = __ldg(rd + 3 * i), dy = __ldg(rd + 3 * i + 1), dz = __ldg(rd + 3 * i + 2);
    const float nrm = __fsqrt_rn(dx * dx + dy * dy + dz * dz);
    float4* o = reinterpret_cast<float4*>(rays + 12 * i);
    o[0] = make_float4(ox, oy, oz, dx);
    o[1] = make_float4(dy, dz, __ldg(rad + i), near);
    o[2] = make_float4(far, __fdiv_rn(dx, nrm), __fdiv_rn(dy, nrm), __fdiv_rn(dz, nrm));
}

}  // namespace
}  // namespace ddnerf

using namespace ddnerf;

extern "C" DDNERF_EXPORT int ddnerf_pack_rays(const float* ray_origins, const float* ray_directions, const float* ray_radii,
                                              float near, float far, int64_t N, float* rays, void* stream) {
    DDNERF_CHECK_ARG(N == 0 || (ray_origins && ray_directions && ray_radii && rays), "pack_rays: null pointer");
    DDNERF_CHECK_ARG((reinterpret_cast<uintptr_t>(rays) & 15u) == 0, "pack_rays: rays must be 16-byte aligned");
    if (N == 0) return 0;
    pack_rays_kernel<<<ceil_div(N, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(ray_origins, ray_directions, ray_radii, near, far,
                                                                                      rays, N);
    DDNERF_LAUNCHED("pack_rays", 1);
    return 0;
}

extern "C" DDNERF_EXPORT int ddnerf_encode(const float* rays, const float* t_vals, float* enc_out, int64_t ld_enc, float* dir_out,
                             int64_t ld_dir, int64_t N, int S, int ray_shape, void* stream) {
    DDNERF_CHECK_ARG(N == 0 || S == 0 || (rays && t_vals && enc_out), "encode: null pointer");
    DDNERF_CHECK_ARG(ray_shape == 0 || ray_shape == 1, "encode: ray_shape=%d (0 cone, 1 cylinder)", ray_shape);
    DDNERF_CHECK_ARG(ld_enc >= 96 && (!dir_out || ld_dir >= 27), "encode: leading dimension too small");
    if (N == 0 || S == 0) return 0;
    encode_kernel<<<ceil_div(N * S, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(rays, t_vals, enc_out, ld_enc,
                                                                                       dir_out, ld_dir, N, S, ray_shape);
    DDNERF_LAUNCHED("encode", 1);
    return 0;
}
