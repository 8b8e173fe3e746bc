// f1 (first "next" row): per-pixel ray generation on the device.
//
// Replaces general_utils/nerf_helpers.py:67-125 (get_ray_bundle) and, for forward-facing scenes,
// data_utils/dataset_helpers.py:3-42 (ndc_mipnerf_rays) of the reference: once the per-ray path renders a
// 1008 x 756 frame in tens of milliseconds, building the rays with a CPU meshgrid and copying 21 MB of them
// to the device every frame is the render loop's bottleneck.  Here a frame needs the 12 floats of its pose.
// One thread per pixel, 28 bytes written per ray (origin 3, direction 3, radius 1), nothing read; the
// neighbour rays the radii need are recomputed in the thread.  Arithmetic follows the reference's fp32
// operation order (compiled with -fmad=false).
#include "common.cuh"

namespace ddnerf {
namespace {

struct RayGenArgs {
    float c2w[12];          // rows of the 3x4 camera-to-world matrix
    const float* c2w_dev;   // if set, the pose is read from device memory instead (CUDA-graph replay with a new pose)
    int H, W; float focal;
    int ndc; float ndc_near;
    int row_lo, row_hi;
    // host-computed scalars the reference evaluates in double and then casts (dataset_helpers.py:12-27)
    float sx, sy, two_near;
};

struct Ray { float o[3], d[3], dir_cam[3]; };

__device__ __forceinline__ Ray camera_ray(const RayGenArgs& a, int j, int i) {
    Ray r;
    r.dir_cam[0] = ((float)i - (float)a.W * 0.5f) / a.focal;            // nerf_helpers.py:101-108
    r.dir_cam[1] = -((float)j - (float)a.H * 0.5f) / a.focal;
    r.dir_cam[2] = -1.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        r.d[k] = (r.dir_cam[0] * a.c2w[k * 4 + 0] + r.dir_cam[1] * a.c2w[k * 4 + 1]) + r.dir_cam[2] * a.c2w[k * 4 + 2];
        r.o[k] = a.c2w[k * 4 + 3];
        if (r.o[k] == 0.f) r.o[k] += 1e-5f;                              // :114-115
        if (r.d[k] == 0.f) r.d[k] += 1e-5f;
    }
    return r;
}

// dataset_helpers.py:8-30, same operation order
__device__ __forceinline__ void ndc_project(const RayGenArgs& a, Ray& r) {
    const float t = -(a.ndc_near + r.o[2]) / r.d[2];
    float o[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) o[k] = r.o[k] + t * r.d[k];
    const float o0 = a.sx * o[0] / o[2];
    const float o1 = a.sy * o[1] / o[2];
    const float o2 = 1.0f + a.two_near / o[2];
    const float d0 = a.sx * (r.d[0] / r.d[2] - o[0] / o[2]);
    const float d1 = a.sy * (r.d[1] / r.d[2] - o[1] / o[2]);
    const float d2 = -a.two_near / o[2];
    r.o[0] = o0; r.o[1] = o1; r.o[2] = o2;
    r.d[0] = d0; r.d[1] = d1; r.d[2] = d2;
}

__device__ __forceinline__ float dist3(const float* p, const float* q) {
    const float x = p[0] - q[0], y = p[1] - q[1], z = p[2] - q[2];
    return sqrtf((x * x + y * y) + z * z);
}

__global__ void __launch_bounds__(256) raygen_kernel(RayGenArgs a, float* __restrict__ ro, float* __restrict__ rd,
                                                      float* __restrict__ rad) {
    if (a.c2w_dev) {
#pragma unroll
        for (int k = 0; k < 12; ++k) a.c2w[k] = __ldg(a.c2w_dev + k);
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = a.row_lo + blockIdx.y;
    if (i >= a.W) return;
    Ray r = camera_ray(a, j, i);
    float radius;
    if (!a.ndc) {
        // radii from the camera-space directions of vertically adjacent pixels; the last row takes dx[-2:-1], i.e.
        // the distance between rows H-3 and H-2 (nerf_helpers.py:119)
        const int j0 = j < a.H - 1 ? j : a.H - 3;
        const Ray p = camera_ray(a, j0, i), q = camera_ray(a, j0 + 1, i);
        radius = dist3(p.dir_cam, q.dir_cam) * 2.0f / 3.46410161513775f;                 // :117-123
    } else {
        ndc_project(a, r);
        const int j0 = j < a.H - 1 ? j : a.H - 3, i0 = i < a.W - 1 ? i : a.W - 3;     // dx[-2:-1], dy[:, -2:-1]
        Ray p = camera_ray(a, j0, i), q = camera_ray(a, j0 + 1, i);
        ndc_project(a, p); ndc_project(a, q);
        const float dx = dist3(p.o, q.o);
        p = camera_ray(a, j, i0); q = camera_ray(a, j, i0 + 1);
        ndc_project(a, p); ndc_project(a, q);
        const float dy = dist3(p.o, q.o);
        radius = (0.5f * (dx + dy)) * 2.0f / 3.46410161513775f;                          // dataset_helpers.py:33-40
    }
    const int64_t px = (int64_t)(j - a.row_lo) * a.W + i;
    ro[px * 3] = r.o[0]; ro[px * 3 + 1] = r.o[1]; ro[px * 3 + 2] = r.o[2];
    rd[px * 3] = r.d[0]; rd[px * 3 + 1] = r.d[1]; rd[px * 3 + 2] = r.d[2];
    rad[px] = radius;
}

}  // namespace
}  // namespace ddnerf

using namespace ddnerf;

static int launch_ray_bundle(const char* name, int H, int W, double focal, const float* c2w_host, const float* c2w_dev,
                             int ndc, float ndc_near, int row_lo, int row_hi, float* ray_origins, float* ray_directions,
                             float* radii, void* stream) {
    DDNERF_CHECK_ARG((c2w_host || c2w_dev) && ray_origins && ray_directions && radii, "%s: null pointer", name);
    DDNERF_CHECK_ARG(H >= 3 && W >= 3 && focal > 0.0, "%s: H=%d W=%d focal=%g unsupported", name, H, W, focal);
    DDNERF_CHECK_ARG(0 <= row_lo && row_lo <= row_hi && row_hi <= H, "%s: rows [%d,%d) outside [0,%d)", name, row_lo, row_hi, H);
    if (row_hi == row_lo) return 0;
    RayGenArgs a;
    for (int k = 0; k < 12; ++k) a.c2w[k] = c2w_host ? c2w_host[k] : 0.f;
    a.c2w_dev = c2w_dev;
    a.H = H; a.W = W; a.focal = (float)focal; a.ndc = ndc; a.ndc_near = ndc_near; a.row_lo = row_lo; a.row_hi = row_hi;
    a.sx = (float)(-1.0 / ((double)W / (2.0 * focal)));
    a.sy = (float)(-1.0 / ((double)H / (2.0 * focal)));
    a.two_near = (float)(2.0 * (double)ndc_near);
    dim3 grid(ceil_div(W, 256), row_hi - row_lo);
    raygen_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, ray_origins, ray_directions, radii);
    DDNERF_LAUNCHED(name, 1);
    return 0;
}

extern "C" DDNERF_EXPORT int ddnerf_ray_bundle(int H, int W, double focal, const float* c2w_host, int ndc, float ndc_near,
                                               int row_lo, int row_hi, float* ray_origins, float* ray_directions,
                                               float* radii, void* stream) {
    return launch_ray_bundle("ray_bundle", H, W, focal, c2w_host, nullptr, ndc, ndc_near, row_lo, row_hi, ray_origins,
                             ray_directions, radii, stream);
}

extern "C" DDNERF_EXPORT int ddnerf_ray_bundle_dev(int H, int W, double focal, const float* c2w_dev, int ndc, float ndc_near,
                                                   int row_lo, int row_hi, float* ray_origins, float* ray_directions,
                                                   float* radii, void* stream) {
    return launch_ray_bundle("ray_bundle_dev", H, W, focal, nullptr, c2w_dev, ndc, ndc_near, row_lo, row_hi, ray_origins,
                             ray_directions, radii, stream);
}
