// f3: the training ray store on the device.
//
// Replaces data_utils/dataset.py:8-59 (TrainDataset) of the reference, which keeps every ray of every training
// image in host memory and, each iteration, draws indices with numpy, gathers four host tensors and copies them
// to the device.  Here the rays live in HBM as one packed row per ray -- {origin 3, direction 3, radius 1,
// target rgb 3, pad 2} = 48 bytes, 16-byte aligned, so a random row costs two 32-byte sectors instead of the four
// it would touch in four separate arrays -- and a batch is one gather kernel over device-resident indices.
#include "common.cuh"

namespace ddnerf {
namespace {

constexpr int kRow = 12;     // floats per packed ray

__global__ void __launch_bounds__(256) raystore_pack_kernel(const float* __restrict__ ro, const float* __restrict__ rd,
                                                             const float* __restrict__ rad, const float* __restrict__ tgt,
                                                             int64_t n, float4* __restrict__ rows) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* o = ro + i * 3; const float* d = rd + i * 3; const float* t = tgt + i * 3;
    rows[i * 3 + 0] = make_float4(__ldg(o), __ldg(o + 1), __ldg(o + 2), __ldg(d));
    rows[i * 3 + 1] = make_float4(__ldg(d + 1), __ldg(d + 2), __ldg(rad + i), __ldg(t));
    rows[i * 3 + 2] = make_float4(__ldg(t + 1), __ldg(t + 2), 0.f, 0.f);
}

__global__ void __launch_bounds__(256) raystore_gather_kernel(const float4* __restrict__ rows, const int64_t* __restrict__ idx,
                                                               int64_t n, int64_t base, int64_t total, float* __restrict__ ro,
                                                               float* __restrict__ rd, float* __restrict__ rad,
                                                               float* __restrict__ tgt, int* __restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t r = base + __ldg(idx + i);
    if (r < 0 || r >= total) { if (bad) atomicExch(bad, 1); r = 0; }     // out-of-range index: flagged, row 0 returned
    const float4 a = __ldg(rows + r * 3), b = __ldg(rows + r * 3 + 1), c = __ldg(rows + r * 3 + 2);
    ro[i * 3] = a.x; ro[i * 3 + 1] = a.y; ro[i * 3 + 2] = a.z;
    rd[i * 3] = a.w; rd[i * 3 + 1] = b.x; rd[i * 3 + 2] = b.y;
    rad[i] = b.z;
    tgt[i * 3] = b.w; tgt[i * 3 + 1] = c.x; tgt[i * 3 + 2] = c.y;
}

}  // namespace
}  // namespace ddnerf

using namespace ddnerf;

extern "C" DDNERF_EXPORT int ddnerf_raystore_pack(const float* ray_origins, const float* ray_directions, const float* radii,
                                                  const float* target_rgb, int64_t n, float* rows, void* stream) {
    DDNERF_CHECK_ARG(ray_origins && ray_directions && radii && target_rgb && rows, "raystore_pack: null pointer");
    DDNERF_CHECK_ARG(reinterpret_cast<uintptr_t>(rows) % 16 == 0, "raystore_pack: rows not 16-byte aligned");
    if (n <= 0) return 0;
    raystore_pack_kernel<<<ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        ray_origins, ray_directions, radii, target_rgb, n, reinterpret_cast<float4*>(rows));
    DDNERF_LAUNCHED("raystore_pack", 1);
    return 0;
}

extern "C" DDNERF_EXPORT int ddnerf_raystore_gather(const float* rows, int64_t total_rows, const int64_t* idx, int64_t n,
                                                    int64_t base, float* ray_origins, float* ray_directions, float* radii,
                                                    float* target_rgb, int* bad_index_flag, void* stream) {
    DDNERF_CHECK_ARG(rows && idx && ray_origins && ray_directions && radii && target_rgb, "raystore_gather: null pointer");
    DDNERF_CHECK_ARG(reinterpret_cast<uintptr_t>(rows) % 16 == 0, "raystore_gather: rows not 16-byte aligned");
    if (n <= 0) return 0;
    raystore_gather_kernel<<<ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(rows), idx, n, base, total_rows, ray_origins, ray_directions, radii, target_rgb,
        bad_index_flag);
    DDNERF_LAUNCHED("raystore_gather", 1);
    return 0;
}
