// f4: frame post-processing on the device.
//
// Replaces validation_utils/visualization.py:11-27 (cast_to_disparity_image, cast_to_image) and the frame assembly
// of render_video.py:96-101 of the reference, which pull the float frame to the host (16 bytes per pixel) and
// convert there.  Here the 8-bit images are produced next to the renderer and 4 (or 6, with the video frame)
// bytes per pixel cross to the host.  Two launches per frame: the disparity range (one pass, ordered-integer
// atomics, last block decodes and re-arms the workspace), then the packing pass.
#include "common.cuh"

namespace ddnerf {
namespace {

// order-preserving map float -> uint32 (NaNs excluded by the callers)
__device__ __forceinline__ unsigned fkey(float f) {
    unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float funkey(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ws[0] = max over ~key (i.e. the minimum), ws[1] = max over key, ws[2] = ticket; all zero between frames.
__global__ void __launch_bounds__(256) frame_minmax_kernel(const float* __restrict__ disp, int64_t n, unsigned* ws,
                                                            float* __restrict__ minmax) {
    unsigned kmin = 0u, kmax = 0u;                                  // identities of the two max-reductions
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = __ldg(disp + i);
        if (v == v) { const unsigned k = fkey(v); kmin = max(kmin, ~k); kmax = max(kmax, k); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kmin = max(kmin, __shfl_xor_sync(FULL, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(FULL, kmax, o));
    }
    __shared__ unsigned s[2][8];
    if ((threadIdx.x & 31) == 0) { s[0][threadIdx.x >> 5] = kmin; s[1][threadIdx.x >> 5] = kmax; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { kmin = max(kmin, s[0][w]); kmax = max(kmax, s[1][w]); }
        atomicMax(ws + 0, kmin);
        atomicMax(ws + 1, kmax);
        __threadfence();
        if (atomicAdd(ws + 2, 1u) == gridDim.x - 1) {               // last block: decode, re-arm
            __threadfence();
            const unsigned a = atomicExch(ws + 0, 0u), b = atomicExch(ws + 1, 0u);
            minmax[0] = funkey(~a);
            minmax[1] = funkey(b);
            ws[2] = 0u;
        }
    }
}

__device__ __forceinline__ unsigned char to_u8(float x255) {          // .byte() / astype(uint8): truncation, clamped
    return (unsigned char)(int)fminf(fmaxf(x255, 0.f), 255.f);
}

__global__ void __launch_bounds__(256) frame_pack_kernel(const float* __restrict__ rgb, const float* __restrict__ disp,
                                                          const float* __restrict__ minmax, unsigned char* __restrict__ rgb8,
                                                          unsigned char* __restrict__ disp8, unsigned char* __restrict__ video,
                                                          int rows, int W) {
    const int64_t px = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (px >= (int64_t)rows * W) return;
    const float mn = __ldg(minmax), mx = __ldg(minmax + 1);
    const float r = __ldg(rgb + px * 3), g = __ldg(rgb + px * 3 + 1), b = __ldg(rgb + px * 3 + 2);
    const unsigned char r8 = to_u8(r * 255.f), g8 = to_u8(g * 255.f), b8 = to_u8(b * 255.f);
    float d = (__ldg(disp + px) - mn) / (mx - mn);                    // visualization.py:12
    d = fminf(fmaxf(d, 0.f), 1.f) * 255.f;                            // :15 (NaN -> 0 through fmaxf)
    const unsigned char d8 = to_u8(d);
    if (rgb8) { rgb8[px * 3] = r8; rgb8[px * 3 + 1] = g8; rgb8[px * 3 + 2] = b8; }
    if (disp8) disp8[px] = d8;
    if (video) {                                                      // render_video.py:96-101: [rgb | disparity], BGR
        const int64_t row = px / W, col = px - row * W;
        unsigned char* v = video + (row * 2 * W + col) * 3;
        v[0] = b8; v[1] = g8; v[2] = r8;
        v += (int64_t)W * 3;
        v[0] = d8; v[1] = d8; v[2] = d8;
    }
}

}  // namespace
}  // namespace ddnerf

using namespace ddnerf;

extern "C" DDNERF_EXPORT int ddnerf_frame_minmax(const float* disp, int64_t n, void* workspace, float* minmax, void* stream) {
    DDNERF_CHECK_ARG(disp && workspace && minmax, "frame_minmax: null pointer");
    DDNERF_CHECK_ARG(n >= 1, "frame_minmax: empty frame");
    const int blocks = (int)((n + 256 * 8 - 1) / (256 * 8));
    frame_minmax_kernel<<<blocks > 1184 ? 1184 : blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        disp, n, static_cast<unsigned*>(workspace), minmax);
    DDNERF_LAUNCHED("frame_minmax", 1);
    return 0;
}

extern "C" DDNERF_EXPORT int ddnerf_frame_pack_u8(const float* rgb, const float* disp, const float* minmax, uint8_t* rgb8,
                                                  uint8_t* disp8, uint8_t* video_bgr, int rows, int W, void* stream) {
    DDNERF_CHECK_ARG(rgb && disp && minmax, "frame_pack_u8: null pointer");
    DDNERF_CHECK_ARG(rgb8 || disp8 || video_bgr, "frame_pack_u8: no output requested");
    if (rows <= 0 || W <= 0) return 0;
    frame_pack_kernel<<<ceil_div((int64_t)rows * W, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        rgb, disp, minmax, rgb8, disp8, video_bgr, rows, W);
    DDNERF_LAUNCHED("frame_pack_u8", 1);
    return 0;
}
