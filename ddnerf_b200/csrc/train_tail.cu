// Training-step tail on device: photometric loss + its gradient, and Adam on the flat bucket.
//
// Replaces train_model.py:156-167 (loss = sum_j coef_j * mse(rgb_j, target)) and :175-177
// (torch.optim.Adam.step over 26 + 24 separate tensors) of the reference.  SURVEY.md section 8f
// row f2: once the per-ray path is fused, dozens of tiny optimizer launches dominate the step,
// so parameters live in one flat fp32 bucket that one kernel updates (and one all-reduce sums).
#include <algorithm>
#include <math.h>

#include "common.cuh"

namespace ddnerf {
namespace {

__global__ void mse_kernel(const float* __restrict__ rgb0, const float* __restrict__ rgb1, const float* __restrict__ target,
                           float coef0, float coef1, float* __restrict__ g0, float* __restrict__ g1,
                           float* __restrict__ mse_out, int64_t n) {
    __shared__ float red[2][8];
    float s0 = 0.f, s1 = 0.f;
    const float inv = 1.0f / (float)n;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        float tg = __ldg(target + e);
        float d0 = __ldg(rgb0 + e) - tg;
        s0 += d0 * d0;
        if (g0) g0[e] = coef0 * 2.0f * d0 * inv;
        if (rgb1) {
            float d1 = __ldg(rgb1 + e) - tg;
            s1 += d1 * d1;
            if (g1) g1[e] = coef1 * 2.0f * d1 * inv;
        }
    }
    s0 = group_sum<32>(s0); s1 = group_sum<32>(s1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = s0; red[1][warp] = s1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += red[0][w]; b += red[1][w]; }
        atomicAdd(mse_out, a * inv);
        atomicAdd(mse_out + 1, b * inv);
        atomicAdd(mse_out + 2, coef0 * (a * inv) + coef1 * (b * inv));     // the weighted photometric loss, train_model.py:163
    }
}

// torch.optim.Adam rounds its python-double hyper-parameters once: the lerp weight is fl32(1 - beta1) and the addcmul value
// fl32(1 - beta2) (NOT 1 - fl32(beta): 1 - fl32(0.999) is off by 1.3e-5 relative, which shows in exp_avg_sq), so the
// complements arrive as their own arguments (omb1, omb2), computed in double on the host.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            int64_t n, float lr, float omb1, float b2, float omb2, float eps, float bc1, float bc2_sqrt,
                            float gscale) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        float gr = __ldg(g + e) * gscale;
        float mm = m[e] + (gr - m[e]) * omb1;                  // exp_avg.lerp_(grad, 1-beta1)
        float vv = v[e] * b2 + omb2 * gr * gr;                 // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1-beta2)
        m[e] = mm; v[e] = vv;
        float denom = sqrtf(vv) / bc2_sqrt + eps;
        p[e] -= (lr / bc1) * (mm / denom);
    }
}

// same update with the scalars read from device memory: lets a CUDA graph of the whole step be replayed while
// the learning rate and the bias corrections change from step to step
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                int64_t n, const float* __restrict__ hyper) {
    const float lr = __ldg(hyper), b2 = __ldg(hyper + 2), eps = __ldg(hyper + 3);
    const float bc1 = __ldg(hyper + 4), bc2_sqrt = __ldg(hyper + 5), gscale = __ldg(hyper + 6);
    const float omb1 = __ldg(hyper + 8), omb2 = __ldg(hyper + 9);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        float gr = __ldg(g + e) * gscale;
        float mm = m[e] + (gr - m[e]) * omb1;
        float vv = v[e] * b2 + omb2 * gr * gr;
        m[e] = mm; v[e] = vv;
        float denom = sqrtf(vv) / bc2_sqrt + eps;
        p[e] -= (lr / bc1) * (mm / denom);
    }
}

// The per-iteration scalars of the driver loop (train_model.py:135-150: annealed gaussian_smooth_factor, learning_rate_decay)
// and Adam's bias corrections, computed ON THE DEVICE from a device-resident iteration counter: a replayed CUDA graph of the
// training step then reads nothing the host mutates between replays (a pinned host buffer rewritten in place races with the
// copy node of an earlier, still queued replay).  One thread, double arithmetic (what the Python driver does), then the
// counters advance.  state = {iteration i, Adam step count t}; hyper = {lr, beta1, beta2, eps, 1-beta1^t, sqrt(1-beta2^t),
// grad_scale, gaussian_smooth_factor, 1-beta1, 1-beta2}.
struct SchedArgs {
    double lr_init, lr_final, max_steps, lr_delay_steps, lr_delay_mult;
    double beta1, beta2, eps, grad_scale;
    double smooth0, dsmooth, final_smooth, finnish_smooth;
};

__global__ void schedule_kernel(long long* __restrict__ state, float* __restrict__ hyper, const SchedArgs a) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const long long i = state[0], t = state[1] + 1;
    double delay = 1.0;
    if (a.lr_delay_steps > 0) {
        const double x = fmin(fmax((double)i / a.lr_delay_steps, 0.0), 1.0);
        delay = a.lr_delay_mult + (1.0 - a.lr_delay_mult) * sin(0.5 * 3.14159265358979323846 * x);
    }
    const double u = fmin(fmax((double)i / a.max_steps, 0.0), 1.0);
    const double lr = delay * exp(log(a.lr_init) * (1.0 - u) + log(a.lr_final) * u);
    hyper[0] = (float)lr;
    hyper[1] = (float)a.beta1;
    hyper[2] = (float)a.beta2;
    hyper[3] = (float)a.eps;
    hyper[4] = (float)(1.0 - pow(a.beta1, (double)t));
    hyper[5] = (float)sqrt(1.0 - pow(a.beta2, (double)t));
    hyper[6] = (float)a.grad_scale;
    hyper[7] = (float)((double)i < a.finnish_smooth ? a.smooth0 - a.dsmooth * (double)i : a.final_smooth);
    hyper[8] = (float)(1.0 - a.beta1);
    hyper[9] = (float)(1.0 - a.beta2);
    state[0] = i + 1;
    state[1] = t;
}

}  // namespace
}  // namespace ddnerf

using namespace ddnerf;

extern "C" DDNERF_EXPORT int ddnerf_train_schedule(int64_t* state, float* hyper, const double* sched, void* stream) {
    DDNERF_CHECK_ARG(state && hyper && sched, "train_schedule: null pointer");
    SchedArgs a;
    a.lr_init = sched[0]; a.lr_final = sched[1]; a.max_steps = sched[2]; a.lr_delay_steps = sched[3]; a.lr_delay_mult = sched[4];
    a.beta1 = sched[5]; a.beta2 = sched[6]; a.eps = sched[7]; a.grad_scale = sched[8];
    a.smooth0 = sched[9]; a.dsmooth = sched[10]; a.final_smooth = sched[11]; a.finnish_smooth = sched[12];
    DDNERF_CHECK_ARG(a.lr_init > 0 && a.lr_final > 0 && a.max_steps > 0, "train_schedule: lr_init, lr_final and max_steps must be positive");
    schedule_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<long long*>(state), hyper, a);
    DDNERF_LAUNCHED("train_schedule", 1);
    return 0;
}

extern "C" DDNERF_EXPORT int ddnerf_mse_loss(const float* rgb0, const float* rgb1, const float* target, float coef0, float coef1,
                               float* g_rgb0, float* g_rgb1, float* mse_out, int64_t N, void* stream) {
    DDNERF_CHECK_ARG(mse_out && (N == 0 || (rgb0 && target)), "mse_loss: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(mse_out, 0, 3 * sizeof(float), st);
    if (N == 0) return 0;
    int64_t n = N * 3;
    int blocks = (int)std::min<int64_t>((n + 255) / 256, 592);
    mse_kernel<<<blocks, 256, 0, st>>>(rgb0, rgb1, target, coef0, coef1, g_rgb0, g_rgb1, mse_out, n);
    DDNERF_LAUNCHED("mse_loss", 1);
    return 0;
}

extern "C" DDNERF_EXPORT int ddnerf_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                                double beta1, double beta2, double eps, int step, float grad_scale, void* stream) {
    DDNERF_CHECK_ARG(param && grad && exp_avg && exp_avg_sq, "adam_step: null pointer");
    DDNERF_CHECK_ARG(step >= 1, "adam_step: step=%d must be >= 1", step);
    if (n == 0) return 0;
    float bc1 = (float)(1.0 - pow(beta1, (double)step));
    float bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)step));
    int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
    adam_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, lr, (float)(1.0 - beta1),
                                                                        (float)beta2, (float)(1.0 - beta2), (float)eps, bc1,
                                                                        bc2_sqrt, grad_scale);
    DDNERF_LAUNCHED("adam_step", 1);
    return 0;
}

extern "C" DDNERF_EXPORT int ddnerf_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                                    const float* hyper, void* stream) {
    DDNERF_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && hyper, "adam_step_dev: null pointer");
    if (n == 0) return 0;
    int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
    adam_dev_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, hyper);
    DDNERF_LAUNCHED("adam_step_dev", 1);
    return 0;
}
