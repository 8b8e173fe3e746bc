// K1 (fp32 parity mode): the NeRF MLP of models/base_architectures.py:3-126 as a chain of
// CUDA-core SGEMMs with fused bias / ReLU / ReLU-mask / accumulate epilogues, plus its backward
// (dX chain, dW as split-K over the sample dimension, db column sums).
//
// This is the bit-faithful fp32 mode (1e-3 parity budget of BASELINE.json); the throughput mode
// is the fused tcgen05 kernel in mlp_tc.cu.  Activations of every layer are kept in a caller
// workspace so the backward needs no recompute.  The concatenations of the reference
// (cat(xyz,h) at layer 5, cat(feat,dirs) at the view branch) are never materialised: layer 4
// writes straight into columns 96.. of the [rows,352] buffer whose first 96 columns the encoder
// filled, and fc_feat writes into columns 0..255 of the [rows,288] buffer holding the direction
// encoding at columns 256..282.
#include <algorithm>

#include "encode.cuh"

namespace ddnerf {
namespace {

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8, NT = 256;

enum : int { EPI_BIAS = 1, EPI_RELU = 2, EPI_MASK = 4, EPI_ACCUM = 8, EPI_ATOMIC = 16 };

struct GemmArgs {
    const float* A; int64_t sam, sak;      // A(m,k) = A[m*sam + k*sak]
    const float* B; int64_t sbk, sbn;      // B(k,n) = B[k*sbk + n*sbn]
    float* C; int64_t ldc;                 // C[m*ldc + n]
    const float* bias;                     // [N]
    const float* mask; int64_t ldm;        // relu mask source: C *= (mask[m*ldm+n] > 0)
    int64_t M; int N; int64_t K;
    int64_t k_per_split;
    int epi;
};

// C[M,N] = epi(A.B).  KCA/KCB: the k index is the unit-stride one for A / B (decides which
// thread->element mapping gives coalesced global loads).
template <bool KCA, bool KCB>
__global__ void __launch_bounds__(NT) sgemm_kernel(GemmArgs g) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;
    const int64_t kbeg = (int64_t)blockIdx.z * g.k_per_split;
    const int64_t kend = min(g.K, kbeg + g.k_per_split);
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
        float ra[8], rb[8];
        // ---- global -> registers
        if (KCA) {
            const int r = tid >> 1, kk = (tid & 1) * 8;
            const int64_t m = m0 + r;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                int64_t k = k0 + kk + e;
                ra[e] = (m < g.M && k < kend) ? __ldg(g.A + m * g.sam + k * g.sak) : 0.f;
            }
        } else {
            const int kk = tid >> 4, mm = (tid & 15) * 8;
            const int64_t k = k0 + kk;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                int64_t m = m0 + mm + e;
                ra[e] = (m < g.M && k < kend) ? __ldg(g.A + m * g.sam + k * g.sak) : 0.f;
            }
        }
        if (KCB) {
            const int r = tid >> 1, kk = (tid & 1) * 8;
            const int n = n0 + r;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                int64_t k = k0 + kk + e;
                rb[e] = (n < g.N && k < kend) ? __ldg(g.B + k * g.sbk + (int64_t)n * g.sbn) : 0.f;
            }
        } else {
            const int kk = tid >> 4, nn = (tid & 15) * 8;
            const int64_t k = k0 + kk;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                int n = n0 + nn + e;
                rb[e] = (n < g.N && k < kend) ? __ldg(g.B + k * g.sbk + (int64_t)n * g.sbn) : 0.f;
            }
        }
        __syncthreads();                                    // previous tile fully consumed
        if (KCA) {
            const int r = tid >> 1, kk = (tid & 1) * 8;
#pragma unroll
            for (int e = 0; e < 8; ++e) As[kk + e][r] = ra[e];
        } else {
            const int kk = tid >> 4, mm = (tid & 15) * 8;
#pragma unroll
            for (int e = 0; e < 8; ++e) As[kk][mm + e] = ra[e];
        }
        if (KCB) {
            const int r = tid >> 1, kk = (tid & 1) * 8;
#pragma unroll
            for (int e = 0; e < 8; ++e) Bs[kk + e][r] = rb[e];
        } else {
            const int kk = tid >> 4, nn = (tid & 15) * 8;
#pragma unroll
            for (int e = 0; e < 8; ++e) Bs[kk][nn + e] = rb[e];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * TM]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * TM + 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * TN]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][tx * TN + 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
    // ---- epilogue
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t m = m0 + ty * TM + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * TN + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            float* c = g.C + m * g.ldc + n;
            if (g.epi & EPI_ATOMIC) { atomicAdd(c, v); continue; }
            if (g.epi & EPI_ACCUM) v += *c;
            if (g.epi & EPI_BIAS) v += __ldg(g.bias + n);
            if (g.epi & EPI_RELU) v = fmaxf(v, 0.f);
            if (g.epi & EPI_MASK) v = __ldg(g.mask + m * g.ldm + n) > 0.f ? v : 0.f;
            *c = v;
        }
    }
}

// out[n] += sum_m X[m*ld + n]
__global__ void colsum_kernel(const float* __restrict__ X, int64_t ld, int64_t M, int N, float* __restrict__ out) {
    __shared__ float red[8][33];
    const int n = blockIdx.x * 32 + threadIdx.x;
    float s = 0.f;
    if (n < N)
        for (int64_t m = (int64_t)blockIdx.y * 8 + threadIdx.y; m < M; m += (int64_t)gridDim.y * 8) s += __ldg(X + m * ld + n);
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && n < N) {
#pragma unroll
        for (int r = 1; r < 8; ++r) s += red[r][threadIdx.x];
        atomicAdd(out + n, s);
    }
}

// x[rows,123] -> XH[:,0:96] and FD[:,256:283]
__global__ void scatter_x_kernel(const float* __restrict__ x, float* __restrict__ XH, float* __restrict__ FD, int64_t rows) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * 123) return;
    int64_t r = e / 123;
    int c = (int)(e - r * 123);
    float v = __ldg(x + e);
    if (c < 96) XH[r * 352 + c] = v; else FD[r * 288 + 256 + (c - 96)] = v;
}

struct Ws {       // fp32 workspace layout for `rows` sample rows
    float *XH, *H[8], *FD, *HD, *Ga, *Gb, *GHD, *GFD;
    static int64_t floats_per_row() { return 352 + 7 * 256 + 288 + 128 + 2 * 256 + 128 + 288; }
    Ws(void* base, int64_t rows) {
        float* p = static_cast<float*>(base);
        XH = p; p += rows * 352;
        for (int i = 0; i < 8; ++i) {
            if (i == 4) { H[i] = XH + 96; continue; }         // layer 4's output lives inside XH (ld 352)
            H[i] = p; p += rows * 256;
        }
        FD = p; p += rows * 288;
        HD = p; p += rows * 128;
        Ga = p; p += rows * 256;
        Gb = p; p += rows * 256;
        GHD = p; p += rows * 128;
        GFD = p; p += rows * 288;
    }
    int64_t ldh(int i) const { return i == 4 ? 352 : 256; }
};

int launch_gemm(cudaStream_t st, bool kca, bool kcb, GemmArgs g, int splits = 1) {
    if (g.M == 0 || g.N == 0) return 0;
    int64_t kps = g.K;
    if (splits > 1) {
        kps = ((g.K + splits - 1) / splits + BK - 1) / BK * BK;
        splits = (int)((g.K + kps - 1) / kps);
        g.epi = EPI_ATOMIC;
    }
    g.k_per_split = kps;
    dim3 grid((g.N + BN - 1) / BN, (unsigned)((g.M + BM - 1) / BM), splits);
    if (kca && kcb) sgemm_kernel<true, true><<<grid, NT, 0, st>>>(g);
    else if (kca && !kcb) sgemm_kernel<true, false><<<grid, NT, 0, st>>>(g);
    else if (!kca && !kcb) sgemm_kernel<false, false><<<grid, NT, 0, st>>>(g);
    else return 1;
    count_launches(1);
    return 0;
}

// Y[rows,N] = epi(X[rows,K] . W[N,K]^T + b)
int linear_fwd(cudaStream_t st, const float* X, int64_t ldx, const float* W, int K, const float* b, float* Y, int64_t ldy,
               int64_t rows, int N, bool relu) {
    GemmArgs g{X, ldx, 1, W, 1, K, Y, ldy, b, nullptr, 0, rows, N, K, 0, EPI_BIAS | (relu ? EPI_RELU : 0)};
    return launch_gemm(st, true, true, g);
}
// dX[rows,Kout] = epi(dY[rows,N] . W[N, koff:koff+Kout])
int linear_dx(cudaStream_t st, const float* dY, int64_t ldy, int N, const float* W, int64_t ldw, int koff, float* dX,
              int64_t ldx, int Kout, int64_t rows, const float* mask, int64_t ldm, bool accum) {
    GemmArgs g{dY, ldy, 1, W + koff, ldw, 1, dX, ldx, nullptr, mask, ldm, rows, Kout, N, 0,
               (mask ? EPI_MASK : 0) | (accum ? EPI_ACCUM : 0)};
    return launch_gemm(st, true, false, g);
}
// dW[N, 0:K] += dY[rows,N]^T . X[rows,K]  (split over rows, atomics);  db[N] += colsum(dY)
int linear_dw(cudaStream_t st, const float* dY, int64_t ldy, int N, const float* X, int64_t ldx, int K, float* dW,
              int64_t ldw, float* db, int64_t rows) {
    GemmArgs g{dY, 1, ldy, X, ldx, 1, dW, ldw, nullptr, nullptr, 0, N, K, rows, 0, EPI_ATOMIC};
    int tiles = ((N + BM - 1) / BM) * ((K + BN - 1) / BN);
    int splits = (int)std::min<int64_t>((rows + 2047) / 2048, std::max(1, 592 / tiles));
    if (splits < 2) {                  // single split still has to accumulate into dW
        g.k_per_split = rows;
        dim3 grid((K + BN - 1) / BN, (N + BM - 1) / BM, 1);
        sgemm_kernel<false, false><<<grid, NT, 0, st>>>(g);
        count_launches(1);
    } else if (launch_gemm(st, false, false, g, splits)) return 1;
    if (db) {
        dim3 grid((N + 31) / 32, (unsigned)std::min<int64_t>((rows + 255) / 256, 256));
        colsum_kernel<<<grid, dim3(32, 8), 0, st>>>(dY, ldy, rows, N, db);
        count_launches(1);
    }
    return 0;
}

int forward_impl(const ddnerf_mlp_params* p, int64_t rows, int C, float* out, Ws& w, cudaStream_t st) {
    const float* in = w.XH;
    int64_t ldin = 352;
    for (int l = 0; l < 8; ++l) {
        int K = l == 0 ? 96 : (l == 5 ? 352 : 256);
        if (l == 5) { in = w.XH; ldin = 352; }
        if (linear_fwd(st, in, ldin, p->w[l], K, p->b[l], w.H[l], w.ldh(l), rows, 256, true)) return 1;
        in = w.H[l]; ldin = w.ldh(l);
    }
    if (linear_fwd(st, w.H[7], 256, p->w[8], 256, p->b[8], w.FD, 288, rows, 256, false)) return 1;    // fc_feat
    if (linear_fwd(st, w.FD, 288, p->w[9], 256, p->b[9], out + 3, C, rows, 1, false)) return 1;       // fc_alpha
    if (linear_fwd(st, w.FD, 288, p->w[10], 283, p->b[10], w.HD, 128, rows, 128, true)) return 1;     // layers_dir.0
    if (linear_fwd(st, w.HD, 128, p->w[11], 128, p->b[11], out, C, rows, 3, false)) return 1;         // fc_rgb
    if (C == 6 && linear_fwd(st, w.HD, 128, p->w[12], 128, p->b[12], out + 4, C, rows, 2, false)) return 1;
    return 0;
}

}  // namespace
}  // namespace ddnerf

using namespace ddnerf;

extern "C" DDNERF_EXPORT int64_t ddnerf_mlp_f32_workspace_bytes(int64_t rows) {
    return rows * Ws::floats_per_row() * (int64_t)sizeof(float);
}

static int check_params(const ddnerf_mlp_params* p, int C, const char* who) {
    DDNERF_CHECK_ARG(p, "%s: null params", who);
    DDNERF_CHECK_ARG(C == 4 || C == 6, "%s: out_channels=%d (4 or 6)", who, C);
    for (int i = 0; i < (C == 6 ? 13 : 12); ++i)
        DDNERF_CHECK_ARG(p->w[i] && p->b[i], "%s: parameter %d is null", who, i);
    return 0;
}

extern "C" DDNERF_EXPORT int ddnerf_mlp_f32_forward(const ddnerf_mlp_params* p, const float* rays, const float* t_vals, int64_t N, int S,
                                      int ray_shape, int out_channels, float* out, void* workspace, void* stream) {
    if (check_params(p, out_channels, "mlp_f32_forward")) return 1;
    DDNERF_CHECK_ARG(rays && t_vals && out && workspace, "mlp_f32_forward: null pointer");
    const int64_t rows = N * S;
    if (rows == 0) return 0;
    Ws w(workspace, rows);
    if (ddnerf_encode(rays, t_vals, w.XH, 352, w.FD + 256, 288, N, S, ray_shape, stream)) return 1;
    DDNERF_CHECK_ARG(forward_impl(p, rows, out_channels, out, w, static_cast<cudaStream_t>(stream)) == 0,
                     "mlp_f32_forward: launch configuration error");
    DDNERF_CHECK_LAUNCH("mlp_f32_forward");
    return 0;
}

extern "C" DDNERF_EXPORT int ddnerf_mlp_f32_forward_x(const ddnerf_mlp_params* p, const float* x, int64_t rows, int out_channels,
                                        float* out, void* workspace, void* stream) {
    if (check_params(p, out_channels, "mlp_f32_forward_x")) return 1;
    DDNERF_CHECK_ARG(x && out && workspace, "mlp_f32_forward_x: null pointer");
    if (rows == 0) return 0;
    Ws w(workspace, rows);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    scatter_x_kernel<<<ceil_div(rows * 123, 256), 256, 0, st>>>(x, w.XH, w.FD, rows);
    count_launches(1);
    DDNERF_CHECK_ARG(forward_impl(p, rows, out_channels, out, w, st) == 0, "mlp_f32_forward_x: launch configuration error");
    DDNERF_CHECK_LAUNCH("mlp_f32_forward_x");
    return 0;
}

extern "C" DDNERF_EXPORT int ddnerf_mlp_f32_backward(const ddnerf_mlp_params* p, const ddnerf_mlp_grads* g, const float* grad_out,
                                       int64_t rows, int out_channels, float* dx, void* workspace, void* stream) {
    if (check_params(p, out_channels, "mlp_f32_backward")) return 1;
    DDNERF_CHECK_ARG(g && grad_out && workspace, "mlp_f32_backward: null pointer");
    if (rows == 0) return 0;
    const int C = out_channels;
    Ws w(workspace, rows);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = 0;
    // heads: d_HD = (g_rgb.W_rgb [+ g_ms.W_ms]) * relu'(HD)
    rc |= linear_dw(st, grad_out, C, 3, w.HD, 128, 128, g->w[11], 128, g->b[11], rows);
    if (C == 6) {
        rc |= linear_dw(st, grad_out + 4, C, 2, w.HD, 128, 128, g->w[12], 128, g->b[12], rows);
        rc |= linear_dx(st, grad_out, C, 3, p->w[11], 128, 0, w.GHD, 128, 128, rows, nullptr, 0, false);
        rc |= linear_dx(st, grad_out + 4, C, 2, p->w[12], 128, 0, w.GHD, 128, 128, rows, w.HD, 128, true);
    } else {
        rc |= linear_dx(st, grad_out, C, 3, p->w[11], 128, 0, w.GHD, 128, 128, rows, w.HD, 128, false);
    }
    // view branch + density head -> d_feat (GFD[:, 0:256]), optionally d_dirs (GFD[:, 256:283])
    rc |= linear_dw(st, w.GHD, 128, 128, w.FD, 288, 283, g->w[10], 283, g->b[10], rows);
    rc |= linear_dw(st, grad_out + 3, C, 1, w.FD, 288, 256, g->w[9], 256, g->b[9], rows);
    rc |= linear_dx(st, w.GHD, 128, 128, p->w[10], 283, 0, w.GFD, 288, dx ? 283 : 256, rows, nullptr, 0, false);
    rc |= linear_dx(st, grad_out + 3, C, 1, p->w[9], 256, 0, w.GFD, 288, 256, rows, nullptr, 0, true);
    // fc_feat (no activation) -> d_h7 masked by relu'(H7)
    rc |= linear_dw(st, w.GFD, 288, 256, w.H[7], 256, 256, g->w[8], 256, g->b[8], rows);
    float* cur = w.Ga;
    float* nxt = w.Gb;
    rc |= linear_dx(st, w.GFD, 288, 256, p->w[8], 256, 0, cur, 256, 256, rows, w.H[7], 256, false);
    // trunk, layers 7..1: cur = dL/d(pre-activation of layer l)
    for (int l = 7; l >= 1; --l) {
        const int K = l == 5 ? 352 : 256;
        const float* xin = l == 5 ? w.XH : w.H[l - 1];
        const int64_t ldx = l == 5 ? 352 : w.ldh(l - 1);
        rc |= linear_dw(st, cur, 256, 256, xin, ldx, K, g->w[l], K, g->b[l], rows);
        // d h_{l-1} = cur . W_l[:, hoff:hoff+256], masked by relu'(h_{l-1})
        const int hoff = l == 5 ? 96 : 0;
        rc |= linear_dx(st, cur, 256, 256, p->w[l], K, hoff, nxt, 256, 256, rows, w.H[l - 1], w.ldh(l - 1), false);
        if (l == 5 && dx)               // skip connection's xyz part -> dx[:, 0:96]
            rc |= linear_dx(st, cur, 256, 256, p->w[5], 352, 0, dx, 123, 96, rows, nullptr, 0, false);
        float* t = cur; cur = nxt; nxt = t;
    }
    rc |= linear_dw(st, cur, 256, 256, w.XH, 352, 96, g->w[0], 96, g->b[0], rows);
    if (dx) {
        rc |= linear_dx(st, cur, 256, 256, p->w[0], 96, 0, dx, 123, 96, rows, nullptr, 0, true);
        // dx[:, 96:123] = GFD[:, 256:283]
        rc |= cudaMemcpy2DAsync(dx + 96, 123 * sizeof(float), w.GFD + 256, 288 * sizeof(float), 27 * sizeof(float), rows,
                                cudaMemcpyDeviceToDevice, st) != cudaSuccess;
    }
    DDNERF_CHECK_ARG(rc == 0, "mlp_f32_backward: launch configuration error");
    DDNERF_CHECK_LAUNCH("mlp_f32_backward");
    return 0;
}
