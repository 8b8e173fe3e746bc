// Descriptor self-test for the tcgen05 path: one CTA multiplies two operand tile IMAGES (already in
// the shared-memory byte layout the MLP kernels use) and returns the fp32 accumulator.  Every
// descriptor field comes from the caller, so the parity tests pin the swizzle / LBO / SBO / major
// conventions that mlp_tc.cu relies on against a CPU matmul, one convention per call.
#define DDNERF_TC_WATCHDOG 1
#include "tc.cuh"

namespace ddnerf {
namespace {

struct SelfTestArgs {
    const uint8_t* a_img; uint32_t a_bytes;
    const uint8_t* b_img; uint32_t b_bytes;
    float* d_out;                       // [128, N]
    int N, nk16;
    uint64_t a_desc, b_desc;            // templates with start address 0
    uint32_t idesc;
    uint32_t a_step, a_steps_per_block, a_block_pitch;   // address of k16-step j: (j / spb) * pitch + (j % spb) * step
    uint32_t b_step, b_steps_per_block, b_block_pitch;
};

__global__ void __launch_bounds__(128, 1) tc_selftest_kernel(SelfTestArgs g) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_base_holder;
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // SWIZZLE_128B atoms: 1024 B alignment
    uint8_t* sa = smem;
    uint8_t* sb = smem + ((g.a_bytes + 1023u) & ~1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tc::mbar_init(&bar_load, 1);
        tc::mbar_init(&bar_mma, 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_holder, 256);
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem = tmem_base_holder;
    if (threadIdx.x == 0) {
        tc::mbar_expect_tx(&bar_load, g.a_bytes + g.b_bytes);
        tc::bulk_g2s(sa, g.a_img, g.a_bytes, &bar_load);
        tc::bulk_g2s(sb, g.b_img, g.b_bytes, &bar_load);
        tc::mbar_wait(&bar_load, 0);
        tc::tc_fence_after_sync();
        const uint32_t a0 = tc::smem_u32(sa), b0 = tc::smem_u32(sb);
        for (int j = 0; j < g.nk16; ++j) {
            uint32_t ao = a0 + (j / g.a_steps_per_block) * g.a_block_pitch + (j % g.a_steps_per_block) * g.a_step;
            uint32_t bo = b0 + (j / g.b_steps_per_block) * g.b_block_pitch + (j % g.b_steps_per_block) * g.b_step;
            tc::mma_f16_ss(tmem, g.a_desc | (uint64_t)((ao >> 4) & 0x3FFF), g.b_desc | (uint64_t)((bo >> 4) & 0x3FFF), g.idesc, j > 0);
        }
        tc::mma_commit(&bar_mma);
    }
    __syncwarp();
    tc::mbar_wait(&bar_mma, 0);
    tc::tc_fence_after_sync();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < g.N; c0 += 16) {
        uint32_t v[16];
        tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (c0 + i < g.N) g.d_out[(size_t)row * g.N + c0 + i] = __uint_as_float(v[i]);
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

}  // namespace
}  // namespace ddnerf

using namespace ddnerf;

extern "C" DDNERF_EXPORT int ddnerf_tc_gemm_selftest(const void* a_img, int64_t a_bytes, const void* b_img, int64_t b_bytes,
                                                     float* d_out, int N, int nk16, uint64_t a_desc, uint64_t b_desc,
                                                     uint32_t idesc, const uint32_t* stepping, void* stream) {
    DDNERF_CHECK_ARG(a_img && b_img && d_out && stepping, "tc_gemm_selftest: null pointer");
    DDNERF_CHECK_ARG(N >= 16 && N <= 256 && N % 16 == 0 && nk16 >= 1, "tc_gemm_selftest: N=%d nk16=%d", N, nk16);
    DDNERF_CHECK_ARG(a_bytes % 16 == 0 && b_bytes % 16 == 0 && a_bytes + b_bytes <= 200 * 1024, "tc_gemm_selftest: image sizes");
    SelfTestArgs g{static_cast<const uint8_t*>(a_img), (uint32_t)a_bytes, static_cast<const uint8_t*>(b_img), (uint32_t)b_bytes,
                   d_out, N, nk16, a_desc, b_desc, idesc, stepping[0], stepping[1], stepping[2], stepping[3], stepping[4], stepping[5]};
    const int smem = (int)(((a_bytes + 1023) & ~1023ll) + b_bytes + 1024);
    cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    tc_selftest_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(g);
    DDNERF_LAUNCHED("tc_gemm_selftest", 1);
    return 0;
}

// ---- tensor-pipe rate probe (diagnostic): back-to-back tcgen05.mma on resident, arbitrary operands ----------
namespace ddnerf {
namespace {

struct RateArgs {
    int n_mma, commit_every, pair, N;
    unsigned long long* cycles;      // [gridDim.x]
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) tc_rate_kernel(RateArgs g) {
    extern __shared__ __align__(1024) uint8_t smem[];      // [A 64 KB | B 32 KB], contents irrelevant (finite garbage)
    __shared__ uint64_t bar_done, bar_stage, bar_stage2;
    __shared__ uint32_t tmem_holder;
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = tc::cluster_ctarank();
    for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        tc::mbar_init(&bar_done, 1);
        tc::mbar_init(&bar_stage, 1);
        tc::mbar_init(&bar_stage2, 1);
        tc::fence_barrier_init();
    }
    tc::fence_proxy_async_smem();
    tc::cluster_sync();
    if (warp == 0) { if (g.pair) tc::tmem_alloc2(&tmem_holder, 512); else tc::tmem_alloc(&tmem_holder, 512); }
    tc::tc_fence_before_sync();
    tc::cluster_sync();
    tc::tc_fence_after_sync();
    const uint32_t tmem = tmem_holder;
    if (threadIdx.x == 0 && (!g.pair || rank == 0)) {
        const uint32_t base16 = tc::smem_u32(smem) >> 4;
        constexpr uint32_t hi128 = (uint32_t)(tc::smem_desc(0, 0, 1024, tc::LAYOUT_SW128) >> 32);
        constexpr uint32_t hi64 = (uint32_t)(tc::smem_desc(0, 0, 512, tc::LAYOUT_SW64) >> 32);
        const uint32_t idesc = tc::idesc_bf16(g.pair ? 256 : 128, g.N, 0, 0);
        const int ce = g.commit_every < 0 ? -g.commit_every : g.commit_every;       // negative: two commits at a time
        int left = ce;
        const unsigned long long t0 = clock64();
        for (int i = 0; i < g.n_mma; ++i) {
            const uint64_t a = ((uint64_t)hi128 << 32) | (uint64_t)(base16 + (uint32_t)((i & 7) / 2) * 1024u + (uint32_t)(i & 1) * 2u);
            const uint64_t b = ((uint64_t)hi64 << 32) | (uint64_t)(base16 + 4096u + (uint32_t)((i >> 1) & 1) * 512u + (uint32_t)(i & 1) * 2u);
            const uint32_t d = tmem + (uint32_t)((i >> 4) & 1) * 256u;
            if (g.pair) tc::mma2_f16_ss(d, a, b, idesc, 1u); else tc::mma_f16_ss(d, a, b, idesc, 1u);
            if (ce > 0 && --left == 0) {
                left = ce;
                if (g.pair) tc::mma2_commit_u32(tc::smem_u32(&bar_stage)); else tc::mma_commit(&bar_stage);
                if (g.commit_every < 0) { if (g.pair) tc::mma2_commit_u32(tc::smem_u32(&bar_stage2)); else tc::mma_commit(&bar_stage2); }
            }
        }
        if (g.pair) tc::mma2_commit_u32(tc::smem_u32(&bar_done)); else tc::mma_commit(&bar_done);
        tc::mbar_wait(&bar_done, 0);
        g.cycles[blockIdx.x] = clock64() - t0;
    }
    tc::tc_fence_before_sync();
    tc::cluster_sync();
    if (warp == 0) { if (g.pair) tc::tmem_dealloc2(tmem, 512); else tc::tmem_dealloc(tmem, 512); }
}

}  // namespace
}  // namespace ddnerf

/* Diagnostic: cycles one CTA (pair = 0) or a CTA pair (pair = 1, cta_group::2) needs to run n_mma back-to-back
 * M = 128 (256) x N x 16 bf16 MMAs on resident operands, committing to an mbarrier every `commit_every` MMAs
 * (0 = never).  cycles_out: one uint64 per CTA of a 148-CTA launch (leader CTAs only in pair mode). */
extern "C" DDNERF_EXPORT int ddnerf_tc_mma_rate(int pair, int N, int n_mma, int commit_every, void* cycles_out, void* stream) {
    DDNERF_CHECK_ARG(cycles_out && n_mma > 0 && N >= 16 && N <= 256 && N % 16 == 0, "tc_mma_rate: bad arguments");
    RateArgs g{n_mma, commit_every, pair, N, static_cast<unsigned long long*>(cycles_out)};
    cudaFuncSetAttribute(tc_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 97 * 1024);
    tc_rate_kernel<<<148, 128, 97 * 1024, static_cast<cudaStream_t>(stream)>>>(g);
    DDNERF_LAUNCHED("tc_mma_rate", 1);
    return 0;
}
