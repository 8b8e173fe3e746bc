// K1 (bf16 throughput mode): the NeRF MLP of models/base_architectures.py:3-126 as persistent,
// warp-specialised tcgen05 chain kernels; all GEMMs of the network run back to back with the
// activations resident on chip.  Two variants with bit-identical results share the epilogues, the
// encoder, the packed weight image and the saved-tile formats:
//
//   mlp_tc_pair_kernel  (default): a cluster of two CTAs works on 512 rows = two super-tiles of 256;
//                    tcgen05.mma.cta_group::2 (M = 256) takes 128 rows of A and half of every weight
//                    chunk from each CTA, the ring stages arrive by tiled TMA that signals the leader's
//                    barrier from both CTAs, and the super-tiles run a whole layer apart (see mlp_tc.cuh
//                    and the section "CTA-pair variant" below);
//   mlp_tc_chain_kernel (DDNERF_TC_PAIR=0 / ddnerf_mlp_tc_set_pair_mode(0)): one CTA owns a 256-row work
//                    item (two 128-row tiles), described next.
//
//   producer warp  : walks the static load program, streaming 16 KB weight stages (and the encoded
//                    xyz / view-direction blocks) from L2 into a 6-slot shared-memory ring with
//                    bulk async copies (TMA engine) completing on mbarriers;
//   MMA warp       : issues tcgen05.mma (M=128, N=256/144/16, K=16, bf16 -> fp32 in TMEM) from a
//                    host-resolved op table; the two tiles share every weight stage and are
//                    interleaved half a layer apart so that one tile's epilogue overlaps the other
//                    tile's MMAs;
//   8 epilogue warps: tcgen05.ld the accumulator, add bias, ReLU, convert to bf16 and write the next
//                    layer's A operand in place (K-major SWIZZLE_128B); density / colour / (mu, sigma)
//                    heads are extra columns of the last two GEMMs and leave as fp32;
//   2 encoder warps : (forward) cone -> Gaussian, integrated positional encoding and view-direction encoding
//                    (models.py:117-133) of the work item AFTER the one in flight, written as the bf16 operand image
//                    the producer streams into the ring -- a per-CTA double buffer that never leaves L2 at inference,
//                    the saved image the weight-gradient kernel reads in training.  No separate encode launch.
//
// FORWARD (program 0): layers_xyz.0-7, fc_feat, [layers_dir.0 | fc_alpha], [fc_rgb | fc_mu_sigma].
// The skip connection cat(xyz, h) of layer 5 and the cat(feat, dirs) of the view branch are extra
// K-chunks of the same accumulation whose A operand is the encoded block in the ring.  In training
// the epilogues also store every layer's bf16 activations (tile images, bulk stores) and ReLU sign
// bitmasks.
//
// BACKWARD (program 1, the dX chain): dZ_dir = (g_rgb.W_rgb + g_musig.W_musig) * relu' on CUDA cores,
// then dZ_feat = [dZ_dir | g_density] . [W_dir[:, :256] ; w_alpha], dZ_l = (dZ_{l+1} . W_{l+1}) *
// relu'_l down to layer 0, with transposed weight stages.  Every dZ tile image is stored for the
// weight-gradient kernel (mlp_tc_dw.cu).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>

#define DDNERF_TC_WATCHDOG 1

#include "encode.cuh"
#include "mlp_tc.cuh"

namespace ddnerf {
namespace {

using namespace tcmlp;

__constant__ Program c_prog[2];          // 0 forward, 1 backward (dX chain)
__constant__ PackTable c_pack;

struct ChainArgs {
    const uint8_t* wimg;      // packed weight stages of this program (program order)
    const float* bias;        // fwd: packed fp32 biases, [n_epis][256]
    const uint8_t* enc;       // fwd: encoded-feature images: enc_mode 0/1 [n_items][64 KB], enc_mode 2 [grid][2][64 KB]
    // in-kernel encoder (enc_mode 1: writes the full image, training; 2: per-CTA double buffer, inference; 0: image given)
    const float* rays;        // [N,12]
    const float* t_vals;      // [N,S+1]
    int64_t N;
    int S, ray_shape, enc_mode;
    float* out;               // fwd: [rows, C]
    const float* gout;        // bwd: [rows, C] cotangent of the output
    uint8_t* save;            // fwd: act_save or null; bwd: dz_save.  [layers][n_tiles][64 KB]
    uint32_t* mask;           // [layers][n_tiles][128][8]: fwd writes (or null), bwd reads
    int64_t rows;
    int n_items, C;
    int enc_yield;            // the in-kernel encoder yields to the epilogue warps (DDNERF_TC_ENC_YIELD, default 1)
    int pslots;               // experiment knob (DDNERF_TC_PSLOTS): ring slots the pair kernels use (<= kSlots)
    int save_alias;           // experiment knob (DDNERF_TC_SAVE_ALIAS): saves go to tile % save_alias (L2-resident)
    unsigned long long* prof; // optional [grid][8] cycle counters (ddnerf_mlp_tc_set_profile_buffer), else null
};

__device__ __forceinline__ unsigned long long clk() { return clock64(); }

// The work-unit loop of a CTA and the tile numbering, for both kernel variants.  Single CTA: a unit is a 256-row item
// (tiles 2 unit + T).  CTA pair: a unit is 512 rows = two super-tiles, tile (2 unit + T) * 2 + rank; a unit may reach
// past the allocated tiles (odd item count) -- such a tile is computed on zeros and neither stored nor read.
struct Topo {
    int first, step, n_units;
    int pair, rank;
    int n_tiles;                    // allocated 128-row tiles (2 per 256-row item)
    uint32_t act_ready_leader;      // pair: shared::cluster address of the leader's act_ready[0]
    __device__ __forceinline__ int tile(int unit, int T) const { return pair ? (2 * unit + T) * 2 + rank : unit * 2 + T; }
};

// descriptor high words: K-major SWIZZLE_128B act buffer (SBO = 8 rows x 128 B), SWIZZLE_64B ring stages
constexpr uint32_t kHi128 = (uint32_t)(tc::smem_desc(0, 0, 1024, tc::LAYOUT_SW128) >> 32);
constexpr uint32_t kHi64 = (uint32_t)(tc::smem_desc(0, 0, 512, tc::LAYOUT_SW64) >> 32);

struct __align__(16) SmemCtl {
    uint64_t full[kSlots], empty[kSlots], acc_full[2], act_ready[2];
    uint64_t enc_ready[2], enc_free[2];  // encoder -> producer: image of item parity p written; MMA warp -> encoder: consumed
    uint32_t tmem_base;
    uint32_t epi_busy;                   // forward: the epilogue warps are converting an accumulator (the encoder warps hold back meanwhile)
    uint32_t pad[2];
    float bias[2][256];
};
constexpr int kSmemBytes = 2 * kActBytes + kSlots * kSlotBytes + (int)sizeof(SmemCtl);

__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// 32 bf16 (16 packed words) of row `row`, columns [c0, c0+32), into the K-major SWIZZLE_128B act buffer
__device__ __forceinline__ void store_chunk32(uint32_t act_u32, int row, int c0, const uint32_t (&pk)[16]) {
    const uint32_t base = act_u32 + (uint32_t)(c0 >> 6) * 16384u + (uint32_t)row * 128u;
    const uint32_t ch0 = (uint32_t)(c0 & 63) >> 3;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        st_shared_v4(base + ((((ch0 + i) ^ (uint32_t)row) & 7u) << 4), pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
}

// ---- epilogues --------------------------------------------------------------------------------
// The two tiles of a work item finish their layers half a layer apart, so their epilogues never
// coincide: all 8 epilogue warps serve whichever tile is ready.  Warp w reads TMEM lane quarter
// w % 4 (rows 32 (w % 4) + lane) and owns column half w / 4 (128 of the 256 accumulator columns).
// Epilogues run in program order, alternating tiles: (e0,T0) (e0,T1) (e1,T0) ...
constexpr int kEpiThreads = 256;

// 32 accumulator columns [c0, c0+32) of this thread's row: + bias, optional ReLU, -> bf16, stored as
// four 16-byte chunks of the act buffer.  Returns the ReLU mask (bit i set: value i has its sign bit
// clear, i.e. is positive -- or +0, where passing the gradient is immaterial), collected with one
// funnel shift per value.
template <bool RELU, bool MASK = true>
__device__ __forceinline__ uint32_t epi_chunk32(const uint32_t (&v)[32], const float* __restrict__ bias_s, int c0,
                                                uint32_t act_u32, int row) {
    uint32_t pk[16];
    uint32_t neg = 0;
    float x[32];
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        const float4 b = *reinterpret_cast<const float4*>(bias_s + c0 + i);
        x[i] = __uint_as_float(v[i]) + b.x;
        x[i + 1] = __uint_as_float(v[i + 1]) + b.y;
        x[i + 2] = __uint_as_float(v[i + 2]) + b.z;
        x[i + 3] = __uint_as_float(v[i + 3]) + b.w;
    }
    if (RELU && MASK) {                     // inference (no mask buffer) skips the 32 funnel shifts
#pragma unroll
        for (int i = 31; i >= 0; --i) neg = __funnelshift_l(__float_as_uint(x[i]), neg, 1);     // bit i = sign of x[i]
    }
#pragma unroll
    for (int i = 0; i < 32; i += 2) pk[i / 2] = RELU ? tc::pack_bf16_relu(x[i], x[i + 1]) : tc::pack_bf16(x[i], x[i + 1]);
    store_chunk32(act_u32, row, c0, pk);
    return ~neg;
}

__device__ __forceinline__ void signal_act_ready(const Topo& tp, SmemCtl* ctl, int T);

__device__ void epilogue_fwd(const ChainArgs& g, const Topo& tp, SmemCtl* ctl, uint8_t* act_all, int warp, int lane) {
    const Program& P = c_prog[0];
    const int q = warp & 3, hf = warp >> 2, row = q * 32 + lane, tid = warp * 32 + lane;
    const uint32_t tmem_q = ctl->tmem_base + ((uint32_t)(q * 32) << 16);
    const int n_tiles = g.n_items * 2;
    const int n_epis = P.n_epis;
    uint32_t acc_phase = 0;                  // bit T: parity of acc_full[T]
    int stores = 0;                          // bulk stores issued by thread 0
    const uint64_t pol_stream = tc::l2_policy_evict_first();     // saved tiles are not re-read before they leave L2
    uint32_t k = 0;                          // running epilogue count: parity selects the bias buffer
    unsigned long long t_wait = 0, t_busy = 0, n_epi = 0, t_pre = 0, t_work = 0;

    ctl->bias[0][tid] = __ldg(g.bias + P.epis[0].bias_off + tid);          // bias of the first epilogue
    named_bar(1, kEpiThreads);
    for (int unit = tp.first; unit < tp.n_units; unit += tp.step) {
        for (int e = 0; e < n_epis; ++e, ++k) {
            const Epi E = P.epis[e];
            const float* bias_s = ctl->bias[k & 1];
            const float nb = __ldg(g.bias + P.epis[(e + 1 == n_epis) ? 0 : e + 1].bias_off + tid);
#pragma unroll 1
            for (int T = 0; T < 2; ++T) {
                const int tile_g = tp.tile(unit, T);
                const bool tile_ok = tile_g < n_tiles;
                const int64_t row_g = (int64_t)tile_g * 128 + row;
                const uint32_t act_u32 = tc::smem_u32(act_all + T * kActBytes);
                const uint32_t tmem_row = tmem_q + (uint32_t)T * 256u;
                const unsigned long long tw0 = g.prof ? clk() : 0;
                tc::mbar_wait(&ctl->acc_full[T], (acc_phase >> T) & 1u);
                acc_phase ^= 1u << T;
                tc::tc_fence_after_sync();
                if (tid == 0) *reinterpret_cast<volatile uint32_t*>(&ctl->epi_busy) = 1u;
                const unsigned long long tw1 = g.prof ? clk() : 0;
                if (g.save) {            // this tile's previous bulk store must have finished reading the act buffer
                    if (tid == 0 && stores >= 2) tc::bulk_wait_read<1>();
                    named_bar(1, kEpiThreads);
                }
                const unsigned long long tw2 = g.prof ? clk() : 0;
                if (E.mode == EPI_ACT || E.mode == EPI_DIR) {
                    const int nc = (E.mode == EPI_DIR) ? 64 : 128;         // columns of this half
                    const int cb = hf * nc;
                    uint32_t mk[4] = {0u, 0u, 0u, 0u};
                    uint32_t va[32], vb[32];
                    tc::tmem_ld32(tmem_row + cb, va);
                    if (E.relu && g.mask) {
#pragma unroll
                        for (int c = 0; c < 4; c += 2) {                   // TMEM loads one chunk ahead of the math
                            tc::tmem_ld_wait();
                            if ((c + 1) * 32 < nc) tc::tmem_ld32(tmem_row + cb + (c + 1) * 32, vb);
                            if (c * 32 < nc) mk[c] = epi_chunk32<true>(va, bias_s, cb + c * 32, act_u32, row);
                            tc::tmem_ld_wait();
                            if ((c + 2) * 32 < nc) tc::tmem_ld32(tmem_row + cb + (c + 2) * 32, va);
                            if ((c + 1) * 32 < nc) mk[c + 1] = epi_chunk32<true>(vb, bias_s, cb + (c + 1) * 32, act_u32, row);
                        }
                    } else if (E.relu) {                                   // inference: no ReLU masks to collect
#pragma unroll
                        for (int c = 0; c < 4; c += 2) {
                            tc::tmem_ld_wait();
                            if ((c + 1) * 32 < nc) tc::tmem_ld32(tmem_row + cb + (c + 1) * 32, vb);
                            if (c * 32 < nc) epi_chunk32<true, false>(va, bias_s, cb + c * 32, act_u32, row);
                            tc::tmem_ld_wait();
                            if ((c + 2) * 32 < nc) tc::tmem_ld32(tmem_row + cb + (c + 2) * 32, va);
                            if ((c + 1) * 32 < nc) epi_chunk32<true, false>(vb, bias_s, cb + (c + 1) * 32, act_u32, row);
                        }
                    } else {                                               // fc_feat: no activation, no mask
#pragma unroll
                        for (int c = 0; c < 4; c += 2) {
                            tc::tmem_ld_wait();
                            tc::tmem_ld32(tmem_row + cb + (c + 1) * 32, vb);
                            epi_chunk32<false>(va, bias_s, cb + c * 32, act_u32, row);
                            tc::tmem_ld_wait();
                            if (c + 2 < 4) tc::tmem_ld32(tmem_row + cb + (c + 2) * 32, va);
                            epi_chunk32<false>(vb, bias_s, cb + (c + 1) * 32, act_u32, row);
                        }
                    }
                    if (E.mode == EPI_DIR && hf == 0) {                    // column 128 = density (fc_alpha)
                        uint32_t v[16];
                        tc::tmem_ld16(tmem_row + 128, v);
                        tc::tmem_ld_wait();
                        if (row_g < g.rows) g.out[row_g * g.C + 3] = __uint_as_float(v[0]) + bias_s[128];
                    }
                    if (g.mask && E.save_layer >= 0 && E.relu && tile_ok) {
                        uint32_t* mp = g.mask + (((size_t)E.save_layer * n_tiles + tile_g) * 128 + row) * 8;
                        if (E.mode == EPI_DIR) {                           // words 0..1 / 2..3 of the 128 view-branch columns
                            *reinterpret_cast<uint2*>(mp + 2 * hf) = make_uint2(mk[0], mk[1]);
                            if (hf == 1) *reinterpret_cast<uint4*>(mp + 4) = make_uint4(0u, 0u, 0u, 0u);
                        } else {
                            *reinterpret_cast<uint4*>(mp + 4 * hf) = make_uint4(mk[0], mk[1], mk[2], mk[3]);
                        }
                    }
                } else if (hf == 0) {                                      // EPI_OUT: colour (+ mu, sigma) heads
                    uint32_t v[16];
                    tc::tmem_ld16(tmem_row, v);
                    tc::tmem_ld_wait();
                    if (row_g < g.rows) {
                        float* o = g.out + row_g * g.C;
                        o[0] = __uint_as_float(v[0]) + bias_s[0];
                        o[1] = __uint_as_float(v[1]) + bias_s[1];
                        o[2] = __uint_as_float(v[2]) + bias_s[2];
                        if (g.C == 6) {
                            o[4] = __uint_as_float(v[3]) + bias_s[3];
                            o[5] = __uint_as_float(v[4]) + bias_s[4];
                        }
                    }
                }
                const unsigned long long tw3 = g.prof ? clk() : 0;
                tc::tc_fence_before_sync();          // TMEM reads done before the MMA warp may overwrite D
                tc::fence_proxy_async_smem();        // act writes visible to tcgen05.mma / bulk store
                if (T == 1) ctl->bias[(k + 1) & 1][tid] = nb;      // (the other buffer: nobody reads it now)
                named_bar(1, kEpiThreads);
                if (tid == 0) {
                    signal_act_ready(tp, ctl, T);
                    *reinterpret_cast<volatile uint32_t*>(&ctl->epi_busy) = 0u;
                    if (g.prof) { t_wait += tw1 - tw0; t_busy += clk() - tw1; ++n_epi; t_pre += tw2 - tw1; t_work += tw3 - tw2; }
                    if (g.save && E.save_layer >= 0) {
                        const int tile_s = g.save_alias ? tile_g % g.save_alias : tile_g;
                        if (tile_ok)
                            tc::bulk_s2g_hint(g.save + ((size_t)E.save_layer * n_tiles + tile_s) * kActBytes, act_all + T * kActBytes,
                                              E.save_bytes, pol_stream);
                        tc::bulk_commit();          // (an empty group for a tile past the end keeps the alternation of the waits)
                        ++stores;
                    }
                }
            }
        }
    }
    if (tid == 0) tc::bulk_wait_all<0>();
    if (g.prof && tid == 0) {
        unsigned long long* o = g.prof + (size_t)blockIdx.x * 8;
        o[3] = t_wait; o[4] = t_busy; o[5] = n_epi;     // epilogues: waiting for MMAs / busy / count
        o[6] = t_pre; o[7] = t_work;                    // of busy: store-drain + barrier / accumulator -> act buffer
    }
}

__device__ __forceinline__ float masked(float x, uint32_t m, int bit) { return ((m >> bit) & 1u) ? x : 0.f; }

// dZ = dA * relu'(Z): convert first, then clear the masked bf16 halves.  Four mask bits become the sign bits of the
// four bytes of one register by a multiply (bit k -> bit 8k+7: 0x10204080 = 2^7 + 2^14 + 2^21 + 2^28, no carries),
// and PRMT's sign-replicate mode spreads each into a 16-bit lane mask: 9 instructions per 4 values instead of 12.
__device__ __forceinline__ void bwd_chunk32(const uint32_t (&v)[32], uint32_t m, int c0, uint32_t act_u32, int row) {
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        const uint32_t r = ((m >> i) & 15u) * 0x10204080u;
        uint32_t w01, w23;
        asm("prmt.b32 %0, %1, %1, 0x9988;" : "=r"(w01) : "r"(r));
        asm("prmt.b32 %0, %1, %1, 0xbbaa;" : "=r"(w23) : "r"(r));
        pk[i / 2] = tc::pack_bf16(__uint_as_float(v[i]), __uint_as_float(v[i + 1])) & w01;
        pk[i / 2 + 1] = tc::pack_bf16(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])) & w23;
    }
    store_chunk32(act_u32, row, c0, pk);
}

__device__ void epilogue_bwd(const ChainArgs& g, const Topo& tp, SmemCtl* ctl, uint8_t* act_all, int warp, int lane) {
    const Program& P = c_prog[1];
    const int q = warp & 3, hf = warp >> 2, row = q * 32 + lane, tid = warp * 32 + lane;
    const uint32_t tmem_q = ctl->tmem_base + ((uint32_t)(q * 32) << 16);
    const int n_tiles = g.n_items * 2;
    const int n_epis = P.n_epis;
    uint32_t acc_phase = 0;
    int stores = 0;
    const uint64_t pol_stream = tc::l2_policy_evict_first();
    unsigned long long t_wait = 0, t_busy = 0, n_epi = 0, t_pre = 0, t_work = 0;

    // ReLU mask words of this thread's row and column half for epilogue (unit, e, T); all ones where no mask applies
    auto fetch_mask = [&](int unit, int e, int T) -> uint4 {
        uint4 m = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        if (unit >= tp.n_units) return m;
        const int ml = P.epis[e].mask_layer;
        const int tile = tp.tile(unit, T);
        if (ml < 0 || tile >= n_tiles) return m;
        const uint32_t* mp = g.mask + (((size_t)ml * n_tiles + tile) * 128 + row) * 8;
        if (P.epis[e].mode == EPI_BWD_IN) {               // view branch: 2 words per half
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(mp + 2 * hf));
            m.x = v.x; m.y = v.y;
        } else {
            m = __ldg(reinterpret_cast<const uint4*>(mp + 4 * hf));
        }
        return m;
    };
    uint4 mk_next = fetch_mask(tp.first, 0, 0);
    for (int unit = tp.first; unit < tp.n_units; unit += tp.step) {
        for (int e = 0; e < n_epis; ++e) {
            const Epi E = P.epis[e];
#pragma unroll 1
            for (int T = 0; T < 2; ++T) {
                const int tile_g = tp.tile(unit, T);
                const bool tile_ok = tile_g < n_tiles;
                const int64_t row_g = (int64_t)tile_g * 128 + row;
                const uint32_t act_u32 = tc::smem_u32(act_all + T * kActBytes);
                const uint32_t tmem_row = tmem_q + (uint32_t)T * 256u;
                const unsigned long long tw0 = g.prof ? clk() : 0;
                // This half's ReLU mask words were requested one epilogue ago: they come from HBM, and the dX chain's epilogues
                // are busy 3/4 of the time, so a load issued here would put its whole latency between "accumulator ready" and
                // the first store.  Request the next epilogue's words now.
                const uint32_t mk[4] = {mk_next.x, mk_next.y, mk_next.z, mk_next.w};
                {
                    int nu = unit, ne = e, nT = T ^ 1;
                    if (T == 1 && ++ne == n_epis) { ne = 0; nu += tp.step; }
                    mk_next = fetch_mask(nu, ne, nT);
                }
                if (E.mode != EPI_BWD_IN) {
                    tc::mbar_wait(&ctl->acc_full[T], (acc_phase >> T) & 1u);
                    acc_phase ^= 1u << T;
                    tc::tc_fence_after_sync();
                }
                const unsigned long long tw1 = g.prof ? clk() : 0;
                if (tid == 0 && stores >= 2) tc::bulk_wait_read<1>();
                named_bar(1, kEpiThreads);
                const unsigned long long tw2 = g.prof ? clk() : 0;
                if (E.mode == EPI_BWD_IN) {
                    // dZ_dir = (g_rgb . W_rgb + g_musig . W_musig) * relu'(dir layer); 64 of its 128 columns per half
                    const float* w_rgb = g.bias + kHeadWRow * 256;          // aligned fp32 copies of the head weights
                    const float* w_musig = w_rgb + 3 * 256;
                    float gr[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    if (row_g < g.rows) {
                        const float* gp = g.gout + row_g * g.C;
#pragma unroll
                        for (int c = 0; c < 6; ++c)
                            if (c < g.C) gr[c] = __ldg(gp + c);
                    }
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        uint32_t pk[16];
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const int col = hf * 64 + c * 32 + i;
                            const float4 w0 = __ldg(reinterpret_cast<const float4*>(w_rgb + col));
                            const float4 w1 = __ldg(reinterpret_cast<const float4*>(w_rgb + 256 + col));
                            const float4 w2 = __ldg(reinterpret_cast<const float4*>(w_rgb + 512 + col));
                            float x0 = gr[0] * w0.x + gr[1] * w1.x + gr[2] * w2.x;
                            float x1 = gr[0] * w0.y + gr[1] * w1.y + gr[2] * w2.y;
                            float x2 = gr[0] * w0.z + gr[1] * w1.z + gr[2] * w2.z;
                            float x3 = gr[0] * w0.w + gr[1] * w1.w + gr[2] * w2.w;
                            if (g.C == 6) {
                                const float4 u0 = __ldg(reinterpret_cast<const float4*>(w_musig + col));
                                const float4 u1 = __ldg(reinterpret_cast<const float4*>(w_musig + 256 + col));
                                x0 += gr[4] * u0.x + gr[5] * u1.x;
                                x1 += gr[4] * u0.y + gr[5] * u1.y;
                                x2 += gr[4] * u0.z + gr[5] * u1.z;
                                x3 += gr[4] * u0.w + gr[5] * u1.w;
                            }
                            pk[i / 2] = tc::pack_bf16(masked(x0, mk[c], i), masked(x1, mk[c], i + 1));
                            pk[i / 2 + 1] = tc::pack_bf16(masked(x2, mk[c], i + 2), masked(x3, mk[c], i + 3));
                        }
                        store_chunk32(act_u32, row, hf * 64 + c * 32, pk);
                    }
                    if (hf == 0) {
                        // columns 128..135 = [g_density, g_r, g_g, g_b, g_mu, g_sigma, 0, 0]: column 128 is the K extension
                        // of the next GEMM (weight rows 129..143 are zero); all six feed the head gradients in mlp_tc_dw.cu
                        const uint32_t b2 = act_u32 + 2u * 16384u + (uint32_t)row * 128u;
                        st_shared_v4(b2 + (((0u ^ (uint32_t)row) & 7u) << 4), tc::pack_bf16(gr[3], gr[0]), tc::pack_bf16(gr[1], gr[2]),
                                     tc::pack_bf16(gr[4], gr[5]), 0u);
                        // the rest of the k-block (columns 136..191) is part of the saved image: zeros, not the previous
                        // layer's leftovers (they only reach accumulator rows nobody reads, but the image stays deterministic)
#pragma unroll
                        for (uint32_t ch = 1; ch < 8; ++ch) st_shared_v4(b2 + (((ch ^ (uint32_t)row) & 7u) << 4), 0u, 0u, 0u, 0u);
                    }
                } else {
                    const int cb = hf * 128;
                    uint32_t va[32], vb[32];
                    tc::tmem_ld32(tmem_row + cb, va);
#pragma unroll
                    for (int c = 0; c < 4; c += 2) {                       // TMEM loads one chunk ahead of the math
                        tc::tmem_ld_wait();
                        tc::tmem_ld32(tmem_row + cb + (c + 1) * 32, vb);
                        bwd_chunk32(va, mk[c], cb + c * 32, act_u32, row);
                        tc::tmem_ld_wait();
                        if (c + 2 < 4) tc::tmem_ld32(tmem_row + cb + (c + 2) * 32, va);
                        bwd_chunk32(vb, mk[c + 1], cb + (c + 1) * 32, act_u32, row);
                    }
                }
                const unsigned long long tw3 = g.prof ? clk() : 0;
                tc::tc_fence_before_sync();
                tc::fence_proxy_async_smem();
                named_bar(1, kEpiThreads);
                if (tid == 0) {
                    if (E.signal) signal_act_ready(tp, ctl, T);
                    if (g.prof) { t_wait += tw1 - tw0; t_busy += clk() - tw1; ++n_epi; t_pre += tw2 - tw1; t_work += tw3 - tw2; }
                    const int tile_s = g.save_alias ? tile_g % g.save_alias : tile_g;
                    if (tile_ok)
                        tc::bulk_s2g_hint(g.save + ((size_t)E.save_layer * n_tiles + tile_s) * kActBytes, act_all + T * kActBytes,
                                          E.save_bytes, pol_stream);
                    tc::bulk_commit();
                    ++stores;
                }
            }
        }
    }
    if (tid == 0) tc::bulk_wait_all<0>();
    if (g.prof && tid == 0) {
        unsigned long long* o = g.prof + (size_t)blockIdx.x * 8;
        o[3] = t_wait; o[4] = t_busy; o[5] = n_epi;
        o[6] = t_pre; o[7] = t_work;
    }
}

// the epilogue of (unit, T) is complete: the act buffer holds the next A operand, the accumulator is drained.
// Pair: both CTAs arrive on the leader's barrier (count 2), released at cluster scope.
__device__ __forceinline__ void signal_act_ready(const Topo& tp, SmemCtl* ctl, int T) {
    if (tp.pair) tc::mbar_arrive_cluster_addr(tp.act_ready_leader + (uint32_t)T * 8u);
    else tc::mbar_arrive(&ctl->act_ready[T]);
}

template <int PI>
__device__ void producer_role(const ChainArgs& g, SmemCtl* ctl, uint8_t* ring) {
    const Program& P = c_prog[PI];
    const int n_loads = P.n_loads;                         // multiple of kSlots
    // every CTA re-reads the 2.4 MB of packed weights for each work item: keep them in L2 against the streaming saves
    const uint64_t pol_w = tc::l2_policy_evict_last(), pol_in = tc::l2_policy_evict_first();
    uint32_t phase = 0, n = 0;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++n) {
        const uint8_t* enc = g.enc + (g.enc_mode == 2 ? (size_t)(blockIdx.x * 2u + (n & 1u)) : (size_t)item) * kEncItemBytes;
        if (PI == 0 && g.enc_mode) tc::mbar_wait(&ctl->enc_ready[n & 1u], (n >> 1) & 1u);      // the encoder warps wrote it
        int slot = 0;
        for (int i = 0; i < n_loads; ++i) {
            const Load L = P.loads[i];
            tc::mbar_wait(&ctl->empty[slot], phase ^ 1);
            const uint8_t* src = (L.kind == LOAD_W ? g.wimg : enc) + L.off;
            tc::mbar_expect_tx(&ctl->full[slot], L.bytes);
            tc::bulk_g2s_hint(ring + slot * kSlotBytes, src, L.bytes, &ctl->full[slot], L.kind == LOAD_W ? pol_w : pol_in);
            if (++slot == kSlots) { slot = 0; phase ^= 1; }
        }
    }
}

// ---- MMA issuer -------------------------------------------------------------------------------
// The whole warp walks the op table (uniform control flow); one elected lane issues the MMAs and commits.
struct Issuer {
    uint32_t base16, tmem, full0, empty0, acc0, act0;
    uint32_t ready_slot, ready_phase, act_phase;
    unsigned long long t_act, t_stage;      // cycles spent waiting (profiling)
    bool prof;

    __device__ __forceinline__ void wait_act(uint32_t tile) {
        const unsigned long long t0 = prof ? clk() : 0;
        tc::mbar_wait_u32(act0 + tile * 8u, (act_phase >> tile) & 1u);
        act_phase ^= 1u << tile;
        if (prof) t_act += clk() - t0;
    }
    __device__ __forceinline__ uint32_t wait_stage() {           // next ring stage in program order
        const uint32_t s = ready_slot;
        const unsigned long long t0 = prof ? clk() : 0;
        tc::mbar_wait_u32(full0 + s * 8u, ready_phase);
        if (prof) t_stage += clk() - t0;
        if (++ready_slot == kSlots) { ready_slot = 0; ready_phase ^= 1u; }
        return s;
    }
};

// One K = 256 layer part: for h in {0,1}: for tile in {0,1}: stages 4h..4h+3.  Tile 0 waits for each stage right
// before the two K = 16 MMAs that read it (waiting for all four up front left the last stage of a layer only ~700
// cycles between the release of its ring slot and its deadline -- less than one L2 bulk fetch: the issuer spent a
// third of its time waiting on the ring); tile 1 finds them resident and releases them.
__device__ __forceinline__ void issue_h_part(Issuer& S, uint32_t idesc, bool commit_acc) {
    uint32_t slot[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        // ---- tile 0: stage by stage ----
        if (h == 0) S.wait_act(0);
#pragma unroll
        for (int c = 4 * h; c < 4 * h + 4; ++c) {
            slot[c] = S.wait_stage();
            tc::tc_fence_after_sync();
            if (tc::elect_one()) {
                const uint64_t a = ((uint64_t)kHi128 << 32) | (uint64_t)(S.base16 + (uint32_t)(c / 2) * 1024u + (uint32_t)(c % 2) * 4u);
                const uint64_t b = ((uint64_t)kHi64 << 32) | (uint64_t)(S.base16 + (kRingOff >> 4) + slot[c] * (kSlotBytes >> 4));
                tc::mma_f16_ss(S.tmem, a, b, idesc, c == 0 ? 0u : 1u);
                tc::mma_f16_ss(S.tmem, a + 2, b + 2, idesc, 1u);
                if (c == 7 && commit_acc) tc::mma_commit_u32(S.acc0);
            }
            __syncwarp();
        }
        // ---- tile 1: the four stages are resident ----
        if (h == 0) S.wait_act(1);
        tc::tc_fence_after_sync();
        if (tc::elect_one()) {
            const uint32_t d = S.tmem + 256u;
#pragma unroll
            for (int c = 4 * h; c < 4 * h + 4; ++c) {
                const uint64_t a = ((uint64_t)kHi128 << 32) |
                                   (uint64_t)(S.base16 + (kActBytes >> 4) + (uint32_t)(c / 2) * 1024u + (uint32_t)(c % 2) * 4u);
                const uint64_t b = ((uint64_t)kHi64 << 32) | (uint64_t)(S.base16 + (kRingOff >> 4) + slot[c] * (kSlotBytes >> 4));
                tc::mma_f16_ss(d, a, b, idesc, c == 0 ? 0u : 1u);
                tc::mma_f16_ss(d, a + 2, b + 2, idesc, 1u);
                tc::mma_commit_u32(S.empty0 + slot[c] * 8u);
            }
            if (h == 1 && commit_acc) tc::mma_commit_u32(S.acc0 + 8u);
        }
        __syncwarp();
    }
}

template <int PI>
__device__ void mma_role(const ChainArgs& g, SmemCtl* ctl, uint32_t smem_base) {
    const Program& P = c_prog[PI];
    Issuer S;
    S.base16 = smem_base >> 4;
    S.tmem = ctl->tmem_base;
    S.full0 = tc::smem_u32(&ctl->full[0]);
    S.empty0 = tc::smem_u32(&ctl->empty[0]);
    S.acc0 = tc::smem_u32(&ctl->acc_full[0]);
    S.act0 = tc::smem_u32(&ctl->act_ready[0]);
    S.ready_slot = S.ready_phase = S.act_phase = 0;
    S.t_act = S.t_stage = 0;
    S.prof = g.prof != nullptr;
    const unsigned long long t_begin = clk();
    const int n_mmas = P.n_mmas;
    bool first_item = true;
    uint32_t n_item = 0;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++n_item) {
        for (int i = 0; i < n_mmas; ++i) {
            const uint4 q0 = *reinterpret_cast<const uint4*>(&P.mmas[i]);
            const uint4 q1 = *(reinterpret_cast<const uint4*>(&P.mmas[i]) + 1);
            const uint32_t nk16 = (q1.y >> 16) & 0xFFu, flags = q1.y >> 24;
            if (q1.w == OP_HPART) {
                issue_h_part(S, q1.x, (flags & F_COMMIT_ACC) != 0);
                continue;
            }
            const uint32_t tile = q1.z & 0xFFu, n_wait = (q1.z >> 8) & 0xFFu, rel0 = (q1.z >> 16) & 0xFFu, rel1 = q1.z >> 24;
            if ((flags & F_WAIT_ACT) || ((flags & F_WAIT_PREV) && !first_item)) S.wait_act(tile);
            for (uint32_t w = 0; w < n_wait; ++w) S.wait_stage();
            if (PI == 0 && (flags & F_ENC_DONE) && tc::elect_one()) tc::mbar_arrive(&ctl->enc_free[n_item & 1u]);
            tc::tc_fence_after_sync();
            if (tc::elect_one()) {
                const uint64_t a = ((uint64_t)q0.z << 32) | (uint64_t)(S.base16 + q0.x);
                const uint64_t b = ((uint64_t)q0.w << 32) | (uint64_t)(S.base16 + q0.y);
                const uint32_t d = S.tmem + (q1.y & 0xFFFFu);
                if (nk16 > 0) tc::mma_f16_ss(d, a, b, q1.x, (flags & F_FIRST) ? 0u : 1u);
                if (nk16 > 1) tc::mma_f16_ss(d, a + 2, b + 2, q1.x, 1u);
                if (rel0 != 0xFFu) tc::mma_commit_u32(S.empty0 + rel0 * 8u);
                if (rel1 != 0xFFu) tc::mma_commit_u32(S.empty0 + rel1 * 8u);
                if (flags & F_COMMIT_ACC) tc::mma_commit_u32(S.acc0 + tile * 8u);
            }
            __syncwarp();
        }
        first_item = false;
    }
    if (S.prof && (threadIdx.x & 31) == 0) {
        unsigned long long* o = g.prof + (size_t)blockIdx.x * 8;
        o[0] = clk() - t_begin;     // issuer: total
        o[1] = S.t_act;             // issuer: waiting for epilogues (act_ready)
        o[2] = S.t_stage;           // issuer: waiting for ring stages (weights / encoded blocks)
    }
}

__device__ void encoder_role(const ChainArgs& g, const Topo& tp, SmemCtl* ctl, int tid);

template <int PI>
__global__ void __launch_bounds__(PI == 0 ? kThreads : kThreads - kEncThreads, 1) mlp_tc_chain_kernel(const ChainArgs g) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* act_all = smem;
    uint8_t* ring = smem + kRingOff;
    SmemCtl* ctl = reinterpret_cast<SmemCtl*>(ring + kSlots * kSlotBytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0 && (tc::smem_u32(smem) & 1023u) != 0) {
        printf("ddnerf mlp_tc: dynamic shared memory is not 1024-byte aligned\n");
        __trap();
    }
    if (warp == 9 && lane == 0) {
        for (int s = 0; s < kSlots; ++s) { tc::mbar_init(&ctl->full[s], 1); tc::mbar_init(&ctl->empty[s], 1); }
        for (int t = 0; t < 2; ++t) { tc::mbar_init(&ctl->acc_full[t], 1); tc::mbar_init(&ctl->act_ready[t], 1); }
        for (int t = 0; t < 2; ++t) { tc::mbar_init(&ctl->enc_ready[t], 1); tc::mbar_init(&ctl->enc_free[t], 1); }
        tc::fence_barrier_init();
    }
    if (threadIdx.x == 0) ctl->epi_busy = 0u;
    if (warp == 8) tc::tmem_alloc(&ctl->tmem_base, 512);
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();

    const Topo tp{(int)blockIdx.x, (int)gridDim.x, g.n_items, 0, 0, g.n_items * 2, 0u};
    if (warp < 8) {
        if (PI == 0) epilogue_fwd(g, tp, ctl, act_all, warp, lane);
        else epilogue_bwd(g, tp, ctl, act_all, warp, lane);
    } else if (warp == 8) {
        if (lane == 0) producer_role<PI>(g, ctl, ring);
    } else if (warp == 9) {
        mma_role<PI>(g, ctl, tc::smem_u32(smem));
    } else {
        if (PI == 0 && g.enc_mode) encoder_role(g, tp, ctl, (int)threadIdx.x - 320);
    }

    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 8) tc::tmem_dealloc(ctl->tmem_base, 512);
}

// ---- CTA-pair variant (cta_group::2), see mlp_tc.cuh ------------------------------------------------------------
__constant__ PProgram c_pprog[2];

struct __align__(64) PairMaps { CUtensorMap m[kPairMaps]; };       // PM_W128, PM_W72, PM_W8 over the packed weights; PM_ENC over the images

// Stage order of one unit (all three roles walk it): for every epilogue e, for super-tile T in {0, 1}: the stages of e.
// The whole warp walks the table (uniform control flow, uniform constant loads); one elected lane issues the copies.
template <int PI>
__device__ void producer_pair(const ChainArgs& g, const Topo& tp, SmemCtl* ctl, uint8_t* ring, const PairMaps& maps) {
    const PProgram& P = c_pprog[PI];
    const uint64_t pol_w = tc::l2_policy_evict_last(), pol_in = tc::l2_policy_evict_first();
    const uint32_t full_leader = tc::mapa_u32(tc::smem_u32(&ctl->full[0]), 0);
    const uint32_t full_own = tc::smem_u32(&ctl->full[0]), empty0 = tc::smem_u32(&ctl->empty[0]);
    const uint32_t ring_u32 = tc::smem_u32(ring);
    const uint32_t rank = (uint32_t)tp.rank;
    uint32_t slot = 0, phase = 0, n = 0;
    const bool prof = g.prof != nullptr && tp.rank == 1;          // the peer's producer reports (its issuer slots are free)
    unsigned long long t_empty = 0, t_enc = 0;
    const unsigned long long t_begin = clk();
    for (int unit = tp.first; unit < tp.n_units; unit += tp.step, ++n) {
        if (PI == 0 && g.enc_mode) {                                // this CTA's encoder warps wrote its blocks
            const unsigned long long t0 = prof ? clk() : 0;
            tc::mbar_wait(&ctl->enc_ready[n & 1u], (n >> 1) & 1u);
            if (prof) t_enc += clk() - t0;
        }
        for (int e = 0; e < P.n_phases; ++e) {
            const int s1 = P.phase_begin[e + 1];
            const uint32_t fast = P.phase_fast[e];
            if (fast) {                 // plain layer: four stages of two N = 256 chunks each, once (shared) or per super-tile
                const uint32_t idx0 = P.phase_idx0[e] + rank;
                for (uint32_t T = 0; T < (fast == 2 ? 1u : 2u); ++T) {
#pragma unroll
                    for (uint32_t st = 0; st < 4; ++st) {
                        {
                            const unsigned long long t0 = prof ? clk() : 0;
                            tc::mbar_wait_u32(empty0 + slot * 8u, phase ^ 1u);
                            if (prof) t_empty += clk() - t0;
                        }
                        if (tc::elect_one()) {
                            if (rank == 0) {
                                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_own + slot * 8u), "r"(32768u) : "memory");
                            }
                            tc::tma_load_3d_pair(ring_u32 + slot * kSlotBytes, &maps.m[PM_W128], 0, 0, (int)(idx0 + 4u * st), full_leader + slot * 8u, pol_w);
                        }
                        __syncwarp();
                        if (++slot == (uint32_t)g.pslots) { slot = 0; phase ^= 1; }
                    }
                }
                continue;
            }
            for (int s = P.phase_begin[e]; s < s1;) {
                // a group: up to two shared stages (loaded once), or a run of stages of one super-tile (loaded for T0, then for T1)
                const bool shared = (P.st[s].flags & PF_SHARED) != 0;
                int ge = s + 1;
                if (shared) { if (ge < s1 && (P.st[ge].flags & PF_SHARED)) ++ge; }
                else { while (ge < s1 && !(P.st[ge].flags & PF_SHARED)) ++ge; }
                for (int T = 0; T < (shared ? 1 : 2); ++T) {
                    // image holding this CTA's encoded blocks of super-tile T (8 blocks of 8 KB each), and which tile of it they are
                    const uint32_t enc_blk0 = (g.enc_mode == 2 ? (blockIdx.x * 2u + (n & 1u)) : (uint32_t)(2 * unit + T)) * 8u;
                    const uint32_t tsel = g.enc_mode == 2 ? (uint32_t)T : rank;
                    for (int j = s; j < ge; ++j) {
                        const uint4 q0 = *reinterpret_cast<const uint4*>(&P.st[j]);
                        const uint4 q1 = *(reinterpret_cast<const uint4*>(&P.st[j]) + 1);       // the two copies
                        const uint32_t n_copies = (q0.y >> 8) & 0xFFu;
                        {
                            const unsigned long long t0 = prof ? clk() : 0;
                            tc::mbar_wait_u32(empty0 + slot * 8u, phase ^ 1u);
                            if (prof) t_empty += clk() - t0;
                        }
                        if (tc::elect_one()) {
                            if (rank == 0) {
                                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_own + slot * 8u), "r"(q0.w) : "memory");
                            }
                            const uint32_t dst = ring_u32 + slot * kSlotBytes, bar = full_leader + slot * 8u;
                            {
                                const uint32_t mul = q1.y & 0xFFu, map = (q1.y >> 8) & 0xFFu, off = q1.y >> 16;
                                const bool is_enc = map == PM_ENC;
                                const uint32_t idx = q1.x + (is_enc ? tsel * mul + enc_blk0 : rank);
                                tc::tma_load_3d_pair(dst + off, &maps.m[map], 0, 0, (int)idx, bar, is_enc ? pol_in : pol_w);
                            }
                            if (n_copies > 1) {
                                const uint32_t map = (q1.w >> 8) & 0xFFu, off = q1.w >> 16;      // the second copy is always a weight chunk
                                tc::tma_load_3d_pair(dst + off, &maps.m[map], 0, 0, (int)(q1.z + rank), bar, pol_w);
                            }
                        }
                        __syncwarp();
                        if (++slot == (uint32_t)g.pslots) { slot = 0; phase ^= 1; }
                    }
                }
                s = ge;
            }
        }
    }
    if (prof && (threadIdx.x & 31) == 0) {
        unsigned long long* o = g.prof + (size_t)blockIdx.x * 8;
        o[0] = clk() - t_begin; o[1] = t_empty; o[2] = t_enc;      // producer: total / waiting for free slots / for the encoder
    }
}

// MMA issuer of the pair (leader CTA).  The whole warp walks the program (uniform control flow keeps descriptors and barrier
// addresses in uniform registers, where tcgen05.mma wants them); one elected lane issues.  The regular layers (a K = 256,
// N = 256 accumulation over the act buffer: four two-chunk stages, 16 MMAs) run through a straight-line path with
// compile-time operand offsets.
struct PairIssuer {
    uint32_t base16, tmem, full0, empty0, acc0, act0;
    uint32_t slot, phase, act_phase;
    unsigned long long t_act, t_stage;
    bool prof;

    __device__ __forceinline__ void wait_act(uint32_t T) {            // both CTAs' epilogues of this super-tile are done
        const unsigned long long t0 = prof ? clk() : 0;
        tc::mbar_wait_u32(act0 + T * 8u, (act_phase >> T) & 1u);
        act_phase ^= 1u << T;
        if (prof) t_act += clk() - t0;
    }
    __device__ __forceinline__ void wait_stage() {                    // the current stage has landed in BOTH CTAs
        const unsigned long long t0 = prof ? clk() : 0;
        tc::mbar_wait_u32(full0 + slot * 8u, phase);
        if (prof) t_stage += clk() - t0;
    }
    uint32_t nslots;
    __device__ __forceinline__ void advance() {
        if (++slot == nslots) { slot = 0; phase ^= 1u; }
    }
};

template <int PI>
__device__ void mma_pair(const ChainArgs& g, const Topo& tp, SmemCtl* ctl, uint32_t smem_base) {
    const PProgram& P = c_pprog[PI];
    PairIssuer S;
    S.base16 = smem_base >> 4;
    S.tmem = ctl->tmem_base;
    S.full0 = tc::smem_u32(&ctl->full[0]);
    S.empty0 = tc::smem_u32(&ctl->empty[0]);
    S.acc0 = tc::smem_u32(&ctl->acc_full[0]);
    S.act0 = tc::smem_u32(&ctl->act_ready[0]);
    S.slot = S.phase = S.act_phase = 0;
    S.nslots = (uint32_t)g.pslots;
    S.t_act = S.t_stage = 0;
    S.prof = g.prof != nullptr;
    const uint32_t encf0 = tc::smem_u32(&ctl->enc_free[0]);
    const uint32_t ring16 = S.base16 + (kRingOff >> 4);
    const uint32_t idesc256 = tc::idesc_bf16(256, 256, 0, 0);
    const unsigned long long t_begin = clk();
    uint32_t n = 0;
    // the four K = 16 MMAs of a two-chunk stage of a plain layer: act chunks 2 st, 2 st + 1 against the two weight chunks in `slot`
    auto layer_stage = [&](uint32_t T, int st, uint32_t slot) {
        const uint32_t d = S.tmem + T * 256u;
        const uint64_t a = ((uint64_t)kHi128 << 32) | (uint64_t)(S.base16 + T * (kActBytes >> 4) + (uint32_t)st * 1024u);
        const uint64_t b = ((uint64_t)kHi64 << 32) | (uint64_t)(ring16 + slot * (kSlotBytes >> 4));
        tc::mma2_f16_ss(d, a, b, idesc256, st == 0 ? 0u : 1u);
        tc::mma2_f16_ss(d, a + 2, b + 2, idesc256, 1u);
        tc::mma2_f16_ss(d, a + 4, b + 512, idesc256, 1u);
        tc::mma2_f16_ss(d, a + 6, b + 514, idesc256, 1u);
    };
    for (int unit = tp.first; unit < tp.n_units; unit += tp.step, ++n) {
        for (int e = 0; e < P.n_phases; ++e) {
            const int s0 = P.phase_begin[e], s1 = P.phase_begin[e + 1];
            if (s0 == s1) continue;
            const uint32_t fast = P.phase_fast[e];
            if (fast == 2) {            // plain layer, shared stages: [T0: st 0, 1] [T1: st 0, 1] [T0: st 2, 3] [T1: st 2, 3]
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t sa = S.slot;
                    S.advance();
                    const uint32_t sb = S.slot;
                    if (h == 0) S.wait_act(0);
                    {
                        const unsigned long long t0 = S.prof ? clk() : 0;
                        tc::mbar_wait_u32(S.full0 + sa * 8u, sb == 0 ? S.phase ^ 1u : S.phase);      // (the phase bit flips when the slot index wraps)
                        if (S.prof) S.t_stage += clk() - t0;
                    }
                    tc::tc_fence_after_sync();
                    if (tc::elect_one()) layer_stage(0, 2 * h, sa);
                    __syncwarp();
                    S.wait_stage();                                                                // slot sb
                    tc::tc_fence_after_sync();
                    if (tc::elect_one()) {
                        layer_stage(0, 2 * h + 1, sb);
                        if (h == 1) tc::mma2_commit_u32(S.acc0);
                    }
                    __syncwarp();
                    if (h == 0) S.wait_act(1);
                    tc::tc_fence_after_sync();
                    if (tc::elect_one()) {
                        layer_stage(1, 2 * h, sa);
                        tc::mma2_commit_u32(S.empty0 + sa * 8u);                                   // frees the slot in both CTAs
                        layer_stage(1, 2 * h + 1, sb);
                        tc::mma2_commit_u32(S.empty0 + sb * 8u);
                        if (h == 1) tc::mma2_commit_u32(S.acc0 + 8u);
                    }
                    __syncwarp();
                    S.advance();
                }
                continue;
            }
            if (fast == 1) {            // plain layer, stages per super-tile
                for (uint32_t T = 0; T < 2; ++T) {
                    S.wait_act(T);
#pragma unroll
                    for (int st = 0; st < 4; ++st) {
                        S.wait_stage();
                        tc::tc_fence_after_sync();
                        if (tc::elect_one()) {
                            layer_stage(T, st, S.slot);
                            tc::mma2_commit_u32(S.empty0 + S.slot * 8u);
                            if (st == 3) tc::mma2_commit_u32(S.acc0 + T * 8u);
                        }
                        __syncwarp();
                        S.advance();
                    }
                }
                continue;
            }
            for (int s = s0; s < s1;) {
                const bool shared = (P.st[s].flags & PF_SHARED) != 0;
                int ge = s + 1;
                if (shared) { if (ge < s1 && (P.st[ge].flags & PF_SHARED)) ++ge; }
                else { while (ge < s1 && !(P.st[ge].flags & PF_SHARED)) ++ge; }
                const uint32_t slot_g = S.slot, phase_g = S.phase;       // ring position of the group's first stage
                for (uint32_t T = 0; T < 2; ++T) {
                    const uint32_t d = S.tmem + T * 256u;
                    const uint32_t act16 = S.base16 + T * (kActBytes >> 4);
                    if (shared) { S.slot = slot_g; S.phase = phase_g; }  // T1 walks the same slots again
                    for (int j = s; j < ge; ++j) {
                        const uint4 q0 = *reinterpret_cast<const uint4*>(&P.st[j]);      // kind..flags | a_chunk0, n_copies, b_stride | idesc | tx
                        const uint32_t kind = q0.x & 0xFFu, n_chunks = (q0.x >> 8) & 0xFFu, last_nk16 = (q0.x >> 16) & 0xFFu, flags = q0.x >> 24;
                        const uint32_t a_chunk0 = q0.y & 0xFFu, b_stride16 = (q0.y >> 16) >> 4, idesc = q0.z;
                        if ((flags & PF_WAIT_ACT) || ((flags & PF_WAIT_PREV) && n > 0)) S.wait_act(T);
                        if (!shared || T == 0) S.wait_stage();
                        tc::tc_fence_after_sync();
                        if (tc::elect_one()) {
                            const uint32_t slot16 = ring16 + S.slot * (kSlotBytes >> 4);
#pragma unroll 1
                            for (uint32_t c = 0; c < n_chunks; ++c) {
                                uint64_t a, b;
                                if (kind == PS_ENCW) {
                                    a = ((uint64_t)kHi64 << 32) | (uint64_t)slot16;
                                    b = ((uint64_t)kHi64 << 32) | (uint64_t)(slot16 + (8192u >> 4));
                                } else {
                                    const uint32_t ch = a_chunk0 + c;
                                    a = ((uint64_t)kHi128 << 32) | (uint64_t)(act16 + (ch >> 1) * 1024u + (ch & 1u) * 4u);
                                    b = ((uint64_t)kHi64 << 32) | (uint64_t)(slot16 + c * b_stride16);
                                }
                                tc::mma2_f16_ss(d, a, b, idesc, ((flags & PF_FIRST) && c == 0) ? 0u : 1u);
                                if (c + 1 < n_chunks || last_nk16 > 1) tc::mma2_f16_ss(d, a + 2, b + 2, idesc, 1u);
                            }
                            if (!shared || T == 1) tc::mma2_commit_u32(S.empty0 + S.slot * 8u);
                            if (flags & PF_COMMIT_ACC) tc::mma2_commit_u32(S.acc0 + T * 8u);
                            if (PI == 0 && (flags & PF_ENC_DONE) && T == 1) tc::mma2_commit_u32(encf0 + (n & 1u) * 8u);
                        }
                        __syncwarp();
                        S.advance();
                    }
                }
                s = ge;
            }
        }
    }
    if (S.prof && (threadIdx.x & 31) == 0) {
        unsigned long long* o = g.prof + (size_t)blockIdx.x * 8;
        o[0] = clk() - t_begin; o[1] = S.t_act; o[2] = S.t_stage;
    }
}

template <int PI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PI == 0 ? kThreads : kThreads - kEncThreads, 1)
mlp_tc_pair_kernel(const ChainArgs g, const __grid_constant__ PairMaps maps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* act_all = smem;
    uint8_t* ring = smem + kRingOff;
    SmemCtl* ctl = reinterpret_cast<SmemCtl*>(ring + kSlots * kSlotBytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)tc::cluster_ctarank();
    if (threadIdx.x == 0 && (tc::smem_u32(smem) & 1023u) != 0) {
        printf("ddnerf mlp_tc: dynamic shared memory is not 1024-byte aligned\n");
        __trap();
    }
    if (warp == 9 && lane == 0) {
        for (int s = 0; s < kSlots; ++s) { tc::mbar_init(&ctl->full[s], 1); tc::mbar_init(&ctl->empty[s], 1); }
        for (int t = 0; t < 2; ++t) { tc::mbar_init(&ctl->acc_full[t], 1); tc::mbar_init(&ctl->act_ready[t], 2); }
        for (int t = 0; t < 2; ++t) { tc::mbar_init(&ctl->enc_ready[t], 1); tc::mbar_init(&ctl->enc_free[t], 1); }
        tc::fence_barrier_init();
    }
    if (warp == 8 && lane == 0) {
        for (int m = 0; m < kPairMaps; ++m) tc::prefetch_tensormap(&maps.m[m]);
        ctl->epi_busy = 0u;
    }
    tc::cluster_sync();                              // the barriers of both CTAs exist before anyone signals them
    if (warp == 8) tc::tmem_alloc2(&ctl->tmem_base, 512);
    tc::tc_fence_before_sync();
    tc::cluster_sync();
    tc::tc_fence_after_sync();

    const int n_units = (g.n_items + 1) >> 1;
    const Topo tp{(int)(blockIdx.x >> 1), (int)(gridDim.x >> 1), n_units, 1, rank, g.n_items * 2,
                  tc::mapa_u32(tc::smem_u32(&ctl->act_ready[0]), 0)};
    if (warp < 8) {
        if (PI == 0) epilogue_fwd(g, tp, ctl, act_all, warp, lane);
        else epilogue_bwd(g, tp, ctl, act_all, warp, lane);
    } else if (warp == 8) {
        producer_pair<PI>(g, tp, ctl, ring, maps);
    } else if (warp == 9) {
        if (rank == 0) mma_pair<PI>(g, tp, ctl, tc::smem_u32(smem));
    } else {
        if (PI == 0 && g.enc_mode) encoder_role(g, tp, ctl, (int)threadIdx.x - 320);
    }

    tc::tc_fence_before_sync();
    tc::cluster_sync();                              // the leader's MMAs read the peer's shared memory and write its TMEM
    if (warp == 8) tc::tmem_dealloc2(ctl->tmem_base, 512);
}

// ---- weight packing: fp32 nn.Linear parameters -> bf16 stage images in program order -----------
struct PackArgs {
    ddnerf_mlp_params p;
    uint8_t* wimg;
    float* bias;
    int C;
};

// element (n, k) of the stage's B operand  (D[m, n] += A[m, k] . B[n, k])
__device__ __forceinline__ float pack_value(const PackArgs& a, const PackEntry& E, int n, int k) {
    const int ld = E.p == 0 ? 96 : (E.p == 5 ? 352 : 256);
    switch (E.kind) {
        case PK_FWD:                       // B[n][k] = W_p[n][k]
            return (n < 256 && k < ld) ? __ldg(a.p.w[E.p] + (size_t)n * ld + k) : 0.f;
        case PK_FWD_DIR:
            if (n < 128) return k < 283 ? __ldg(a.p.w[10] + (size_t)n * 283 + k) : 0.f;
            if (n == 128) return k < 256 ? __ldg(a.p.w[9] + k) : 0.f;
            return 0.f;
        case PK_FWD_HEADS:
            if (n < 3) return __ldg(a.p.w[11] + n * 128 + k);
            if (n < 5 && a.C == 6) return __ldg(a.p.w[12] + (n - 3) * 128 + k);
            return 0.f;
        case PK_BWD:                       // B[n][k] = W_p[k][n0 + n]  (dX = dZ . W)
            return k < 256 ? __ldg(a.p.w[E.p] + (size_t)k * ld + E.n0 + n) : 0.f;
        case PK_BWD_DIR:                   // k < 128: W_dir[k][n]; k == 128: w_alpha[n]  (n = feat index)
            if (k < 128) return __ldg(a.p.w[10] + (size_t)k * 283 + n);
            if (k == 128) return __ldg(a.p.w[9] + n);
            return 0.f;
        default:
            return 0.f;
    }
}

// One block per stage: the [n_total x 32 k] bf16 image is assembled in shared memory (element order chosen so that a warp
// reads consecutive fp32 parameters: along k for the forward stages, along n for the transposed stages of the dX program)
// and leaves as 16-byte stores.  Runs after every optimizer step (20 us -> 5 us per network).
__device__ __forceinline__ void pack_bias_row(const PackArgs& a, int l, int c);

// Blocks [0, n_stages) pack one weight stage each, blocks [n_stages, n_stages + kBiasRows) one row of the fp32 table
// (one launch per network and optimizer step instead of two).
__global__ void __launch_bounds__(256) pack_weights_kernel(const PackArgs a, int n_stages) {
    __shared__ __align__(16) uint8_t img[256 * 64];
    if ((int)blockIdx.x >= n_stages) { pack_bias_row(a, (int)blockIdx.x - n_stages, (int)threadIdx.x); return; }
    const PackEntry E = c_pack.e[blockIdx.x];
    const int total = E.n_total * 32;
    const bool along_n = E.kind == PK_BWD || E.kind == PK_BWD_DIR;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int n = along_n ? e % E.n_total : e >> 5, kk = along_n ? e / E.n_total : e & 31;
        *reinterpret_cast<__nv_bfloat16*>(img + tc::sw64_off(n, kk)) = __float2bfloat16_rn(pack_value(a, E, n, E.k0 + kk));
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(a.wimg + E.dst_off);
    for (int i = threadIdx.x; i < E.n_total * 4; i += blockDim.x) dst[i] = reinterpret_cast<const uint4*>(img)[i];
}

// bias table [16][256]: rows 0..8 = layers_xyz.0-7, fc_feat; row 9 = [layers_dir.0 (128) | fc_alpha];
// row 10 = [fc_rgb (3) | fc_mu_sigma (2)]; rows 11..13 = fc_rgb.weight rows, 14..15 = fc_mu_sigma.weight
// rows (fp32 copies at aligned addresses, read by the backward chain's first epilogue)
__device__ __forceinline__ void pack_bias_row(const PackArgs& a, int l, int c) {
    float v = 0.f;
    if (l <= 8) v = __ldg(a.p.b[l] + c);
    else if (l == 9) v = c < 128 ? __ldg(a.p.b[10] + c) : (c == 128 ? __ldg(a.p.b[9]) : 0.f);
    else if (l == 10) v = c < 3 ? __ldg(a.p.b[11] + c) : ((c < 5 && a.C == 6) ? __ldg(a.p.b[12] + c - 3) : 0.f);
    else if (l <= 13) v = c < 128 ? __ldg(a.p.w[11] + (l - 11) * 128 + c) : 0.f;
    else v = (c < 128 && a.C == 6) ? __ldg(a.p.w[12] + (l - 14) * 128 + c) : 0.f;
    a.bias[l * 256 + c] = v;
}

// ---- encoder writing bf16 operand images -------------------------------------------------------
// One thread per sample row: cone -> Gaussian, the 96 IPE features (three [128 x 32] SWIZZLE_64B blocks
// of the item image) and the 27 (+5 zero) view-direction features, stored as 16-byte chunks.  The
// values are rounded to bf16 (2^-9 relative), so the transcendental functions use the SFU: the argument
// follows the reference's fp32 steps (y = x 2^l, the floored remainder by fl32(100 pi) of safe_sin, the
// fp32 sum y + fl32(pi/2) of the cosine half, math_utils.py:128-166), then a two-term Cody-Waite
// reduction to [-pi, pi] and sin.approx (absolute error ~1e-6).
__device__ __forceinline__ float sfu_sin(float x) {                // |x| <= 100 pi
    const float k = rintf(x * 0.15915494309189535f);
    float r = fmaf(-k, 6.2831854820251465f, x);
    r = fmaf(-k, -1.7484555e-7f, r);
    return __sinf(r);
}
__device__ __forceinline__ float safe_arg_fast(float x) {          // math_utils.py:154-166
    const float T = 314.15927124f;
    if (fabsf(x) < T) return x;
    const float q = floorf(x * (1.0f / T));
    return fmaf(-q, T, x);
}

// one sample row of the operand image: `row` = global sample row, (T, r) = its tile and row inside the 256-row item at `ib`
// `busy` (encoder warps of the chain kernel): shared flag the epilogue warps raise while they work; the encoder yields the issue
// slots it shares with them between octaves (bounded wait), i.e. it runs in the gaps in which they wait for the tensor pipe
__device__ __forceinline__ void encode_row_image(const float* __restrict__ rays, const float* __restrict__ t_vals, int64_t N, int S,
                                                 int ray_shape, int64_t row, uint8_t* __restrict__ ib, int T, int r,
                                                 const volatile uint32_t* busy = nullptr) {
    uint32_t w[48];                       // 96 bf16: feature f = h*48 + l*3 + a in word f/2
    uint32_t dw[16];                      // 32 bf16 of the direction block
#pragma unroll
    for (int i = 0; i < 48; ++i) w[i] = 0u;
#pragma unroll
    for (int i = 0; i < 16; ++i) dw[i] = 0u;
    if (row < N * S) {
        const int64_t ray = row / S;
        const int i = (int)(row - ray * S);
        const RayGeom g = load_ray(rays, ray);
        const float* tp = t_vals + ray * (S + 1) + i;
        const Gauss3 s = cast_interval(g, __ldg(tp), __ldg(tp + 1), ray_shape);
        const float m[3] = {s.mx, s.my, s.mz}, c[3] = {s.cx, s.cy, s.cz};
        float f[96];
#pragma unroll
        for (int l = 0; l < 16; ++l) {
            const float scale = (float)(1 << l), sc2 = scale * scale;
            if (busy && (l & 1) == 0) {
                for (int spins = 0; *busy && spins < 48; ++spins) __nanosleep(64);
            }
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float y = m[a] * scale;
                const float e = __expf(-0.5f * (c[a] * sc2));
                f[l * 3 + a] = e * sfu_sin(safe_arg_fast(y));
                f[48 + l * 3 + a] = e * sfu_sin(safe_arg_fast(y + 1.57079637f));
            }
        }
#pragma unroll
        for (int i = 0; i < 48; ++i) w[i] = tc::pack_bf16(f[2 * i], f[2 * i + 1]);
        float d[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) d[i] = 0.f;
        const float v[3] = {g.vx, g.vy, g.vz};
#pragma unroll
        for (int a = 0; a < 3; ++a) {                    // [v | sin v, cos v | sin 2v, cos 2v | sin 4v, cos 4v | sin 8v, cos 8v]
            d[a] = v[a];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float x = v[a] * (float)(1 << k);  // |x| <= 8
                d[3 + 6 * k + a] = __sinf(x);
                d[6 + 6 * k + a] = __cosf(x);
            }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) dw[i] = tc::pack_bf16(d[2 * i], d[2 * i + 1]);
    }
    const uint32_t sw = ((uint32_t)r >> 1) & 3u;
#pragma unroll
    for (int b = 0; b < 3; ++b) {                        // xyz block b = features 32 b .. 32 b + 31
        uint8_t* bp = ib + (b < 2 ? (uint32_t)T * 16384u + (uint32_t)b * 8192u : 32768u + (uint32_t)T * 8192u) + (uint32_t)r * 64u;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(bp + (((uint32_t)j ^ sw) << 4)) =
                make_uint4(w[16 * b + 4 * j], w[16 * b + 4 * j + 1], w[16 * b + 4 * j + 2], w[16 * b + 4 * j + 3]);
    }
    uint8_t* dp = ib + 49152u + (uint32_t)T * 8192u + (uint32_t)r * 64u;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(dp + (((uint32_t)j ^ sw) << 4)) = make_uint4(dw[4 * j], dw[4 * j + 1], dw[4 * j + 2], dw[4 * j + 3]);
}

__global__ void __launch_bounds__(128) encode_img_kernel(const float* __restrict__ rays, const float* __restrict__ t_vals,
                                                         uint8_t* __restrict__ img, int64_t N, int S, int ray_shape,
                                                         int64_t rows_padded) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows_padded) return;
    const int64_t item = row / kItemRows;
    encode_row_image(rays, t_vals, N, S, ray_shape, row, img + item * kEncItemBytes, (int)(row % kItemRows) / 128, (int)(row % 128));
}

// The encoder warps of the forward chain kernel (2 warps, 4 sample rows per thread and item).  They run one to two work
// items ahead of the GEMMs: item n of this CTA goes to image slot (n & 1) as soon as the MMA warp has seen the last encoded
// block of item n - 2 land in the ring, the producer starts item n's loads when the image is complete.  The image is
// written with ordinary stores and read back by the TMA engine (bulk copies): fence.proxy.async orders the two proxies,
// the named barrier + mbarrier arrive / wait carry the release / acquire.
__device__ void encoder_role(const ChainArgs& g, const Topo& tp, SmemCtl* ctl, int tid) {
    uint32_t n = 0;
    const volatile uint32_t* busy = g.enc_yield ? &ctl->epi_busy : nullptr;
    for (int unit = tp.first; unit < tp.n_units; unit += tp.step, ++n) {
        const uint32_t par = n & 1u;
        if (n >= 2) tc::mbar_wait(&ctl->enc_free[par], ((n >> 1) - 1u) & 1u);
        uint8_t* scratch = const_cast<uint8_t*>(g.enc) + (size_t)(blockIdx.x * 2u + par) * kEncItemBytes;
#pragma unroll 1
        for (int rr = tid; rr < kItemRows; rr += kEncThreads) {
            const int T = rr >> 7, r = rr & 127;
            if (!tp.pair) {              // this CTA's 256-row item: tile T of the item image
                uint8_t* ib = g.enc_mode == 2 ? scratch : const_cast<uint8_t*>(g.enc) + (size_t)unit * kEncItemBytes;
                encode_row_image(g.rays, g.t_vals, g.N, g.S, g.ray_shape, (int64_t)unit * kItemRows + rr, ib, T, r, busy);
            } else {                     // rows [128 rank, +128) of super-tile T = tile `rank` of 256-row item 2 unit + T
                const int item = 2 * unit + T;
                const int64_t row = ((int64_t)item * 2 + tp.rank) * 128 + r;
                if (g.enc_mode == 2) encode_row_image(g.rays, g.t_vals, g.N, g.S, g.ray_shape, row, scratch, T, r, busy);
                else if (item < g.n_items)
                    encode_row_image(g.rays, g.t_vals, g.N, g.S, g.ray_shape, row, const_cast<uint8_t*>(g.enc) + (size_t)item * kEncItemBytes,
                                     tp.rank, r, busy);
            }
        }
        tc::fence_proxy_async_all();
        named_bar(2, kEncThreads);
        if (tid == 0) tc::mbar_arrive(&ctl->enc_ready[par]);
    }
}

// ---- host: program construction ----------------------------------------------------------------

struct RawMma {
    int tile, flags, nk16, n;            // n: MMA N (instruction descriptor)
    int a_seq;                           // ring stage holding A, -1: the tile's act buffer
    uint32_t a_off;
    int b_seq;                           // ring stage holding B (-1: none, nk16 == 0)
    uint32_t b_off;
    int rel0, rel1;                      // ring stages released after this group
    int kind = OP_GENERIC;               // OP_HPART: b_seq = first of 8 consecutive stages
};

struct Builder {
    Program P{};
    PackTable* K = nullptr;
    std::vector<RawMma> ops;
    int seq = 0;
    uint32_t woff = 0, wbase = 0;        // wbase: byte offset of this program's region in the weight image

    int load_w(uint32_t bytes, uint16_t kind, uint16_t p, uint16_t n_total, uint16_t k0, uint16_t n0 = 0, int n_chunks = 1) {
        for (int c = 0; c < n_chunks; ++c)
            K->e[K->n++] = PackEntry{wbase + woff + (uint32_t)c * n_total * 64u, kind, p, n_total, (uint16_t)(k0 + 32 * c), n0, 0};
        P.loads[P.n_loads++] = Load{LOAD_W, woff, bytes};
        woff += bytes;
        return seq++;
    }
    int load_enc(uint32_t off, uint32_t bytes) {
        P.loads[P.n_loads++] = Load{LOAD_ENC, off, bytes};
        return seq++;
    }
    void mma(int tile, int flags, int nk16, int n, int a_seq, uint32_t a_off, int b_seq, uint32_t b_off, int rel0 = -1,
             int rel1 = -1) {
        ops.push_back(RawMma{tile, flags, nk16, n, a_seq, a_off, b_seq, b_off, rel0, rel1, OP_GENERIC});
    }
    // the 16 MMA groups an OP_HPART entry stands for (issue_h_part executes exactly this order)
    static void expand_h_part(const RawMma& hp, std::vector<RawMma>& out) {
        for (int h = 0; h < 2; ++h)
            for (int T = 0; T < 2; ++T)
                for (int c = 4 * h; c < 4 * h + 4; ++c) {
                    int flags = 0;
                    if (c == 0) flags |= F_FIRST | F_WAIT_ACT;
                    if (c == 7 && (hp.flags & F_COMMIT_ACC)) flags |= F_COMMIT_ACC;
                    out.push_back(RawMma{T, flags, 2, 256, -1, (c / 2) * 16384u + (c % 2) * 64u, hp.b_seq + c, 0,
                                         T == 1 ? hp.b_seq + c : -1, -1, OP_GENERIC});
                }
    }
    void epi(int mode, int relu, int save_layer, int mask_layer, int ncols, int bias_row, uint32_t save_bytes, int signal = 1) {
        P.epis[P.n_epis++] = Epi{(uint8_t)mode, (uint8_t)relu, (int8_t)save_layer, (int8_t)mask_layer, (uint16_t)ncols,
                                 (uint16_t)(bias_row * 256), save_bytes, (uint32_t)signal};
    }
    // the encoded xyz part (K = 96) of layers 0 and 5: A operand from the item image, 3 weight stages
    void xyz_part(int p, bool first_layer) {
        const int x0 = load_enc(0, 16384), x1 = load_enc(16384, 16384), x2 = load_enc(32768, 16384);
        int w[3];
        for (int c = 0; c < 3; ++c) w[c] = load_w(16384, PK_FWD, p, 256, 32 * c);
        for (int T = 0; T < 2; ++T)
            for (int c = 0; c < 3; ++c) {
                const int a_seq = c < 2 ? (T == 0 ? x0 : x1) : x2;
                const uint32_t a_off = c < 2 ? c * 8192u : T * 8192u;
                int flags = 0;
                if (first_layer && c == 0) flags |= F_FIRST | F_WAIT_PREV;
                if (c == 2) flags |= F_COMMIT_ACC;
                int rel0 = -1, rel1 = -1;
                if (T == 0) { if (c == 1) rel0 = x0; }
                else { rel0 = w[c]; if (c == 1) rel1 = x1; if (c == 2) rel1 = x2; }
                mma(T, flags, 2, 256, a_seq, a_off, w[c], 0, rel0, rel1);
            }
    }
    // a K = 256 accumulation over the act buffer (8 stages), tiles interleaved half a layer apart
    void h_part(uint16_t kind, int p, int k0, int n0, bool commit) {
        int w[8];
        for (int c = 0; c < 8; ++c) w[c] = load_w(16384, kind, p, 256, k0 + 32 * c, n0);
        ops.push_back(RawMma{0, commit ? F_COMMIT_ACC : 0, 0, 256, -1, 0, w[0], 0, -1, -1, OP_HPART});
    }
    // pad the stage count to a multiple of the ring size (static slot indices), resolve the op table
    bool finish(char* why, size_t n) {
        std::vector<int> dummies;
        while (seq % kSlots) {
            P.loads[P.n_loads++] = Load{LOAD_W, 0, 16};
            dummies.push_back(seq++);
        }
        for (size_t i = 0; i < dummies.size(); i += 2)
            mma(0, 0, 0, 256, -1, 0, dummies[i], 0, dummies[i], i + 1 < dummies.size() ? dummies[i + 1] : -1);
        if (P.n_loads > kMaxLoads || (int)ops.size() > kMaxMmas) { snprintf(why, n, "program too large"); return false; }
        int ready = 0;
        bool released[kMaxLoads] = {};
        for (size_t i = 0; i < ops.size(); ++i) {
            // schedule check on the expanded groups: every stage <= hi - kSlots must have been released by
            // earlier groups, else the producer (which refills slots in order) and the issuer deadlock
            std::vector<RawMma> sub;
            if (ops[i].kind == OP_HPART) expand_h_part(ops[i], sub); else sub.push_back(ops[i]);
            const int ready_before = ready;
            for (const RawMma& o : sub) {
                int hi = std::max(o.a_seq, o.b_seq);
                if (o.rel1 > hi && o.nk16 == 0) hi = o.rel1;
                for (int s = 0; s <= hi - kSlots; ++s)
                    if (!released[s]) { snprintf(why, n, "mma group %zu needs stage %d but stage %d is still held", i, hi, s); return false; }
                ready = std::max(ready, hi + 1);
                if (o.rel0 >= 0) released[o.rel0] = true;
                if (o.rel1 >= 0) released[o.rel1] = true;
            }
            const RawMma& o = ops[i];
            Mma m{};
            m.kind = (uint32_t)o.kind;
            m.idesc = tc::idesc_bf16(128, o.n, 0, 0);
            m.flags = (uint8_t)o.flags;
            if (o.kind == OP_HPART) {
                // the specialised routine takes its stages from the running ring position
                if (ready_before != o.b_seq || ready != o.b_seq + 8) { snprintf(why, n, "h-part %zu does not start at the ring position", i); return false; }
                P.mmas[P.n_mmas++] = m;
                continue;
            }
            const int hi = ready - 1;
            const uint32_t a_byte = o.a_seq >= 0 ? kRingOff + (uint32_t)(o.a_seq % kSlots) * kSlotBytes + o.a_off
                                                 : (uint32_t)o.tile * kActBytes + o.a_off;
            const uint32_t b_byte = kRingOff + (uint32_t)((o.b_seq < 0 ? 0 : o.b_seq) % kSlots) * kSlotBytes + o.b_off;
            m.a_lo = a_byte >> 4;
            m.b_lo = b_byte >> 4;
            m.a_hi = o.a_seq >= 0 ? kHi64 : kHi128;
            m.b_hi = kHi64;
            m.d_col = (uint16_t)(o.tile * 256);
            m.nk16 = (uint8_t)o.nk16;
            m.tile = (uint8_t)o.tile;
            m.n_wait = (uint8_t)std::max(0, hi + 1 - ready_before);
            m.rel0 = o.rel0 >= 0 ? (uint8_t)(o.rel0 % kSlots) : 0xFF;
            m.rel1 = o.rel1 >= 0 ? (uint8_t)(o.rel1 % kSlots) : 0xFF;
            P.mmas[P.n_mmas++] = m;
        }
        if (ready != seq) { snprintf(why, n, "%d stages loaded but %d consumed", seq, ready); return false; }
        for (int s = 0; s < seq; ++s)
            if (!released[s]) { snprintf(why, n, "stage %d is never released", s); return false; }
        return true;
    }
};


// ---- CTA-pair programs: the same packed weight image, walked per super-tile ------------------------------------
struct PairBuilder {
    PProgram P{};
    std::vector<uint32_t> woff;      // byte offsets of the weight stages in the packed image, in load order
    size_t wi = 0;
    bool ok = true;
    char why[160] = "";

    uint32_t next_w(uint32_t bytes, const std::vector<uint32_t>& sizes) {
        if (wi >= woff.size() || sizes[wi] != bytes) { ok = false; snprintf(why, sizeof(why), "pair program: weight stage %zu mismatch", wi); return 0; }
        return woff[wi++];
    }
    PStage& add(uint8_t kind, int n_chunks, int last_nk16, int flags, int a_chunk0, int n, uint16_t b_stride) {
        PStage& S = P.st[P.n_stages++];
        S = PStage{};
        S.kind = kind; S.n_chunks = (uint8_t)n_chunks; S.last_nk16 = (uint8_t)last_nk16; S.flags = (uint8_t)flags;
        S.a_chunk0 = (uint8_t)a_chunk0; S.b_stride = b_stride; S.idesc = tc::idesc_bf16(256, n, 0, 0);
        return S;
    }
    static void copy(PStage& S, uint8_t map, uint32_t idx, uint8_t mul, uint16_t dst_off, uint32_t bytes_per_cta) {
        PCopy& C = S.c[S.n_copies++];
        C = PCopy{idx, mul, map, dst_off};
        S.tx_bytes += 2 * bytes_per_cta;
    }
};

// `single` = the finished single-CTA builder of the same program (source of the weight offsets and the epilogue list)
bool build_pair(const Builder& single, bool fwd, PProgram& out, uint32_t (&bases)[3], char* why, size_t n) {
    PairBuilder b;
    std::vector<uint32_t> sizes;
    for (int i = 0; i < single.P.n_loads; ++i)
        if (single.P.loads[i].kind == LOAD_W && single.P.loads[i].bytes > 16) {
            b.woff.push_back(single.wbase + single.P.loads[i].off);
            sizes.push_back(single.P.loads[i].bytes);
        }
    // The N = 256 chunks of a program are contiguous from `w128_base` (16 KB each), the view-branch chunks from `w72_base`
    // (9216 B each); a copy with the traversal stride of the maps brings this CTA's halves of chunks c and c + 1; one-chunk
    // stages use the single-half maps.
    uint32_t w128_base = 0xFFFFFFFFu, w72_base = 0xFFFFFFFFu, w8_base = 0;
    for (size_t i = 0; i < b.woff.size(); ++i) {
        if (sizes[i] == 16384 && w128_base == 0xFFFFFFFFu) w128_base = b.woff[i];
        if (sizes[i] == 9216 && w72_base == 0xFFFFFFFFu) w72_base = b.woff[i];
        if (sizes[i] == 4096) w8_base = b.woff[i];
    }
    // Default: every stage per super-tile, the super-tiles a whole layer apart.  DDNERF_TC_PAIR_SHARE=1 selects the shared-stage
    // schedule (half a layer apart, half the L2 -> SM traffic): measured 10 % slower at inference (the epilogue no longer fits
    // behind the other super-tile's MMAs), equal in training.
    const char* share_env = getenv("DDNERF_TC_PAIR_SHARE");
    const int share = (share_env && atoi(share_env) == 1) ? PF_SHARED : 0;
    auto w128x2 = [&](PStage& S, int nc, uint16_t dst) {          // nc chunks (1 or 2) starting at the next weight stage
        const uint32_t off = b.next_w(16384, sizes);
        if (nc > 1) b.next_w(16384, sizes);
        if ((off - w128_base) % 16384) { b.ok = false; snprintf(b.why, sizeof(b.why), "pair program: N = 256 chunks are not contiguous"); }
        PairBuilder::copy(S, nc > 1 ? PM_W128 : PM_W128S, 2 * ((off - w128_base) / 16384), 1, dst, nc > 1 ? 16384 : 8192);
    };
    auto w72x2 = [&](PStage& S, int nc, uint16_t dst) {
        const uint32_t off = b.next_w(9216, sizes);
        if (nc > 1) b.next_w(9216, sizes);
        if ((off - w72_base) % 9216) { b.ok = false; snprintf(b.why, sizeof(b.why), "pair program: view-branch chunks are not contiguous"); }
        PairBuilder::copy(S, nc > 1 ? PM_W72 : PM_W72S, 2 * ((off - w72_base) / 9216), 1, dst, nc > 1 ? 9216 : 4608);
    };
    // encoded 8 KB blocks of tile `tsel` inside a 64 KB item image: b0, b1 = blocks 2 tsel, 2 tsel + 1; b2 = 4 + tsel; view directions 6 + tsel
    auto enc = [&](PStage& S, int blk) {
        const uint32_t idx = blk < 2 ? (uint32_t)blk : (blk == 2 ? 4u : 6u);
        PairBuilder::copy(S, PM_ENC, idx, blk < 2 ? 2 : 1, 0, 8192);
    };
    auto act_layer = [&](int n_chunks_total, int nmma, bool first, bool commit, int wait_flag, bool dir) {      // K = 32 n_chunks_total over the act buffer
        for (int c = 0; c < n_chunks_total; c += 2) {
            const int nc = std::min(2, n_chunks_total - c);
            int flags = 0;
            if (c == 0 && first) flags |= PF_FIRST | wait_flag;
            if (c + 2 >= n_chunks_total && commit) flags |= PF_COMMIT_ACC;
            PStage& S = b.add(PS_ACT, nc, 2, flags | share, c, nmma, dir ? 4608 : 8192);
            if (dir) w72x2(S, nc, 0); else w128x2(S, nc, 0);
        }
    };
    auto xyz_part = [&](bool first, int wait_flag) {       // K = 96 over the encoded xyz blocks, commits the accumulator
        for (int c = 0; c < 3; ++c) {
            int flags = 0;
            if (c == 0 && first) flags |= PF_FIRST | wait_flag;
            if (c == 2) flags |= PF_COMMIT_ACC;
            PStage& S = b.add(PS_ENCW, 1, 2, flags, 0, 256, 8192);
            enc(S, c);
            w128x2(S, 1, 8192);
        }
    };
    int e = 0;
    auto begin_phase = [&] { b.P.phase_begin[e++] = (uint16_t)b.P.n_stages; };
    if (fwd) {
        begin_phase(); xyz_part(true, PF_WAIT_PREV);                                   // layers_xyz.0
        for (int l = 1; l <= 8; ++l) {
            begin_phase();
            if (l == 5) { act_layer(8, 256, true, false, PF_WAIT_ACT, false); xyz_part(false, 0); }
            else act_layer(8, 256, true, true, PF_WAIT_ACT, false);
        }
        begin_phase();                                                                  // view branch + density, N = 144
        act_layer(8, 144, true, false, PF_WAIT_ACT, true);
        {
            PStage& S = b.add(PS_ENCW, 1, 2, PF_COMMIT_ACC | PF_ENC_DONE, 0, 144, 8192);
            enc(S, 3);
            w72x2(S, 1, 8192);
        }
        begin_phase();                                                                  // heads, N = 16: four [16 x 32] chunks, 8 rows per CTA
        {
            PStage& S = b.add(PS_HEADS, 4, 2, PF_FIRST | PF_WAIT_ACT | PF_COMMIT_ACC | share, 0, 16, 512);
            b.next_w(4096, sizes);
            PairBuilder::copy(S, PM_W8, 0, 1, 0, 2048);
        }
    } else {
        begin_phase();                                                                  // EPI_BWD_IN: no MMAs
        begin_phase();                                                                  // dZ_feat: K = 144 (4.5 chunks)
        for (int c = 0; c < 5; c += 2) {
            const int nc = std::min(2, 5 - c);
            int flags = 0;
            if (c == 0) flags |= PF_FIRST | PF_WAIT_ACT;
            if (c == 4) flags |= PF_COMMIT_ACC;
            PStage& S = b.add(PS_ACT, nc, c == 4 ? 1 : 2, flags | share, c, 256, 8192);
            w128x2(S, nc, 0);
        }
        for (int p = 8; p >= 1; --p) { begin_phase(); act_layer(8, 256, true, true, PF_WAIT_ACT, false); }
    }
    b.P.phase_begin[e] = (uint16_t)b.P.n_stages;
    b.P.n_phases = e;
    for (int ph = 0; ph < e; ++ph) {        // plain layers take the issuer's unrolled path
        const int s0 = b.P.phase_begin[ph], s1 = b.P.phase_begin[ph + 1];
        bool fast = s1 - s0 == 4;
        for (int s2 = s0; fast && s2 < s1; ++s2) {
            const PStage& S = b.P.st[s2];
            const int want = (s2 == s0 ? (PF_FIRST | PF_WAIT_ACT) : 0) | (s2 == s1 - 1 ? PF_COMMIT_ACC : 0) | share;
            fast = S.kind == PS_ACT && S.n_chunks == 2 && S.last_nk16 == 2 && S.a_chunk0 == 2 * (s2 - s0) && S.b_stride == 8192 &&
                   S.flags == want && S.idesc == tc::idesc_bf16(256, 256, 0, 0);
        }
        b.P.phase_fast[ph] = fast ? (share ? 2 : 1) : 0;
        b.P.phase_idx0[ph] = b.P.st[s0].c[0].idx;
        for (int s2 = s0; fast && s2 < s1; ++s2)
            if (b.P.st[s2].n_copies != 1 || b.P.st[s2].c[0].map != PM_W128 || b.P.st[s2].c[0].idx != b.P.st[s0].c[0].idx + 4u * (s2 - s0) ||
                b.P.st[s2].tx_bytes != 32768) { b.ok = false; snprintf(b.why, sizeof(b.why), "pair program: phase %d is not a plain layer", ph); }
    }
    if (b.ok && b.wi != b.woff.size()) { b.ok = false; snprintf(b.why, sizeof(b.why), "pair program: %zu of %zu weight stages used", b.wi, b.woff.size()); }
    if (b.ok && e != single.P.n_epis) { b.ok = false; snprintf(b.why, sizeof(b.why), "pair program: %d phases for %d epilogues", e, single.P.n_epis); }
    if (b.ok && b.P.n_stages > kMaxPStages) { b.ok = false; snprintf(b.why, sizeof(b.why), "pair program too large"); }
    if (!b.ok) { snprintf(why, n, "%s", b.why); return false; }
    out = b.P;
    if (fwd) { bases[0] = w128_base; bases[1] = w72_base; bases[2] = w8_base; }
    else { bases[0] = w128_base; bases[1] = bases[2] = w128_base; }
    return true;
}

struct Programs {
    Builder fwd, bwd;
    PProgram pfwd{}, pbwd{};
    uint32_t pbase[2][3] = {};       // byte offsets of the pair maps' weight regions (N = 256, view branch, heads), per program
    PackTable pack{};
    bool ok = false;
    char why[160] = "";
};

void build_fwd(Builder& b) {
    // layer 0: xyz (96) -> 256
    b.xyz_part(0, true);
    b.epi(EPI_ACT, 1, 0, -1, 256, 0, kActBytes);
    for (int l = 1; l <= 8; ++l) {          // layers_xyz.1-7 and fc_feat (l == 8, no activation)
        if (l == 5) {                       // cat(xyz, h): weight columns 0..95 = xyz, 96..351 = h
            b.h_part(PK_FWD, 5, 96, 0, false);
            b.xyz_part(5, false);
        } else {
            b.h_part(PK_FWD, l, 0, 0, true);
        }
        b.epi(EPI_ACT, l < 8 ? 1 : 0, l, -1, 256, l, kActBytes);
    }
    {   // view branch + density: N = 144 = [layers_dir.0 (128) | fc_alpha | 15 x 0], K = 256 feat + 32 dir
        int w[9];
        for (int c = 0; c < 8; ++c) w[c] = b.load_w(144 * 64, PK_FWD_DIR, 10, 144, 32 * c);
        const int d = b.load_enc(49152, 16384);
        w[8] = b.load_w(144 * 64, PK_FWD_DIR, 10, 144, 256);
        for (int h = 0; h < 2; ++h)
            for (int T = 0; T < 2; ++T) {
                for (int c = 4 * h; c < 4 * h + 4; ++c)
                    b.mma(T, c == 0 ? (F_FIRST | F_WAIT_ACT) : 0, 2, 144, -1, (c / 2) * 16384u + (c % 2) * 64u, w[c], 0, T == 1 ? w[c] : -1);
                if (h == 1) b.mma(T, F_COMMIT_ACC | (T == 0 ? F_ENC_DONE : 0), 2, 144, d, T * 8192u, w[8], 0, T == 1 ? w[8] : -1, T == 1 ? d : -1);
            }
        b.epi(EPI_DIR, 1, 9, -1, 144, 9, 32768);
    }
    {   // colour (+ mu, sigma) heads: N = 16, K = 128 (the view-branch activations, k-blocks 0..1)
        const int wl = b.load_w(4096, PK_FWD_HEADS, 11, 16, 0, 0, 4);
        for (int T = 0; T < 2; ++T)
            for (int c = 0; c < 4; ++c) {
                int flags = 0;
                if (c == 0) flags |= F_FIRST | F_WAIT_ACT;
                if (c == 3) flags |= F_COMMIT_ACC;
                b.mma(T, flags, 2, 16, -1, (c / 2) * 16384u + (c % 2) * 64u, wl, c * 1024u, (T == 1 && c == 3) ? wl : -1);
            }
        b.epi(EPI_OUT, 0, -1, -1, 16, 10, 0);
    }
}

void build_bwd(Builder& b) {
    // dZ_dir (+ density column) from the output cotangents, on CUDA cores
    b.epi(EPI_BWD_IN, 0, 9, 9, 144, 0, 49152);
    {   // dZ_feat = [dZ_dir | g_density | 0] (K = 144) . [W_dir[:, :256] ; w_alpha]
        int w[5];
        for (int c = 0; c < 5; ++c) w[c] = b.load_w(16384, PK_BWD_DIR, 10, 256, 32 * c);
        for (int h = 0; h < 2; ++h)
            for (int T = 0; T < 2; ++T)
                for (int c = (h == 0 ? 0 : 3); c < (h == 0 ? 3 : 5); ++c) {
                    int flags = 0;
                    if (c == 0) flags |= F_FIRST | F_WAIT_ACT;
                    if (c == 4) flags |= F_COMMIT_ACC;
                    b.mma(T, flags, c < 4 ? 2 : 1, 256, -1, (c / 2) * 16384u + (c % 2) * 64u, w[c], 0, T == 1 ? w[c] : -1);
                }
        b.epi(EPI_BWD_MASK, 0, 8, -1, 256, 0, kActBytes);           // fc_feat has no activation
    }
    for (int p = 8; p >= 1; --p) {          // dZ_{p-1} = (dZ_p . W_p) * relu'_{p-1}; layer 5 skips its xyz columns
        b.h_part(PK_BWD, p, 0, p == 5 ? 96 : 0, true);
        b.epi(EPI_BWD_MASK, 0, p - 1, p - 1, 256, 0, kActBytes, p > 1 ? 1 : 0);
    }
}

void build_into(Programs& S) {
    S.fwd.K = &S.pack;
    S.bwd.K = &S.pack;
    build_fwd(S.fwd);
    if (!S.fwd.finish(S.why, sizeof(S.why))) return;
    S.bwd.wbase = S.fwd.woff;
    build_bwd(S.bwd);
    if (!S.bwd.finish(S.why, sizeof(S.why))) return;
    if (S.pack.n > kMaxPack) { snprintf(S.why, sizeof(S.why), "pack table too large"); return; }
    S.pack.total_bytes = S.fwd.woff + S.bwd.woff;
    if (!build_pair(S.fwd, true, S.pfwd, S.pbase[0], S.why, sizeof(S.why))) return;
    if (!build_pair(S.bwd, false, S.pbwd, S.pbase[1], S.why, sizeof(S.why))) return;
    S.ok = true;
}

Programs* build_programs() {            // host tables are built exactly once
    static Programs* S = [] { Programs* s = new Programs(); build_into(*s); return s; }();
    return S;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_tiled = nullptr;
int g_pair_mode = -1;                   // -1: from DDNERF_TC_PAIR (default on); 0 single-CTA kernels; 1 CTA-pair kernels

std::once_flag g_once;
Programs* g_programs = nullptr;
int g_upload_rc = 0;

int ensure_programs() {
    std::call_once(g_once, [] {
        g_programs = build_programs();
        if (!g_programs->ok) { g_upload_rc = 1; return; }
        if (cudaMemcpyToSymbol(c_prog, &g_programs->fwd.P, sizeof(Program), 0) != cudaSuccess) g_upload_rc = 2;
        if (cudaMemcpyToSymbol(c_prog, &g_programs->bwd.P, sizeof(Program), sizeof(Program)) != cudaSuccess) g_upload_rc = 2;
        if (cudaMemcpyToSymbol(c_pack, &g_programs->pack, sizeof(PackTable)) != cudaSuccess) g_upload_rc = 2;
        if (cudaFuncSetAttribute(mlp_tc_chain_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess) g_upload_rc = 3;
        if (cudaFuncSetAttribute(mlp_tc_chain_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess) g_upload_rc = 3;
        if (cudaMemcpyToSymbol(c_pprog, &g_programs->pfwd, sizeof(PProgram), 0) != cudaSuccess) g_upload_rc = 2;
        if (cudaMemcpyToSymbol(c_pprog, &g_programs->pbwd, sizeof(PProgram), sizeof(PProgram)) != cudaSuccess) g_upload_rc = 2;
        if (cudaFuncSetAttribute(mlp_tc_pair_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess) g_upload_rc = 3;
        if (cudaFuncSetAttribute(mlp_tc_pair_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess) g_upload_rc = 3;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) g_upload_rc = 4;
        g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
    });
    return g_upload_rc;
}

unsigned long long* g_prof_buffer = nullptr;      // diagnostic hook, see ddnerf_mlp_tc_set_profile_buffer

int sm_count() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

}  // namespace
}  // namespace ddnerf

using namespace ddnerf;

#define TC_ENSURE(who)                                                                                      \
    do {                                                                                                    \
        int rc__ = ensure_programs();                                                                       \
        DDNERF_CHECK_ARG(rc__ != 1, "%s: invalid kernel program: %s", who, g_programs->why);                \
        DDNERF_CHECK_ARG(rc__ == 0, "%s: device setup failed (%d): %s", who, rc__, cudaGetErrorString(cudaGetLastError())); \
    } while (0)

/* Diagnostic hook: when set to a device buffer of >= 8 * n_SMs uint64, the chain kernels write per-CTA cycle
 * counters [issuer total, issuer waiting on epilogues, issuer waiting on ring stages, epilogues waiting on
 * MMAs, epilogues busy, epilogue count, -, -]; NULL (default) switches the instrumentation off. */
extern "C" DDNERF_EXPORT int ddnerf_mlp_tc_set_profile_buffer(void* dev_u64) {
    g_prof_buffer = static_cast<unsigned long long*>(dev_u64);
    return 0;
}
/* Kernel variant of the forward / dX chains: 1 = CTA pairs (cluster of 2, tcgen05 cta_group::2; the default), 0 = one CTA
 * per work item, -1 = re-read DDNERF_TC_PAIR.  Returns the previous setting.  Results are bit-identical. */
extern "C" DDNERF_EXPORT int ddnerf_mlp_tc_set_pair_mode(int mode) {
    const int prev = g_pair_mode;
    g_pair_mode = mode < 0 ? -1 : (mode ? 1 : 0);
    return prev;
}
/* 0 when the static kernel programs (ring schedule, op tables) are consistent; host-only check */
extern "C" DDNERF_EXPORT int ddnerf_mlp_tc_program_check(void) {
    Programs* S = build_programs();
    DDNERF_CHECK_ARG(S->ok, "mlp_tc: invalid kernel program: %s", S->why);
    return 0;
}
extern "C" DDNERF_EXPORT int64_t ddnerf_mlp_tc_wimg_bytes(void) {
    Programs* S = build_programs();
    return (int64_t)S->fwd.woff + S->bwd.woff;
}
extern "C" DDNERF_EXPORT int64_t ddnerf_mlp_tc_bias_floats(void) { return tcmlp::kBiasRows * 256; }
extern "C" DDNERF_EXPORT int64_t ddnerf_mlp_tc_items(int64_t rows) { return (rows + tcmlp::kItemRows - 1) / tcmlp::kItemRows; }
extern "C" DDNERF_EXPORT int64_t ddnerf_mlp_tc_enc_bytes(int64_t rows) { return ddnerf_mlp_tc_items(rows) * tcmlp::kEncItemBytes; }
extern "C" DDNERF_EXPORT int64_t ddnerf_mlp_tc_act_save_bytes(int64_t rows) {
    return tcmlp::kSaveLayers * ddnerf_mlp_tc_items(rows) * 2 * (int64_t)tcmlp::kActBytes;
}
extern "C" DDNERF_EXPORT int64_t ddnerf_mlp_tc_mask_save_bytes(int64_t rows) {
    return tcmlp::kSaveLayers * ddnerf_mlp_tc_items(rows) * 2 * 128 * 32;
}

extern "C" DDNERF_EXPORT int ddnerf_mlp_tc_pack(const ddnerf_mlp_params* p, int out_channels, void* wimg, float* bias_pack,
                                                void* stream) {
    DDNERF_CHECK_ARG(p && wimg && bias_pack, "mlp_tc_pack: null pointer");
    DDNERF_CHECK_ARG(out_channels == 4 || out_channels == 6, "mlp_tc_pack: out_channels=%d (4 or 6)", out_channels);
    for (int i = 0; i < (out_channels == 6 ? 13 : 12); ++i) DDNERF_CHECK_ARG(p->w[i] && p->b[i], "mlp_tc_pack: parameter %d is null", i);
    TC_ENSURE("mlp_tc_pack");
    PackArgs a{*p, static_cast<uint8_t*>(wimg), bias_pack, out_channels};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    pack_weights_kernel<<<g_programs->pack.n + kBiasRows, 256, 0, st>>>(a, g_programs->pack.n);
    DDNERF_LAUNCHED("mlp_tc_pack", 1);
    return 0;
}

extern "C" DDNERF_EXPORT int ddnerf_mlp_tc_encode(const float* rays, const float* t_vals, int64_t N, int S, int ray_shape,
                                                  void* enc_img, void* stream) {
    DDNERF_CHECK_ARG(rays && t_vals && enc_img, "mlp_tc_encode: null pointer");
    DDNERF_CHECK_ARG(ray_shape == 0 || ray_shape == 1, "mlp_tc_encode: ray_shape=%d (0 cone, 1 cylinder)", ray_shape);
    if (N * S == 0) return 0;
    const int64_t rows_padded = ddnerf_mlp_tc_items(N * S) * tcmlp::kItemRows;
    encode_img_kernel<<<ceil_div(rows_padded, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        rays, t_vals, static_cast<uint8_t*>(enc_img), N, S, ray_shape, rows_padded);
    DDNERF_LAUNCHED("mlp_tc_encode", 1);
    return 0;
}

static bool pair_mode() {
    if (g_pair_mode < 0) {
        const char* e = getenv("DDNERF_TC_PAIR");
        g_pair_mode = (e && atoi(e) == 0) ? 0 : 1;
    }
    return g_pair_mode == 1;
}

// A {512-byte row, rows_per_half, n_halves} view of the bytes at `base`, no swizzle (the images are pre-swizzled): a box is
// `box_halves` halves traversed with element stride `estride` along the last dimension, landing contiguously in shared memory.
static bool make_map(CUtensorMap* m, const void* base, uint32_t rows_per_half, uint64_t n_halves, uint32_t box_halves, uint32_t estride) {
    const cuuint64_t dims[3] = {256, rows_per_half, n_halves}, strides[2] = {512, (cuuint64_t)rows_per_half * 512};
    const cuuint32_t box[3] = {256, rows_per_half, box_halves}, estr[3] = {1, 1, estride};
    return g_encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// wimg_base: start of the whole packed image (both programs); enc / enc_blocks: the encoded images the producer reads (or null)
template <int PI>
static int launch_pair(const char* who, const ChainArgs& g_in, const void* wimg_base, const void* enc, uint64_t enc_blocks, int max_ctas,
                       void* stream) {
    ChainArgs g = g_in;
    g.pslots = kSlots;
    if (const char* e = getenv("DDNERF_TC_PSLOTS")) g.pslots = std::max(2, std::min(kSlots, atoi(e)));
    PairMaps maps;
    const uint8_t* wb = static_cast<const uint8_t*>(wimg_base);
    const uint32_t* base = g_programs->pbase[PI];
    const uint64_t total = (uint64_t)g_programs->fwd.woff + g_programs->bwd.woff;
    bool ok = make_map(&maps.m[PM_W128], wb + base[0], 16, (total - base[0]) / 8192, 3, 2) &&
              make_map(&maps.m[PM_W72], wb + base[1], 9, (total - base[1]) / 4608, 3, 2) &&
              make_map(&maps.m[PM_W8], wb + base[2], 1, (total - base[2]) / 512, 7, 2) &&
              make_map(&maps.m[PM_W128S], wb + base[0], 16, (total - base[0]) / 8192, 1, 1) &&
              make_map(&maps.m[PM_W72S], wb + base[1], 9, (total - base[1]) / 4608, 1, 1);
    ok = ok && (enc ? make_map(&maps.m[PM_ENC], enc, 16, enc_blocks, 1, 1) : make_map(&maps.m[PM_ENC], wb, 16, total / 8192, 1, 1));
    DDNERF_CHECK_ARG(ok, "%s: cuTensorMapEncodeTiled failed", who);
    const int n_units = (g.n_items + 1) / 2;
    int pairs = std::min(n_units, sm_count() / 2);
    if (max_ctas > 0) pairs = std::max(1, std::min(pairs, max_ctas / 2));
    mlp_tc_pair_kernel<PI><<<2 * pairs, PI == 0 ? kThreads : kThreads - kEncThreads, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(g, maps);
    DDNERF_LAUNCHED(who, 1);
    return 0;
}

static int launch_forward(const char* who, ChainArgs g, int64_t rows, void* stream) {
    const int64_t n_items = ddnerf_mlp_tc_items(rows);
    DDNERF_CHECK_ARG(n_items < (1 << 30), "%s: too many rows", who);
    g.rows = rows;
    g.n_items = (int)n_items;
    if (const char* e = getenv("DDNERF_TC_SAVE_ALIAS")) g.save_alias = atoi(e);
    g.enc_yield = 1;
    if (const char* e = getenv("DDNERF_TC_ENC_YIELD")) g.enc_yield = atoi(e);
    g.prof = g_prof_buffer;
    if (pair_mode()) {
        const uint64_t enc_blocks = (g.enc_mode == 2 ? (uint64_t)sm_count() * 2 : (uint64_t)n_items) * (kEncItemBytes / 8192);
        return launch_pair<0>(who, g, g.wimg, g.enc, enc_blocks, 0, stream);
    }
    const int grid = (int)std::min<int64_t>(n_items, sm_count());
    mlp_tc_chain_kernel<0><<<grid, kThreads, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(g);
    DDNERF_LAUNCHED(who, 1);
    return 0;
}

extern "C" DDNERF_EXPORT int ddnerf_mlp_tc_forward(const void* wimg, const float* bias_pack, const void* enc_img, int64_t rows,
                                                   int out_channels, float* out, void* act_save, void* mask_save, void* stream) {
    DDNERF_CHECK_ARG(wimg && bias_pack && enc_img && out, "mlp_tc_forward: null pointer");
    DDNERF_CHECK_ARG(out_channels == 4 || out_channels == 6, "mlp_tc_forward: out_channels=%d (4 or 6)", out_channels);
    DDNERF_CHECK_ARG((act_save == nullptr) == (mask_save == nullptr), "mlp_tc_forward: act_save and mask_save go together");
    DDNERF_CHECK_ARG(ddnerf_device_is_sm100(), "mlp_tc_forward: the bf16 MLP needs an sm_100 device (tcgen05)");
    if (rows == 0) return 0;
    TC_ENSURE("mlp_tc_forward");
    ChainArgs g{};
    g.wimg = static_cast<const uint8_t*>(wimg);
    g.bias = bias_pack;
    g.enc = static_cast<const uint8_t*>(enc_img);
    g.enc_mode = 0;
    g.out = out;
    g.save = static_cast<uint8_t*>(act_save);
    g.mask = static_cast<uint32_t*>(mask_save);
    g.C = out_channels;
    return launch_forward("mlp_tc_forward", g, rows, stream);
}

extern "C" DDNERF_EXPORT int64_t ddnerf_mlp_tc_enc_scratch_bytes(void) { return (int64_t)sm_count() * 2 * tcmlp::kEncItemBytes; }

extern "C" DDNERF_EXPORT int ddnerf_mlp_tc_forward_rays(const void* wimg, const float* bias_pack, const float* rays,
                                                        const float* t_vals, int64_t N, int S, int ray_shape, int out_channels,
                                                        float* out, void* enc_img, void* enc_scratch, void* act_save,
                                                        void* mask_save, void* stream) {
    DDNERF_CHECK_ARG(wimg && bias_pack && rays && t_vals && out, "mlp_tc_forward_rays: null pointer");
    DDNERF_CHECK_ARG((enc_img != nullptr) != (enc_scratch != nullptr),
                     "mlp_tc_forward_rays: give either enc_img (kept for the backward) or enc_scratch (inference)");
    DDNERF_CHECK_ARG(out_channels == 4 || out_channels == 6, "mlp_tc_forward_rays: out_channels=%d (4 or 6)", out_channels);
    DDNERF_CHECK_ARG((act_save == nullptr) == (mask_save == nullptr), "mlp_tc_forward_rays: act_save and mask_save go together");
    DDNERF_CHECK_ARG(!act_save || enc_img, "mlp_tc_forward_rays: the training forward needs enc_img (the weight-gradient kernel reads it)");
    DDNERF_CHECK_ARG(ray_shape == 0 || ray_shape == 1, "mlp_tc_forward_rays: ray_shape=%d (0 cone, 1 cylinder)", ray_shape);
    DDNERF_CHECK_ARG(S >= 1 && N >= 0, "mlp_tc_forward_rays: N=%lld S=%d", (long long)N, S);
    DDNERF_CHECK_ARG(ddnerf_device_is_sm100(), "mlp_tc_forward_rays: the bf16 MLP needs an sm_100 device (tcgen05)");
    if (N == 0) return 0;
    TC_ENSURE("mlp_tc_forward_rays");
    ChainArgs g{};
    g.wimg = static_cast<const uint8_t*>(wimg);
    g.bias = bias_pack;
    g.enc = static_cast<const uint8_t*>(enc_img ? enc_img : enc_scratch);
    g.enc_mode = enc_img ? 1 : 2;
    g.rays = rays;
    g.t_vals = t_vals;
    g.N = N;
    g.S = S;
    g.ray_shape = ray_shape;
    g.out = out;
    g.save = static_cast<uint8_t*>(act_save);
    g.mask = static_cast<uint32_t*>(mask_save);
    g.C = out_channels;
    return launch_forward("mlp_tc_forward_rays", g, N * (int64_t)S, stream);
}

extern "C" DDNERF_EXPORT int ddnerf_mlp_tc_backward_dx(const void* wimg, const float* bias_pack, const float* grad_out,
                                                       int64_t rows, int out_channels, const void* mask_save, void* dz_save,
                                                       int max_ctas, void* stream) {
    DDNERF_CHECK_ARG(wimg && bias_pack && grad_out && mask_save && dz_save, "mlp_tc_backward_dx: null pointer");
    DDNERF_CHECK_ARG(out_channels == 4 || out_channels == 6, "mlp_tc_backward_dx: out_channels=%d (4 or 6)", out_channels);
    DDNERF_CHECK_ARG(ddnerf_device_is_sm100(), "mlp_tc_backward_dx: the bf16 MLP needs an sm_100 device (tcgen05)");
    if (rows == 0) return 0;
    TC_ENSURE("mlp_tc_backward_dx");
    const int64_t n_items = ddnerf_mlp_tc_items(rows);
    DDNERF_CHECK_ARG(n_items < (1 << 30), "mlp_tc_backward_dx: too many rows");
    ChainArgs g{};
    g.wimg = static_cast<const uint8_t*>(wimg) + g_programs->fwd.woff;
    g.bias = bias_pack;
    g.gout = grad_out;
    g.save = static_cast<uint8_t*>(dz_save);
    g.mask = static_cast<uint32_t*>(const_cast<void*>(mask_save));
    g.rows = rows;
    g.n_items = (int)n_items;
    g.C = out_channels;
    if (const char* e = getenv("DDNERF_TC_SAVE_ALIAS")) g.save_alias = atoi(e);
    g.prof = g_prof_buffer;
    if (pair_mode()) return launch_pair<1>("mlp_tc_backward_dx", g, wimg, nullptr, 0, max_ctas, stream);
    int grid = (int)std::min<int64_t>(n_items, sm_count());
    if (max_ctas > 0) grid = std::min(grid, max_ctas);
    mlp_tc_chain_kernel<1><<<grid, kThreads - kEncThreads, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(g);
    DDNERF_LAUNCHED("mlp_tc_backward_dx", 1);
    return 0;
}
