// K1c (bf16 throughput mode): weight / bias gradients of the NeRF MLP (models/base_architectures.py)
// from the tile images the chain kernels saved:  dW_l = dZ_l^T . A_{l-1},  db_l = colsum(dZ_l).
//
// The reduction runs over the SAMPLE dimension (K = rows), so both operands are the saved
// [128 rows x features] K-major tile images read as MN-major operands (no transpose pass):
//   A operand = dZ_l  (M = 128 output features per half, two halves per layer),
//   B operand = A_{l-1} / encoded xyz / encoded view directions (N = input features).
// A CTA owns ONE (layer-op, tile range) work item: it streams 64-row half tiles of both images
// through a 3-stage shared-memory ring (bulk async copies), accumulates the whole [256 x N] fp32
// gradient in TMEM (2 x 256 columns) across its tile range, and flushes once with fp32 atomics.
// Four SIMT warps read the same stages for the small reductions that do not fit the MMA shape:
// bias gradients (column sums of dZ), fc_alpha (g_density-weighted column sum of feat) and the
// colour / (mu, sigma) heads (5 weighted column sums of the view-branch activations).
// HBM-bound by design: 128 KB of tile images per 16.8 MFLOP.
#include <algorithm>
#include <cstdlib>
#include <mutex>

#define DDNERF_TC_WATCHDOG 1

#include "mlp_tc.cuh"

namespace ddnerf {
namespace {

using namespace tcmlp;

constexpr int kDwStages = 3;
constexpr int kDwStageBytes = 65536;           // [A half tile 32 KB | B half tile 32 KB]
constexpr int kDwThreads = 192;                // warp 0 producer, 1 MMA issuer, 2..5 SIMT + flush
constexpr int kMaxWork = 480;                  // work items (a CTA takes up to three)
constexpr int kMaxCtas = 160;

enum : int8_t { B_ACT = 0, B_XYZ = 1, B_DIRX = 2 };
enum : int8_t { FLUSH_STD = 0, FLUSH_DIR = 1, FLUSH_HEADS = 2 };

struct DwOp {
    int8_t a_layer, a_block0, a_blocks, m_halves;      // dZ image: first block / blocks loaded per half tile
    int8_t b_kind, b_layer, b_blocks, param;           // B image; parameter index of the standard flush
    int8_t colsum, flush, pad[2];
    uint16_t ld, col0, out_rows, n_cols;
    uint32_t tx_bytes;                                 // bytes landing in one stage
};
constexpr int kNumOps = 13;
// parameter indices: 0..7 layers_xyz, 8 fc_feat, 9 fc_alpha, 10 layers_dir.0, 11 fc_rgb, 12 fc_mu_sigma.
// The dZ image of layer 9 holds dZ_dir in columns 0..127 and the raw output cotangents
// [g_density, g_r, g_g, g_b, g_mu, g_sigma] in columns 128..133 (written by the dX chain), so fc_alpha and
// the colour / (mu, sigma) heads are M-rows of ordinary MMAs: rows 128.. of the dir op against feat,
// and of the heads op against the view-branch activations.
__constant__ DwOp c_ops[kNumOps] = {
    //a_l b0 nb mh  b_kind    b_l b_b  p  cs flush                ld  col0 rows cols  tx
    {0, 0, 4, 2, B_XYZ, -1, 0, 0, 1, FLUSH_STD, {0, 0}, 96, 0, 256, 96, 32768 + 12288},
    {1, 0, 4, 2, B_ACT, 0, 4, 1, 1, FLUSH_STD, {0, 0}, 256, 0, 256, 256, 65536},
    {2, 0, 4, 2, B_ACT, 1, 4, 2, 1, FLUSH_STD, {0, 0}, 256, 0, 256, 256, 65536},
    {3, 0, 4, 2, B_ACT, 2, 4, 3, 1, FLUSH_STD, {0, 0}, 256, 0, 256, 256, 65536},
    {4, 0, 4, 2, B_ACT, 3, 4, 4, 1, FLUSH_STD, {0, 0}, 256, 0, 256, 256, 65536},
    {5, 0, 4, 2, B_ACT, 4, 4, 5, 1, FLUSH_STD, {0, 0}, 352, 96, 256, 256, 65536},
    {5, 0, 4, 2, B_XYZ, -1, 0, 5, 0, FLUSH_STD, {0, 0}, 352, 0, 256, 96, 32768 + 12288},
    {6, 0, 4, 2, B_ACT, 5, 4, 6, 1, FLUSH_STD, {0, 0}, 256, 0, 256, 256, 65536},
    {7, 0, 4, 2, B_ACT, 6, 4, 7, 1, FLUSH_STD, {0, 0}, 256, 0, 256, 256, 65536},
    {8, 0, 4, 2, B_ACT, 7, 4, 8, 1, FLUSH_STD, {0, 0}, 256, 0, 256, 256, 65536},
    {9, 0, 3, 2, B_ACT, 8, 4, 10, 1, FLUSH_DIR, {0, 0}, 283, 0, 128, 256, 24576 + 32768},
    {9, 0, 2, 1, B_DIRX, -1, 0, 10, 0, FLUSH_STD, {0, 0}, 283, 256, 128, 27, 16384 + 4096},
    {9, 2, 1, 1, B_ACT, 9, 2, 11, 0, FLUSH_HEADS, {0, 0}, 128, 0, 6, 128, 8192 + 16384},
};
// Cost model of the work split, measured on a B200 with the kernel's own per-item cycle counters
// (ddnerf_mlp_tc_dw_set_profile_buffer; all 148 CTAs streaming, 4096 tiles): cycles per tile of each op -- the bytes moved
// (88 / 128 / 112 / 40 / 48 KB) plus about 800 cycles per tile that do not scale with them -- and cycles of one flush
// (fp32 reductions of the op's [M x N] accumulator, four columns per instruction where the rows are 16-byte aligned:
// 4 tiles' worth for a 256 x 256 layer; 18 with scalar atomics).
const uint32_t h_op_weight[kNumOps] = {3600, 4120, 4120, 4120, 4120, 4120, 3320, 4120, 4120, 4120, 3930, 2370, 2630};
const uint32_t h_op_flush[kNumOps] = {6700, 17500, 17500, 17500, 17500, 17500, 6700, 17500, 17500, 17500, 29000, 3400, 1100};

struct WorkItem { uint32_t op, t0, t1; };

struct DwArgs {
    const uint8_t* act;       // act_save  [10][n_tiles][64 KB]
    const uint8_t* dz;        // dz_save   [10][n_tiles][64 KB]
    const uint8_t* enc;       // encoded images [n_items][64 KB]
    const float* gout;        // [rows, C]
    ddnerf_mlp_grads grads;   // fp32 accumulators (caller zeroes them)
    int64_t rows;
    int n_tiles, C;
    int n_work;
    unsigned long long* prof;   // diagnostic: per work item {op, tiles, cycles to the last MMA, cycles of the flush}
    uint32_t dz_alias;          // experiment knob (DDNERF_TC_DW_DZ_ALIAS): dZ images are read from tile % dz_alias (L2-resident)
};
struct DwWork { WorkItem w[kMaxWork]; uint16_t first[kMaxCtas + 1]; };   // CTA i: items first[i] .. first[i+1]-1

struct __align__(16) DwCtl {
    uint64_t full[kDwStages], empty[kDwStages], acc_done, acc_free;
    uint32_t tmem_base, pad;
};
constexpr int kDwSmemBytes = kDwStages * kDwStageBytes + (int)sizeof(DwCtl);

__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return (uint32_t)v;
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

__device__ void dw_producer(const DwArgs& g, const DwOp& op, const WorkItem wk, DwCtl* ctl, uint8_t* ring, uint32_t& s, uint32_t& phase) {
    const size_t layer_pitch = (size_t)g.n_tiles * kActBytes;
    for (uint32_t t = wk.t0; t < wk.t1; ++t) {
        const uint8_t* a_src = op.a_layer >= 0 ? g.dz + op.a_layer * layer_pitch + (size_t)(g.dz_alias ? t % g.dz_alias : t) * kActBytes : nullptr;
        const uint8_t* b_src = op.b_layer >= 0 ? g.act + op.b_layer * layer_pitch + (size_t)t * kActBytes : nullptr;
        const uint8_t* e_src = g.enc + (size_t)(t >> 1) * kEncItemBytes;
        const uint32_t T = t & 1u;
        for (uint32_t hf = 0; hf < 2; ++hf) {
            tc::mbar_wait(&ctl->empty[s], phase ^ 1);
            uint8_t* st = ring + s * kDwStageBytes;
            uint64_t* bar = &ctl->full[s];
            tc::mbar_expect_tx(bar, op.tx_bytes);
            for (int b = 0; b < op.a_blocks; ++b)
                tc::bulk_g2s(st + b * 8192, a_src + (op.a_block0 + b) * 16384 + hf * 8192, 8192, bar);
            for (int b = 0; b < op.b_blocks; ++b) tc::bulk_g2s(st + 32768 + b * 8192, b_src + b * 16384 + hf * 8192, 8192, bar);
            if (op.b_kind == B_XYZ) {
                tc::bulk_g2s(st + 32768, e_src + T * 16384 + hf * 4096, 4096, bar);
                tc::bulk_g2s(st + 32768 + 4096, e_src + T * 16384 + 8192 + hf * 4096, 4096, bar);
                tc::bulk_g2s(st + 32768 + 8192, e_src + 32768 + T * 8192 + hf * 4096, 4096, bar);
            } else if (op.b_kind == B_DIRX) {
                tc::bulk_g2s(st + 32768, e_src + 49152 + T * 8192 + hf * 4096, 4096, bar);
            }
            if (++s == kDwStages) { s = 0; phase ^= 1; }
        }
    }
}

__device__ void dw_mma(const DwOp& op, const WorkItem wk, DwCtl* ctl, uint32_t ring_u32, uint32_t& s, uint32_t& phase, int item) {
    const uint32_t tmem = ctl->tmem_base;
    if (item > 0) {                         // the previous item's accumulator has been read out of TMEM
        tc::mbar_wait(&ctl->acc_free, (uint32_t)(item - 1) & 1u);
        tc::tc_fence_after_sync();
    }
    // MN-major operand descriptors (tile image rows = K): SW128 blocks [rows x 64 features], LBO = block
    // pitch, SBO = 8 rows; SW64 blocks [rows x 32 features]
    constexpr uint64_t kD128 = tc::smem_desc(0, 8192, 1024, tc::LAYOUT_SW128);
    constexpr uint64_t kD64 = tc::smem_desc(0, 4096, 512, tc::LAYOUT_SW64);
    const uint32_t id_main = tc::idesc_bf16(128, op.b_blocks > 0 ? op.b_blocks * 64 : 256, 1, 1);
    const uint32_t id64 = tc::idesc_bf16(128, 64, 1, 1), id32 = tc::idesc_bf16(128, 32, 1, 1);
    const uint32_t n_stages = (wk.t1 - wk.t0) * 2;
    for (uint32_t it = 0; it < n_stages; ++it) {
        tc::mbar_wait(&ctl->full[s], phase);
        tc::tc_fence_after_sync();
        if (tc::elect_one()) {
            const uint32_t st = ring_u32 + s * kDwStageBytes;
            const uint32_t acc = it > 0 ? 1u : 0u;
            for (int hm = 0; hm < op.m_halves; ++hm) {
                const uint32_t d = tmem + (uint32_t)hm * 256u;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t a = kD128 | (uint64_t)(((st + hm * 16384 + k * 2048) >> 4) & 0x3FFF);
                    const uint32_t ak = (acc | (uint32_t)k) ? 1u : 0u;
                    if (op.b_kind == B_XYZ) {
                        tc::mma_f16_ss(d, a, kD64 | (uint64_t)(((st + 32768 + k * 1024) >> 4) & 0x3FFF), id64, ak);
                        tc::mma_f16_ss(d + 64, a, kD64 | (uint64_t)(((st + 32768 + 8192 + k * 1024) >> 4) & 0x3FFF), id32, ak);
                    } else if (op.b_kind == B_DIRX) {              // cat(feat, dirs): the 27 (+5) direction columns
                        tc::mma_f16_ss(d, a, kD64 | (uint64_t)(((st + 32768 + k * 1024) >> 4) & 0x3FFF), id32, ak);
                    } else {
                        tc::mma_f16_ss(d, a, kD128 | (uint64_t)(((st + 32768 + k * 2048) >> 4) & 0x3FFF), id_main, ak);
                    }
                }
            }
            tc::mma_commit(&ctl->empty[s]);
            if (it + 1 == n_stages) tc::mma_commit(&ctl->acc_done);
        }
        __syncwarp();
        if (++s == kDwStages) { s = 0; phase ^= 1; }
    }
}

// byte offset of the bf16 pair (row r, columns c, c+1; c even) inside a [rows x 64] SWIZZLE_128B block
__device__ __forceinline__ uint32_t sw128_pair(uint32_t r, uint32_t c) {
    return r * 128u + ((((c >> 3) ^ r) & 7u) << 4) + ((c & 7u) << 1);
}

__device__ void dw_simt(const DwArgs& g, const DwOp& op, const WorkItem wk, DwCtl* ctl, uint32_t ring_u32, int tid, uint32_t& s,
                        uint32_t& phase, int item, int witem) {
    // bias gradients: thread tid sums columns 2 tid, 2 tid + 1 of the dZ half tiles
    float cs0 = 0.f, cs1 = 0.f;
    const uint32_t c2 = 2u * (uint32_t)tid;
    const uint32_t blk2 = (c2 >> 6) * 8192u, cc2 = c2 & 63u;
    const bool cs_on = op.colsum && (int)c2 < op.a_blocks * 64;
    const uint32_t n_stages = (wk.t1 - wk.t0) * 2;
    const long long t_start = g.prof ? clock64() : 0;
    for (uint32_t it = 0; it < n_stages; ++it) {
        tc::mbar_wait(&ctl->full[s], phase);
        if (cs_on) {
            const uint32_t st = ring_u32 + s * kDwStageBytes + blk2;
            float p0 = 0.f, p1 = 0.f, q0 = 0.f, q1 = 0.f;          // two independent chains per column
#pragma unroll 8
            for (uint32_t r = 0; r < 64; r += 2) {
                const uint32_t w0 = lds32(st + sw128_pair(r, cc2)), w1 = lds32(st + sw128_pair(r + 1, cc2));
                p0 += bf_lo(w0); p1 += bf_hi(w0);
                q0 += bf_lo(w1); q1 += bf_hi(w1);
            }
            cs0 += p0 + q0;
            cs1 += p1 + q1;
        }
        tc::mbar_arrive(&ctl->empty[s]);
        if (++s == kDwStages) { s = 0; phase ^= 1; }
    }
    // ---- flush ---------------------------------------------------------------------------------
    if (cs_on) {
        if (op.flush == FLUSH_DIR) {       // columns: 0..127 layers_dir.0, 128 fc_alpha, 129..131 fc_rgb, 132..133 fc_mu_sigma
            for (int j = 0; j < 2; ++j) {
                const int c = (int)c2 + j;
                const float v = j ? cs1 : cs0;
                if (c < 128) atomicAdd(g.grads.b[10] + c, v);
                else if (c == 128) atomicAdd(g.grads.b[9], v);
                else if (c < 132) atomicAdd(g.grads.b[11] + (c - 129), v);
                else if (c < 134 && g.C == 6) atomicAdd(g.grads.b[12] + (c - 132), v);
            }
        } else {
            float* db = g.grads.b[op.param];
            atomicAdd(db + c2, cs0);
            atomicAdd(db + c2 + 1, cs1);
        }
    }
    tc::mbar_wait(&ctl->acc_done, (uint32_t)item & 1u);
    tc::tc_fence_after_sync();
    const long long t_acc = g.prof ? clock64() : 0;
    // a warp may read the TMEM lane quarter (CTA warp index % 4): warps 2..5 -> quarters 2,3,0,1
    const int quarter = ((tid >> 5) + 2) & 3;
    const int lane = tid & 31;
    for (int hm = 0; hm < op.m_halves; ++hm) {
        const int o = hm * 128 + quarter * 32 + lane;             // row of the [M, N] accumulator
        float* dst = nullptr;                                      // start of this row's destination
        const int n_cols = op.n_cols;
        if (op.flush == FLUSH_STD) {
            if (o < op.out_rows) dst = g.grads.w[op.param] + (size_t)o * op.ld + op.col0;
        } else if (op.flush == FLUSH_DIR) {
            if (hm == 0) dst = g.grads.w[10] + (size_t)o * 283;
            else if (o == 128) dst = g.grads.w[9];                 // row 128 = g_density^T . feat
        } else {                                                   // rows 1..3 fc_rgb, 4..5 fc_mu_sigma
            if (o >= 1 && o <= 3) dst = g.grads.w[11] + (o - 1) * 128;
            else if (o >= 4 && o <= 5 && g.C == 6) dst = g.grads.w[12] + (o - 4) * 128;
        }
        const uint32_t taddr = ctl->tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)hm * 256u;
        for (int c0 = 0; c0 < n_cols; c0 += 32) {
            uint32_t v[32];
            tc::tmem_ld32(taddr + c0, v);
            tc::tmem_ld_wait();
            if (dst) {
                // four columns per reduction where the row is 16-byte aligned (every 256- and 352-wide weight); the
                // flush of a 256 x 256 accumulator was 65,536 scalar atomics = 18 tiles' worth of streaming time
                if (((reinterpret_cast<uintptr_t>(dst + c0) & 15u) == 0) && c0 + 32 <= n_cols) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + i), "f"(__uint_as_float(v[i])),
                                     "f"(__uint_as_float(v[i + 1])), "f"(__uint_as_float(v[i + 2])), "f"(__uint_as_float(v[i + 3]))
                                     : "memory");
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c0 + i < n_cols) atomicAdd(dst + c0 + i, __uint_as_float(v[i]));
                }
            }
        }
    }
    tc::tc_fence_before_sync();
    tc::mbar_arrive(&ctl->acc_free);         // TMEM may be overwritten by the CTA's next work item
    if (g.prof && tid == 0) {
        unsigned long long* o = g.prof + (size_t)witem * 4;
        o[0] = wk.op; o[1] = wk.t1 - wk.t0; o[2] = (unsigned long long)(t_acc - t_start); o[3] = (unsigned long long)(clock64() - t_acc);
    }
}

__global__ void __launch_bounds__(kDwThreads, 1) mlp_tc_dw_kernel(const DwArgs g, const DwWork work) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* ring = smem;
    DwCtl* ctl = reinterpret_cast<DwCtl*>(smem + kDwStages * kDwStageBytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w0 = work.first[blockIdx.x], w1 = work.first[blockIdx.x + 1];
    if (threadIdx.x == 0) {
        if ((tc::smem_u32(smem) & 1023u) != 0) { printf("ddnerf mlp_tc_dw: shared memory is not 1024-byte aligned\n"); __trap(); }
        for (int s = 0; s < kDwStages; ++s) { tc::mbar_init(&ctl->full[s], 1); tc::mbar_init(&ctl->empty[s], 1 + 128); }
        tc::mbar_init(&ctl->acc_done, 1);
        tc::mbar_init(&ctl->acc_free, 128);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(&ctl->tmem_base, 512);
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();

    // every role walks the CTA's work items in the same order; the ring position carries over from item to item
    uint32_t s = 0, phase = 0;
    for (int w = w0; w < w1; ++w) {
        const WorkItem wk = work.w[w];
        const DwOp op = c_ops[wk.op];
        if (warp == 0) { if (lane == 0) dw_producer(g, op, wk, ctl, ring, s, phase); }
        else if (warp == 1) dw_mma(op, wk, ctl, tc::smem_u32(ring), s, phase, w - w0);
        else dw_simt(g, op, wk, ctl, tc::smem_u32(ring), threadIdx.x - 64, s, phase, w - w0, w);
    }

    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(ctl->tmem_base, 512);
}

// Work split.  The (op, tile) pairs are laid on one line, op after op, and cut into at most `sms` pieces of equal
// COST: a piece pays every tile's cycles plus one flush per op it touches, so a piece that crosses an op boundary (two
// work items, run back to back by one CTA) gets fewer tiles.  The common piece cost is found by bisection.  (Round 1a
// gave every op a whole number of CTAs: with 13 ops on 148 SMs the rounding alone left a 13:12 imbalance, and ncu
// showed the SMs active for 71 % of the kernel's duration.)  Every (op, tile) pair is covered exactly once.
// Returns the number of CTAs.
int assign_pieces(uint32_t n_tiles, int sms, uint64_t budget, DwWork* work) {
    int n_items = 0, n_ctas = 0, op = 0;
    uint32_t tile = 0;
    if (work) work->first[0] = 0;
    while (op < kNumOps) {
        if (n_ctas >= sms && !work) return sms + 1;                 // does not fit
        uint64_t left = budget;
        bool opened = false;
        while (op < kNumOps) {
            const uint64_t w = h_op_weight[op], f = h_op_flush[op];
            // open an item of this op only if a few tiles fit behind its flush (or the piece is still empty)
            if (opened && left < f + 4 * w) break;
            const uint64_t room = left > f ? (left - f) / w : 0;
            const uint32_t take = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(room, 1), n_tiles - tile);
            if (work && n_items < kMaxWork) work->w[n_items] = WorkItem{(uint32_t)op, tile, tile + take};
            ++n_items;
            opened = true;
            const uint64_t cost = f + w * take;
            left = left > cost ? left - cost : 0;
            tile += take;
            if (tile == n_tiles) { ++op; tile = 0; } else break;    // the piece ends inside this op
            if (left == 0) break;
        }
        ++n_ctas;
        if (work) {
            if (n_ctas > kMaxCtas) return -1;
            work->first[n_ctas] = (uint16_t)std::min(n_items, kMaxWork);
        }
    }
    return n_ctas;
}

int plan_work(uint32_t n_tiles, int sms, DwWork& work) {
    sms = std::max(1, std::min(sms, kMaxCtas));
    uint64_t lo = 1, hi = 0;
    for (int o = 0; o < kNumOps; ++o) hi += (uint64_t)h_op_weight[o] * n_tiles + h_op_flush[o];
    while (lo < hi) {                                                  // smallest piece cost that needs <= sms pieces
        const uint64_t mid = (lo + hi) / 2;
        if (assign_pieces(n_tiles, sms, mid, nullptr) <= sms) hi = mid; else lo = mid + 1;
    }
    return assign_pieces(n_tiles, sms, lo, &work);
}

unsigned long long* g_dw_prof = nullptr;
std::once_flag g_dw_once;
int g_dw_rc = 0;

}  // namespace
}  // namespace ddnerf

using namespace ddnerf;

extern "C" DDNERF_EXPORT int ddnerf_mlp_tc_backward_dw(const void* act_save, const void* dz_save, const void* enc_img,
                                                       const float* grad_out, const ddnerf_mlp_grads* grads, int64_t rows,
                                                       int out_channels, int max_ctas, void* stream) {
    DDNERF_CHECK_ARG(act_save && dz_save && enc_img && grad_out && grads, "mlp_tc_backward_dw: null pointer");
    DDNERF_CHECK_ARG(out_channels == 4 || out_channels == 6, "mlp_tc_backward_dw: out_channels=%d (4 or 6)", out_channels);
    for (int i = 0; i < (out_channels == 6 ? 13 : 12); ++i)
        DDNERF_CHECK_ARG(grads->w[i] && grads->b[i], "mlp_tc_backward_dw: gradient buffer %d is null", i);
    DDNERF_CHECK_ARG(ddnerf_device_is_sm100(), "mlp_tc_backward_dw: the bf16 MLP needs an sm_100 device (tcgen05)");
    if (rows == 0) return 0;
    std::call_once(g_dw_once, [] {
        if (cudaFuncSetAttribute(mlp_tc_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwSmemBytes) != cudaSuccess) g_dw_rc = 1;
    });
    DDNERF_CHECK_ARG(g_dw_rc == 0, "mlp_tc_backward_dw: device setup failed: %s", cudaGetErrorString(cudaGetLastError()));
    const int64_t n_tiles64 = ddnerf_mlp_tc_items(rows) * 2;
    DDNERF_CHECK_ARG(n_tiles64 < (1ll << 31), "mlp_tc_backward_dw: too many rows");
    const uint32_t n_tiles = (uint32_t)n_tiles64;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);

    if (max_ctas > 0) sms = std::min(sms, max_ctas);
    DwWork work{};
    const int n_work = plan_work(n_tiles, sms, work);        // number of CTAs
    DwArgs g{};
    g.act = static_cast<const uint8_t*>(act_save);
    g.dz = static_cast<const uint8_t*>(dz_save);
    g.enc = static_cast<const uint8_t*>(enc_img);
    g.gout = grad_out;
    g.grads = *grads;
    g.rows = rows;
    g.n_tiles = (int)n_tiles;
    g.C = out_channels;
    g.n_work = n_work;
    g.prof = g_dw_prof;
    if (const char* e = getenv("DDNERF_TC_DW_DZ_ALIAS")) g.dz_alias = (uint32_t)atoi(e);
    mlp_tc_dw_kernel<<<n_work, kDwThreads, kDwSmemBytes, static_cast<cudaStream_t>(stream)>>>(g, work);
    DDNERF_LAUNCHED("mlp_tc_backward_dw", 1);
    return 0;
}

/* Diagnostic hook: a device buffer of >= 4 * 480 uint64 receives, per work item of the next dW launches,
 * {layer-op, tiles, cycles until its last MMA completed, cycles of its flush}; NULL switches it off. */
extern "C" DDNERF_EXPORT int ddnerf_mlp_tc_dw_set_profile_buffer(void* dev_u64) {
    g_dw_prof = static_cast<unsigned long long*>(dev_u64);
    return 0;
}

/* host-side work split of ddnerf_mlp_tc_backward_dw for `rows` sample rows on `sms` SMs: writes
 * (op, first tile, end tile) triples, returns their number (test hook, no device work) */
extern "C" DDNERF_EXPORT int ddnerf_mlp_tc_dw_plan(int64_t rows, int sms, uint32_t* triples, int max_items) {
    DDNERF_CHECK_ARG(triples && rows >= 0 && sms > 0, "mlp_tc_dw_plan: bad arguments");
    DwWork work{};
    const int n_ctas = rows == 0 ? 0 : plan_work((uint32_t)(ddnerf_mlp_tc_items(rows) * 2), sms, work);
    const int n = n_ctas == 0 ? 0 : work.first[n_ctas];
    for (int i = 0; i < n && i < max_items; ++i) {
        triples[3 * i] = work.w[i].op;
        triples[3 * i + 1] = work.w[i].t0;
        triples[3 * i + 2] = work.w[i].t1;
    }
    return n;
}
