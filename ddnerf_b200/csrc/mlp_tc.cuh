// Shared definitions of the tensor-core (bf16, tcgen05) MLP kernels: the static per-work-item
// PROGRAM (ring loads, MMA groups, epilogues) that the producer / MMA-issuer / epilogue warps of the
// chain kernels walk in lock step, and the global-memory image formats.
//
// Work item  = 256 sample rows = two 128-row tiles T0, T1 processed by one CTA.
// Ring       = 6 slots of 16 KB fed by bulk async copies (weight stages, encoded-feature blocks).
//              Every program loads a multiple of 6 stages per item, so the slot of a stage is static.
// Act buffer = per tile 64 KB: [4 k-blocks][128 rows][128 B], K-major SWIZZLE_128B, overwritten in
//              place by every layer's epilogue (the next layer's A operand).
// TMEM       = 512 columns: tile T owns columns [256 T, 256 T + 256) (fp32 accumulator 128 x 256).
//
// Saved tile images (training): act_save / dz_save = [10 layers][n_tiles][64 KB] in the act-buffer
// format; layers 0..7 = layers_xyz.0-7, 8 = fc_feat, 9 = layers_dir.0 (+ density column 128 in dz).
#pragma once
#include "tc.cuh"

namespace ddnerf {
namespace tcmlp {

constexpr int kSlots = 6;
constexpr int kSlotBytes = 16384;
constexpr int kActBytes = 65536;
constexpr int kItemRows = 256;
constexpr int kEncItemBytes = 65536;     // [T0 b0,b1 | T1 b0,b1 | T0 b2 | T1 b2 | T0 dir | T1 dir], blocks [128 x 32] SW64
constexpr int kThreads = 384;            // warps 0-7 epilogue (4 per tile), 8 producer, 9 MMA issuer, 10-11 encoder (forward)
constexpr int kEncThreads = 64;
constexpr int kMaxLoads = 96, kMaxMmas = 208, kMaxEpis = 12, kMaxPack = 192;
constexpr int kSaveLayers = 10;
constexpr int kBiasRows = 16, kHeadWRow = 11;     // packed fp32 table: 11 bias rows + 5 head-weight rows
constexpr uint32_t kRingOff = 2 * kActBytes;          // smem byte offset of the ring

enum : uint8_t { F_WAIT_ACT = 1, F_FIRST = 2, F_COMMIT_ACC = 4, F_WAIT_PREV = 16,
                 F_ENC_DONE = 32 };       // the item's last encoded block has landed in the ring: its image may be overwritten
enum : uint8_t { EPI_ACT = 0, EPI_DIR = 1, EPI_OUT = 2, EPI_BWD_IN = 3, EPI_BWD_MASK = 4 };
enum : uint8_t { LOAD_W = 0, LOAD_ENC = 1 };

struct Load { uint32_t kind, off, bytes; };
// One MMA group, fully resolved on the host: descriptor words relative to the dynamic shared
// memory base, ring slots to wait for / release, accumulator column.
struct __align__(16) Mma {
    uint32_t a_lo, b_lo;              // (byte offset from the smem base) >> 4
    uint32_t a_hi, b_hi;              // descriptor bits 32..63 (SBO, version, layout)
    uint32_t idesc;
    uint16_t d_col;                   // TMEM column of the accumulator
    uint8_t nk16, flags;
    uint8_t tile, n_wait, rel0, rel1; // n_wait: ring stages to wait for first; rel*: slots released (0xFF none)
    uint32_t kind;                    // OP_GENERIC: the fields above; OP_HPART: a whole K = 256 layer part (below)
};
// OP_HPART: the regular bulk of both programs -- one K = 256 accumulation over the act buffer for both
// tiles (8 consecutive ring stages, tiles interleaved half a layer apart), issued by a specialised
// routine instead of 16 table entries.  Uses idesc and flags & F_COMMIT_ACC only.
enum : uint32_t { OP_GENERIC = 0, OP_HPART = 1 };
struct Epi {
    uint8_t mode, relu;
    int8_t save_layer;                // index into the activation / dZ save area, -1: none
    int8_t mask_layer;                // bwd: ReLU mask applied (index into mask_save), -1: none
    uint16_t ncols;
    uint16_t bias_off;                // float offset into the packed bias table
    uint32_t save_bytes;              // bytes of the act buffer stored when saving
    uint32_t signal;                  // arrive on act_ready after the epilogue
};
struct Program {
    int n_loads, n_mmas, n_epis, pad;
    Mma mmas[kMaxMmas];
    Load loads[kMaxLoads];
    Epi epis[kMaxEpis];
};
// one weight stage of the packed image: n_total rows x 32 k, SWIZZLE_64B
struct PackEntry { uint32_t dst_off; uint16_t kind, p, n_total, k0, n0, pad; };
struct PackTable { int n; uint32_t total_bytes; PackEntry e[kMaxPack]; };

enum : uint16_t { PK_FWD = 0, PK_FWD_DIR = 1, PK_FWD_HEADS = 2, PK_BWD = 3, PK_BWD_DIR = 4 };

}  // namespace tcmlp
}  // namespace ddnerf
