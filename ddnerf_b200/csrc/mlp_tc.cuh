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

// ---- CTA-pair variant (cluster of 2, tcgen05 cta_group::2) ----------------------------------------
// Work unit = 512 sample rows = two super-tiles of 256 rows; CTA `rank` of the pair holds rows [128 rank, 128 rank + 128)
// of each super-tile (its act buffer T, its TMEM columns [256 T, 256 T + 256)), i.e. global 128-row tile
// (2 unit + T) * 2 + rank.  One M = 256 MMA covers a super-tile; its B operand (a weight chunk [N x 32 k]) is split by
// rows between the two CTAs, so a 16 KB ring slot holds TWO K = 32 chunks (K = 64) and the 6-slot ring 1.5 layers.
// The two super-tiles run a whole layer apart: every stage is fetched once per super-tile (same L2 -> SM bytes per row as the
// single-CTA kernel) and released right behind its MMAs, and a super-tile's epilogue has the other one's whole layer of MMAs
// to hide in.  With PF_SHARED on the weight-only stages the same code runs them half a layer (two stages) apart instead: T0
// takes the stages as they land, T1 follows and releases them (half the L2 -> SM bytes; measured slower, see build_pair).
// Stages arrive by tiled TMA (.cta_group::2) that signals the LEADER's barrier from both CTAs.
enum : uint8_t { PS_ACT = 0,      // 1-2 weight chunks; A = act-buffer K = 32 chunks a_chunk0, a_chunk0 + 1
                 PS_ENCW = 1,     // [encoded block (A operand) | weight chunk] in one slot
                 PS_HEADS = 2 };  // four [16 x 32] chunks of the colour / (mu, sigma) heads; A = act chunks 0..3
enum : uint8_t { PF_FIRST = 1, PF_COMMIT_ACC = 2, PF_WAIT_ACT = 4, PF_WAIT_PREV = 8, PF_ENC_DONE = 16,
                 PF_SHARED = 32 };    // the stage serves BOTH super-tiles (weights only): loaded once, T0 then T1 consume it, T1 releases
// Tensor maps: 3-D views {512-byte row, rows of one chunk HALF, half index}.  The halves of a weight chunk lie one after the
// other (rank 0's N/2 rows, then rank 1's), so half 2 c + rank belongs to chunk c; a box of 3 halves traversed with element
// stride 2 delivers this CTA's halves of TWO consecutive chunks as one copy (contiguous in the slot).  PM_W128: 8 KB halves
// of the N = 256 chunks of the launched program; PM_W72: 4.5 KB halves of the view-branch chunks (N = 144); PM_W8: 512-byte
// halves of the four head chunks (box 7 -> four halves); PM_ENC: the 8 KB encoded blocks, one per copy.
enum : uint8_t { PM_W128 = 0, PM_W72 = 1, PM_W8 = 2, PM_ENC = 3,
                 PM_W128S = 4, PM_W72S = 5 };      // single halves (one-chunk stages)
constexpr int kPairMaps = 6;
struct PCopy {
    uint32_t idx;        // half / block index in the map; weights: + rank; encoded blocks: + tsel * mul + 8 * image
    uint8_t mul, map;
    uint16_t dst_off;    // byte offset inside the ring slot
};
struct __align__(16) PStage {
    uint8_t kind, n_chunks, last_nk16, flags;
    uint8_t a_chunk0, n_copies;
    uint16_t b_stride;   // bytes between the B blocks of consecutive chunks inside the slot
    uint32_t idesc;
    uint32_t tx_bytes;   // bytes landing in BOTH CTAs
    PCopy c[2];
};
constexpr int kMaxPStages = 48;
struct PProgram {
    int n_stages, n_phases;
    uint16_t phase_begin[kMaxEpis + 1];      // stages [phase_begin[e], phase_begin[e+1]) run before epilogue e
    uint8_t phase_fast[kMaxEpis + 3];        // a plain K = 256, N = 256 layer over the act buffer (four two-chunk stages) takes an unrolled
                                             // issue path: 1 = per super-tile, 2 = shared stages
    uint32_t phase_idx0[kMaxEpis + 1];       // fast phases: PM_W128 index of the first stage's copy (stage st: + 4 st, + rank)
    PStage st[kMaxPStages];
};

}  // namespace tcmlp
}  // namespace ddnerf
