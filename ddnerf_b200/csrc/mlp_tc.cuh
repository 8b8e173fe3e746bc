// Shared definitions of the tensor-core (bf16, tcgen05) MLP kernels: the static per-work-item
// PROGRAM (ring loads, MMA groups, epilogues) that the producer / MMA-issuer / epilogue warps of the
// chain kernel walk in lock step, and the global-memory image formats.
//
// Work item  = 256 sample rows = two 128-row tiles T0, T1 processed by one CTA.
// Ring       = 6 slots of 16 KB fed by bulk async copies (weights stages, encoded-feature blocks).
// Act buffer = per tile 64 KB: [4 k-blocks][128 rows][128 B], K-major SWIZZLE_128B, overwritten in
//              place by every layer's epilogue (the next layer's A operand).
// TMEM       = 512 columns: tile T owns columns [256 T, 256 T + 256) (fp32 accumulator 128 x 256).
#pragma once
#include "tc.cuh"

namespace ddnerf {
namespace tcmlp {

constexpr int kSlots = 6;
constexpr int kSlotBytes = 16384;
constexpr int kActBytes = 65536;
constexpr int kItemRows = 256;
constexpr int kEncItemBytes = 65536;     // [T0 b0,b1 | T1 b0,b1 | T0 b2 | T1 b2 | T0 dir | T1 dir], blocks [128 x 32] SW64
constexpr int kThreads = 320;            // warps 0-7 epilogue (4 per tile), 8 producer, 9 MMA issuer
constexpr int kMaxLoads = 112, kMaxMmas = 208, kMaxEpis = 12, kMaxPack = 112;

enum : uint8_t { F_WAIT_ACT = 1, F_FIRST = 2, F_COMMIT_ACC = 4, F_A_SLOT = 8, F_WAIT_PREV = 16 };
enum : uint8_t { EPI_ACT = 0, EPI_DIR = 1, EPI_OUT = 2, EPI_BWD_IN = 3, EPI_BWD_MASK = 4, EPI_BWD_HD = 5, EPI_BWD_LAST = 6 };
enum : uint8_t { LOAD_W = 0, LOAD_ENC = 1 };

struct Load { uint32_t kind, off, bytes; };
struct Mma {
    uint8_t tile, flags, nk16, idesc_sel;
    int16_t a_slot, b_slot;           // ring sequence numbers relative to the item (a_slot only with F_A_SLOT)
    uint32_t a_off, b_off;            // byte offsets (act buffer of the tile / inside the slot)
    int16_t rel0, rel1;               // ring sequence numbers released after this group (-1: none)
};
struct Epi {
    uint8_t mode, relu;
    int16_t save_layer;               // index into the activation / mask save areas, -1: none
    uint16_t ncols;
    uint16_t bias_off;                // float offset into the packed bias table
    uint32_t save_bytes;              // bytes of the act buffer stored when saving
};
struct Program {
    int n_loads, n_mmas, n_epis;
    Load loads[kMaxLoads];
    Mma mmas[kMaxMmas];
    Epi epis[kMaxEpis];
};
// one weight stage of the packed image: n_total rows x 32 k, SWIZZLE_64B
struct PackEntry { uint32_t dst_off; uint16_t kind, p, n_total, k0; };
struct PackTable { int n; uint32_t total_bytes; PackEntry e[kMaxPack]; };

enum : uint16_t { PK_FWD = 0, PK_FWD_DIR = 1, PK_FWD_HEADS = 2, PK_BWD = 3, PK_BWD_DIR = 4, PK_BWD_HEADS = 5 };

}  // namespace tcmlp
}  // namespace ddnerf
