// Shared helpers for the hyprec_b200 kernels (host + device).
#pragma once
#include <stdint.h>

#ifdef HYPREC_EMU
#include "emu_cuda.h"
#define HYPREC_HD inline
#else
#include <cuda_runtime.h>
#define HYPREC_HD __host__ __device__ __forceinline__
#define HYPREC_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace hyprec {

// Confidence weights of build_confidence_matrix (collaborative_filtering.py:135-142).
constexpr double kConfPos = 1.0;
constexpr double kConfNeg = 0.1;

// The per-row normal matrix is held as 8x8 tiles of its lower triangle, "structure of arrays
// over tiles": element (r, c) of tile t lives at [(r * 8 + c) * n_tiles + t], so that a warp
// whose lanes own consecutive tiles touches consecutive words (bank-conflict free).
constexpr int kTile = 8;
constexpr int kTileElems = kTile * kTile;

HYPREC_HD int tri(int i) { return i * (i + 1) / 2; }              // tiles in the first i block rows
HYPREC_HD int tile_id(int bi, int bj) { return tri(bi) + bj; }    // requires bj <= bi
// Block rows of the augmented system: k factor rows plus one row carrying the right-hand side.
HYPREC_HD int aug_blocks(int k) { return (k + 1 + kTile - 1) / kTile; }

// Inverse of tile_id on the lower triangle: linear index q -> (i, j), j <= i.
HYPREC_HD void tri_coords(int q, int* i_out, int* j_out) {
    int i = 0;
    while (tri(i + 1) <= q) ++i;   // callers build a table once per kernel; q < ~1000
    *i_out = i;
    *j_out = q - tri(i);
}

template <typename T>
struct Acc { typedef double type; };   // reductions are float64 in every precision mode

}  // namespace hyprec
