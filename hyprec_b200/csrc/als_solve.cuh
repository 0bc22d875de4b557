// Fused per-row ALS update: CSR gather of the fixed side's factor rows, rank-1 accumulation of
// the row's normal matrix, Cholesky factorisation and both triangular solves, all inside one
// CTA's shared memory -- the k x k system never exists in HBM.
//
// Replaces, for one row r of the side being solved (/root/reference/lib/collaborative_filtering.py):
//   confidence = build_confidence_matrix(r)                       :83, :91, :120-144
//   A = (Y.T * confidence).dot(Y) + lambda * I                    :84-85, :92
//   b = (ratings_r * confidence).dot(Y) [+ lambda * theta_r]      :86, :93-98
//   latent[r, :] = solve(A, b)                                    :85, :94, :98
// using the identity A = 0.1 * Y^T Y + 0.9 * sum_{j in R_r} y_j y_j^T + lambda * I,
// b = sum_{j in R_r} r_rj y_j (confidence is 1.0 on R_r and 0.1 elsewhere, and the rating is 0
// outside R_r).  `base` holds 0.1 * Y^T Y + lambda * I, shared by all rows (gram.cuh).
//
// During the factorisation the first 8 rows of the gather staging buffer hold the current
// panel (the 8 freshly solved columns of L for all rows below the diagonal tile) in the same
// interleaved order as gathered factor rows, so the trailing update reads its operands with
// broadcast / consecutive-word accesses exactly like the rank-1 accumulation.
//
// Layout: the (k+1) x (k+1) augmented matrix [A b; b^T .] is kept as 8x8 tiles of its lower
// triangle, element (r, c) of tile t at [(r*8+c) * n_tiles + t] (common.cuh).  Carrying b as
// row k makes the forward substitution part of the factorisation: after the last panel, row k
// holds z = L^-1 b.  Diagonal tiles are replaced by the inverse of their Cholesky factor, so
// the panel solve and the back substitution are small mat-vecs instead of dependent chains.
//
// One CTA = one row at a time; CTAs pull rows from an atomic counter (rows differ in degree).
//
// Low-rank mode (rows with deg << k).  With B = 0.1 G + lambda I = L L^T and the whitened factors
// Q = Y L^-T, the same normal matrix is A = L (I_k + 0.9 Q_r^T Q_r) L^T (Q_r = the deg gathered
// rows of Q), and by the Woodbury identity
//     x = L^-T y,   y = Q_r^T S^-1 (r - 0.9 Q_r p) + p,   S = I_deg + 0.9 Q_r Q_r^T,   p = lambda L^-1 theta_r
// (`prior` then holds the rows of theta L^-T; lambda is applied here)
// so only a deg x deg SPD system is factorised.  The kernel below is reused as is: the system
// dimension becomes deg, the rank-1 vectors are the k columns of Q_r, the base matrix is the
// identity, the right-hand side is r - 0.9 Q_r p, and the epilogue forms y; the caller turns the
// y rows into x with one small GEMM against L^-1.  Mathematically identical to the direct path,
// ~k/deg times fewer serial factorisation steps.
#pragma once
#include "common.cuh"
#ifndef HYPREC_EMU
#include "tc_syrk.cuh"
#endif

namespace hyprec {

constexpr int kSolveThreads = 256;

struct SolveLayout {
    int nb, n_tiles, kpad, ystride, chunk_rows;
    size_t off_a, off_y, off_z, off_x, off_idx, off_map, off_bar, total;
};

// Shared-memory plan for systems of dimension <= max_dim (k in direct mode, the largest row
// degree of the launch in low-rank mode) held as ts x ts tiles (8 for the k x k systems; 4 for
// the small low-rank systems, where 8 x 8 tiles would leave most threads without a tile and
// serialise 64 FMAs per rank-1 vector in the few that have one).  chunk_rows (>= ts: the buffer
// doubles as the panel buffer) rank-1 vectors are staged at a time; rows of the staging buffer are ystride apart
// (odd, so the low-rank gather's column-wise stores spread over banks).
HYPREC_HD int aug_blocks_ts(int dim, int ts) { return (dim + 1 + ts - 1) / ts; }

// stage_bytes > 0 (tensor-core launches, tc_syrk.cuh): the two operand staging stages alias the tile
// storage and the staging buffer from offset 0, the buffer is 1024-byte aligned by the kernel (the
// slack is part of `total`), and two mbarriers + the TMEM base slot follow the tile map.
HYPREC_HD SolveLayout solve_layout(int ts, int max_dim, int chunk_rows, bool a_in_smem, size_t elem, size_t stage_bytes = 0) {
    SolveLayout l;
    l.nb = aug_blocks_ts(max_dim, ts);
    l.n_tiles = tri(l.nb);
    l.kpad = l.nb * ts;
    l.ystride = l.kpad + 1;
    l.chunk_rows = chunk_rows;
    size_t o = 0;
    l.off_a = o;
    if (a_in_smem) o += elem * (size_t)(ts * ts) * l.n_tiles;
    l.off_y = o;
    o += elem * (size_t)chunk_rows * l.ystride;
    if (o < 2 * stage_bytes) o = 2 * stage_bytes;
    l.off_z = o;
    o += elem * (size_t)l.kpad;
    l.off_x = o;
    o += elem * (size_t)l.kpad;
    l.off_idx = (o + 7) & ~(size_t)7;
    o = l.off_idx + 4 * (size_t)l.kpad;
    l.off_map = o;
    o += 2 * (size_t)l.n_tiles;
    l.off_bar = (o + 7) & ~(size_t)7;
    if (stage_bytes > 0) o = l.off_bar + 32 + 1024;
    l.total = (o + 15) & ~(size_t)15;
    return l;
}

template <typename T>
struct SolveArgs {
    const int64_t* indptr;   // CSR of the side being solved, indexed by global row id
    const int32_t* indices;  // ids into `fixed`
    const T* values;         // ratings (b = sum r y); never null
    const T* fixed;          // [n_fixed][k]
    T* latent;               // [n_rows][k], rows [row_lo, row_hi) are written
    const T* base;           // tiled 0.1 * G + lambda * I (build_base_tiles_kernel)
    const T* prior;          // [n_rows][k] theta for update_with_items, or null
    T* a_scratch;            // per-CTA tile storage in global memory, or null => shared memory
    int* work_counter;       // zeroed before launch
    int* status;             // set to 1 when a solution is not finite
    T lambda;
    int k;
    int64_t row_lo, row_hi;
    int chunk_rows;
    // ---- work list / modes --------------------------------------------------------------------
    const int32_t* row_list;  // optional: row ids to process instead of [row_lo, row_hi)
    int64_t n_list;
    int max_dim;              // system dimension the shared-memory layout is sized for
    int lowrank;              // 0: direct k x k systems; 1: deg x deg systems on whitened factors
    T* ymat;                  // low-rank output, COMPACT: work item w of the launch writes row ymat_row0 + w
    int64_t ymat_row0;
    T* export_tiles;          // direct mode: factored tiles of the processed row (single-row launch)
    int tmem_cols;            // tensor-core launches: TMEM columns to allocate (tc::syrk_tmem_cols)
    long long* prof;          // optional: per-phase clock64() totals of block 0 (tools/solve_phases.py)
    // optional [n_rows]: sum_j r_j (fixed_j . x) of the row -- the RMSE cross term of evaluator.py:249 as a by-product
    // of the solve (direct: b.x = |L^-1 b|^2; low-rank: (|r|^2 - |z|^2) / 0.9); only meaningful without a prior
    double* cross_out = nullptr;
};

// float64 dot product of two k-vectors by one warp (lane-strided partial sums, xor-shuffle tree)
template <typename TA, typename TB>
__device__ __forceinline__ double warp_dot_t(const TA* __restrict__ a, const TB* __restrict__ b, int k, int lane) {
    double s = 0.0;
    for (int l = lane; l < k; l += 32) s += (double)a[l] * (double)b[l];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    return s;
}

// 1/sqrt(x) in float64 from a float32 seed and two Newton steps (2^-22 -> ~2^-44 -> rounding level):
// a far shorter dependent chain than the library rsqrt, which sits on the factorisation's
// critical path once per column.  Falls back to rsqrt outside the float32 range.
__device__ __forceinline__ double inv_sqrt(double x) {
#ifdef HYPREC_EMU
    return rsqrt(x);
#else
    if (!(x > 1e-30 && x < 1e30)) return rsqrt(x);
    double y = (double)rsqrtf((float)x);
    const double hx = 0.5 * x;
    y = y * (1.5 - hx * y * y);
    y = y * (1.5 - hx * y * y);
    return y;
#endif
}
__device__ __forceinline__ float inv_sqrt(float x) { return rsqrtf(x); }

// Cholesky of the leading nvalid x nvalid block of a TS x TS tile held in registers (lower
// triangle); the rows below it are carried along as in a panel factorisation (that is where the
// right-hand-side row picks up its part of z).  The diagonal is left INVERTED.
template <typename T, int TS>
__device__ __forceinline__ void chol_tile_factor(T (&a)[TS][TS], int nvalid) {
#pragma unroll
    for (int j = 0; j < TS; ++j) {
        if (j < nvalid) {
            const T inv = inv_sqrt(a[j][j]);
            a[j][j] = inv;
#pragma unroll
            for (int i = j + 1; i < TS; ++i) a[i][j] *= inv;
#pragma unroll
            for (int l = j + 1; l < TS; ++l)
#pragma unroll
                for (int i = l; i < TS; ++i) a[i][l] -= a[i][j] * a[l][j];
        }
    }
}

// In-place inverse of the leading lower-triangular nvalid x nvalid block (diagonal already
// inverted): column j of L^-1 is -L22^-1 L[j+1:, j] / L[j][j].
template <typename T, int TS>
__device__ __forceinline__ void chol_tile_invert(T (&a)[TS][TS], int nvalid) {
#pragma unroll
    for (int j = TS - 1; j >= 0; --j) {
        if (j < nvalid) {
            const T njj = -a[j][j];
#pragma unroll
            for (int i = TS - 1; i > j; --i) {
                if (i < nvalid) {
                    T s = T(0);
#pragma unroll
                    for (int l = j + 1; l <= i; ++l) s += a[i][l] * a[l][j];
                    a[i][j] = s * njj;
                }
            }
        }
    }
}

// The serial part of a diagonal tile, one thread, tile in registers; out of line so that its
// register needs stay out of the callers' accumulator tiles.
#ifndef HYPREC_EMU
template <typename T, int TS, bool GA>
__device__ __noinline__ void factor_diag_tile_serial(T* A, int NT, int t, int p, int nvalid, T* zvec) {
    if (!GA) __builtin_assume(__isShared(A));   // shared-memory loads/stores across the call boundary
    __builtin_assume(__isShared(zvec));
#else
template <typename T, int TS, bool GA>
void factor_diag_tile_serial(T* A, int NT, int t, int p, int nvalid, T* zvec) {
#endif
    T a[TS][TS];
#pragma unroll
    for (int r = 0; r < TS; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) a[r][c] = A[(size_t)(r * TS + c) * NT + t];
    chol_tile_factor<T, TS>(a, nvalid);
    if (nvalid < TS) {
#pragma unroll
        for (int r = 1; r < TS; ++r)
            if (r == nvalid) {
#pragma unroll
                for (int j = 0; j < r; ++j) zvec[p * TS + j] = a[r][j];
            }
    }
    chol_tile_invert<T, TS>(a, nvalid);
#pragma unroll
    for (int r = 0; r < TS; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) A[(size_t)(r * TS + c) * NT + t] = a[r][c];
}

// Factor diagonal tile `p` (tile id `t` of the tiled matrix A) in place: Cholesky of its leading
// block (rows below it carried along; the right-hand-side row's z entries are published to zvec
// when it sits in this tile), then that block is replaced by its inverse, diagonal kept inverted.
// Called by all 32 lanes of one warp.  With `panel` != null the tile first receives its trailing
// update from the current panel (A_pp -= P_p P_p^T) spread over the lanes -- lane l holds column
// l % TS of rows l / TS + g * (32 / TS).  The factorisation itself is a chain of dependent
// rsqrt / FMA steps: measured on B200, one thread with the tile in registers runs it faster than a
// shuffle-based warp version, so lane 0 does it while the other warps carry the trailing update.
template <typename T, int TS, bool GA>
__device__ __forceinline__ void factor_diag_tile_warp(T* A, int NT, int t, int p, int dim, T* zvec, int lane,
                                                      const T* panel = nullptr, int ystride = 0, int nb = 0) {
    int nvalid = dim - p * TS;
    if (nvalid > TS) nvalid = TS;
    if (panel != nullptr) {
        constexpr int RPP = 32 / TS;                   // rows covered per pass of 32 lanes
        constexpr int EPL = (TS * TS + 31) / 32;       // elements per lane
        const int c = lane % TS, r0 = lane / TS;
        T a[EPL];
#pragma unroll
        for (int g = 0; g < EPL; ++g) {
            const int r = r0 + g * RPP;
            a[g] = (r < TS) ? A[(size_t)(r * TS + c) * NT + t] : T(0);
        }
#pragma unroll
        for (int kk = 0; kk < TS; ++kk) {
            const T* prow = panel + (size_t)kk * ystride;
            const T lc = prow[c * nb + p];
#pragma unroll
            for (int g = 0; g < EPL; ++g) {
                const int r = r0 + g * RPP;
                if (r < TS) a[g] -= prow[r * nb + p] * lc;
            }
        }
#pragma unroll
        for (int g = 0; g < EPL; ++g) {
            const int r = r0 + g * RPP;
            if (r < TS) A[(size_t)(r * TS + c) * NT + t] = a[g];
        }
        __syncwarp();
    }
    if (nvalid > 0 && lane == 0) factor_diag_tile_serial<T, TS, GA>(A, NT, t, p, nvalid, zvec);
    __syncwarp();
}

// A_tile(bi, bj) -= P_bi P_bj^T with the current panel P (TS columns of L) read from the panel buffer
template <typename T, int TS>
__device__ __forceinline__ void trailing_update_tile(T* A, int NT, const T* ybuf, int ystride, int nb, int bi, int bj) {
    const int t = tile_id(bi, bj);
    T acc[TS][TS];
#pragma unroll
    for (int r = 0; r < TS; ++r)
#pragma unroll
        for (int c = 0; c < TS; ++c) acc[r][c] = A[(size_t)(r * TS + c) * NT + t];
#pragma unroll 2
    for (int kk = 0; kk < TS; ++kk) {
        const T* prow = ybuf + (size_t)kk * ystride;
        T li[TS], lj[TS];
#pragma unroll
        for (int r = 0; r < TS; ++r) li[r] = prow[r * nb + bi];
#pragma unroll
        for (int c = 0; c < TS; ++c) lj[c] = prow[c * nb + bj];
#pragma unroll
        for (int r = 0; r < TS; ++r)
#pragma unroll
            for (int c = 0; c < TS; ++c) acc[r][c] -= li[r] * lj[c];
    }
#pragma unroll
    for (int r = 0; r < TS; ++r)
#pragma unroll
        for (int c = 0; c < TS; ++c) A[(size_t)(r * TS + c) * NT + t] = acc[r][c];
}

#ifdef HYPREC_EMU
#define HYPREC_PROF(slot)
#else
// phase timing for tools/solve_phases.py: thread 0 of block 0 accumulates clock64() deltas
#define HYPREC_PROF(slot)                                                            \
    if (p.prof != nullptr && blockIdx.x == 0 && tid == 0) {                          \
        const long long now_ = clock64();                                            \
        p.prof[slot] += now_ - prof_t;                                               \
        prof_t = now_;                                                               \
    }
#endif

// Sum over the CTA of one value per thread in a FIXED order (shuffle tree per warp, warp partials added by thread 0
// in warp order): deterministic, unlike floating-point atomics.  red: >= 33 doubles of shared memory.  Every thread
// returns the total.
__device__ __forceinline__ double block_sum_fixed(double v, double* red) {
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32, nwarps = (blockDim.x + 31) / 32;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < nwarps; ++w) s += red[w];
        red[32] = s;
    }
    __syncthreads();
    return red[32];
}

// TC (float32, TS == 8, low-rank launches, sm_100a only): the rank-k accumulation runs on the tensor
// cores (tc_syrk.cuh) instead of the SIMT rank-1 loop; everything after it is shared.
// GA: the tiles live in per-CTA global scratch (p.a_scratch) instead of shared memory -- a separate
// instantiation so that the common case addresses its tiles with shared-memory instructions.
template <typename T, int TS, bool TC = false, bool GA = false>
__global__ void __launch_bounds__(kSolveThreads, TS == 4 ? 2 : 1) als_solve_rows_kernel(SolveArgs<T> p) {
    HYPREC_DYN_SMEM(smem_dyn);
    unsigned char* smem_raw = smem_dyn;
#ifndef HYPREC_EMU
    // SWIZZLE_128B staging needs a 1024-byte aligned base (offset arithmetic keeps the pointer a
    // shared-memory pointer for the compiler)
    if (TC) smem_raw = smem_dyn + ((1024u - (tc::smem_u32(smem_dyn) & 1023u)) & 1023u);
#endif
    long long prof_t = 0;
    __shared__ long long s_row;
    __shared__ double s_red[33];
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int lane = tid % 32, warp = tid / 32, nwarps = nthreads / 32;
    const int k = p.k;
    const bool lowrank = p.lowrank != 0;
#ifndef HYPREC_EMU
    const SolveLayout lay = solve_layout(TS, p.max_dim, p.chunk_rows, !GA, sizeof(T),
                                         TC ? tc::syrk_stage_bytes(p.max_dim) : 0);
#else
    const SolveLayout lay = solve_layout(TS, p.max_dim, p.chunk_rows, !GA, sizeof(T));
#endif
    const int CH = lay.chunk_rows, ystride = lay.ystride;
    T* A = GA ? p.a_scratch + (size_t)blockIdx.x * (TS * TS) * lay.n_tiles : (T*)(smem_raw + lay.off_a);
    T* ybuf = (T*)(smem_raw + lay.off_y);
    T* zvec = (T*)(smem_raw + lay.off_z);
    T* xvec = (T*)(smem_raw + lay.off_x);
    int* idxs = (int*)(smem_raw + lay.off_idx);
    unsigned char* tmap = smem_raw + lay.off_map;
    const T pos_scale = T(kConfPos - kConfNeg);
    const int64_t n_work = p.row_list ? p.n_list : p.row_hi - p.row_lo;

    for (int t = tid; t < lay.n_tiles; t += nthreads) {
        int bi, bj;
        tri_coords(t, &bi, &bj);
        tmap[2 * t] = (unsigned char)bi;
        tmap[2 * t + 1] = (unsigned char)bj;
    }
#ifndef HYPREC_EMU
    tc::SyrkState tcs;
    if (TC) {
        tcs.bars = (uint64_t*)(smem_raw + lay.off_bar);
        uint32_t* tmem_slot = (uint32_t*)(tcs.bars + 2);
        tcs.phase[0] = tcs.phase[1] = 0;
        tcs.stage = smem_raw;
        tcs.rows_pad = tc::syrk_rows_pad(p.max_dim);
        if (tid == 0) {
            tc::mbar_init(&tcs.bars[0], 1);
            tc::mbar_init(&tcs.bars[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (warp == 0) tc::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        tcs.tmem_base = *tmem_slot;
    }
#endif

    while (true) {
        __syncthreads();   // previous row finished with A / zvec / xvec / idxs / s_row; tmap visible
        if (tid == 0) s_row = (long long)atomicAdd(p.work_counter, 1);
        __syncthreads();
        if (s_row >= n_work) break;
        const int64_t row = p.row_list ? (int64_t)p.row_list[s_row] : p.row_lo + s_row;
#ifndef HYPREC_EMU
        if (p.prof != nullptr && blockIdx.x == 0 && tid == 0) prof_t = clock64();
#endif
        const int64_t lo = p.indptr[row];
        const int deg = (int)(p.indptr[row + 1] - lo);
        T* out = lowrank ? p.ymat + (p.ymat_row0 + s_row) * k : p.latent + row * k;
        const T* prior_row = p.prior ? p.prior + row * k : nullptr;
        if (deg == 0 && (lowrank || (p.prior == nullptr && p.export_tiles == nullptr))) {
            // direct: b == 0 => the solve returns exactly 0; low-rank: y = p
            for (int c = tid; c < k; c += nthreads) out[c] = (lowrank && prior_row) ? p.lambda * prior_row[c] : T(0);
            continue;
        }
        // dimension of this row's system, number of rank-1 vectors that build it
        const int dim = lowrank ? deg : k;
        const int nvec = lowrank ? k : deg;
        const int nb = aug_blocks_ts(dim, TS), NT = tri(nb), kpad = nb * TS;
        const int nbk = (dim + TS - 1) / TS;   // block columns that hold system columns

        if (lowrank) {
            // positives' ids; right-hand side r - 0.9 * Q_r p kept in xvec until the finalize step
            for (int i = tid; i < dim; i += nthreads) idxs[i] = p.indices[lo + i];
            __syncthreads();
            for (int i = warp; i < dim; i += nwarps) {
                T v = p.values[lo + i];
                if (prior_row) {
                    const double d = warp_dot_t(p.fixed + (int64_t)idxs[i] * k, prior_row, k, lane);
                    v -= pos_scale * p.lambda * (T)d;
                }
                if (lane == 0) xvec[i] = v;
            }
        }
        double rr = 0.0;       // |r|^2 of the row's ratings (low-rank cross term)
        if (p.cross_out != nullptr && lowrank) {
            __syncthreads();
            double part = 0.0;
            for (int i = tid; i < dim; i += nthreads) part += (double)xvec[i] * (double)xvec[i];
            rr = block_sum_fixed(part, s_red);
        }

#ifndef HYPREC_EMU
        if constexpr (TC) {
            // ---- the rank-k accumulation on the tensor cores, then the rhs row and the padding --
            __syncthreads();   // idxs / xvec complete (low-rank); previous row done with zvec / xvec
            if (lowrank) {
                // S = I + 0.9 Q_r Q_r^T; rhs row = r - 0.9 Q_r p (already in xvec)
                tc::syrk_lowrank(p.fixed, idxs, dim, k, tcs, (float*)A, NT, (float)pos_scale);
            } else if (deg > 0) {
                // 0.9 sum y y^T (the shared base is added below); xvec = sum r y
                int batch = (kpad / tc::kSyrkKB) * tc::kSyrkKB;
                if (batch > 256) batch = 256;
                tc::syrk_direct(p.fixed, p.indices, p.values, lo, deg, k, tcs, (float*)A, NT, (float)pos_scale, idxs,
                                (float*)zvec, batch, (float*)xvec);
            } else {
                for (int e = tid; e < TS * TS * NT; e += nthreads) A[e] = T(0);
                for (int c = tid; c < k; c += nthreads) xvec[c] = T(0);
                __syncthreads();
            }
            for (int e = tid; e < nb * TS * TS; e += nthreads) {
                const int bj = e / (TS * TS), r = (e / TS) % TS, c = e % TS;
                const int i = (nb - 1) * TS + r, j = bj * TS + c;
                if (i >= dim) {
                    T v = T(0);
                    if (i == dim && j < dim) {
                        v = xvec[j];
                        if (!lowrank && prior_row != nullptr) v += p.lambda * prior_row[j];
                    }
                    A[(size_t)(r * TS + c) * NT + tile_id(nb - 1, bj)] = v;
                }
            }
            __syncthreads();
        } else
#endif
        {
        // ---- gather + rank-1 accumulation, CH vectors at a time --------------------------------
        const int nchunks = nvec == 0 ? 1 : (nvec + CH - 1) / CH;
        for (int ch = 0; ch < nchunks; ++ch) {
            const int c0 = ch * CH;
            int cr = nvec - c0;
            if (cr > CH) cr = CH;
            if (!lowrank) {
                // one warp per gathered row: coalesced read of the factor row, stored with
                // column 8*b+c at [c * nb + b] so tile owners read consecutive words
                for (int t = warp; t < cr; t += nwarps) {
                    const int64_t src = (int64_t)p.indices[lo + c0 + t] * k;
                    const T rating = p.values[lo + c0 + t];
                    T* dst = ybuf + (size_t)t * ystride;
                    for (int col = lane; col < kpad; col += 32) {
                        T v = T(0);
                        if (col < k) v = p.fixed[src + col];
                        else if (col == k) v = rating;
                        dst[(col % TS) * nb + col / TS] = v;
                    }
                }
            } else {
                // vector t = column c0+t of Q restricted to this row's positives: one warp per
                // positive reads cr consecutive values of its Q row
#pragma unroll 4
                for (int i = warp; i < dim; i += nwarps) {
                    const T* src = p.fixed + (int64_t)idxs[i] * k + c0;
                    const int pos = (i % TS) * nb + i / TS;
                    for (int t = lane; t < cr; t += 32) ybuf[(size_t)t * ystride + pos] = src[t];
                }
                const int npad = kpad - dim;   // augmented slot + padding carry zeros
                for (int e = tid; e < cr * npad; e += nthreads) {
                    const int t = e / npad, i = dim + e % npad;
                    ybuf[(size_t)t * ystride + (i % TS) * nb + i / TS] = T(0);
                }
            }
            __syncthreads();
            const bool last = (ch == nchunks - 1);
#pragma unroll 1
            for (int q = tid; q < NT; q += nthreads) {
                const int bi = tmap[2 * q], bj = tmap[2 * q + 1];
                T acc[TS][TS];
                if (ch == 0) {
#pragma unroll
                    for (int r = 0; r < TS; ++r)
#pragma unroll
                        for (int c = 0; c < TS; ++c) acc[r][c] = T(0);
                } else {
#pragma unroll
                    for (int r = 0; r < TS; ++r)
#pragma unroll
                        for (int c = 0; c < TS; ++c) acc[r][c] = A[(size_t)(r * TS + c) * NT + q];
                }
#pragma unroll 2
                for (int t = 0; t < cr; ++t) {
                    const T* yrow = ybuf + (size_t)t * ystride;
                    T vi[TS], vj[TS];
#pragma unroll
                    for (int r = 0; r < TS; ++r) vi[r] = yrow[r * nb + bi];
#pragma unroll
                    for (int c = 0; c < TS; ++c) vj[c] = yrow[c * nb + bj];
#pragma unroll
                    for (int r = 0; r < TS; ++r)
#pragma unroll
                        for (int c = 0; c < TS; ++c) acc[r][c] += vi[r] * vj[c];
                }
                if (last) {
                    // direct:   A = base + 0.9 * sum y y^T, rhs row = sum r y (+ lambda theta)
                    // low-rank: S = I + 0.9 * Q_r Q_r^T,    rhs row = r - 0.9 Q_r p
#pragma unroll
                    for (int r = 0; r < TS; ++r) {
                        const int i = bi * TS + r;
#pragma unroll
                        for (int c = 0; c < TS; ++c) {
                            const int j = bj * TS + c;
                            T v;
                            if (i < dim) {
                                // (direct mode: the shared base 0.1 G + lambda I is added below, in one
                                // coalesced pass, instead of 64 scattered loads per tile here)
                                v = pos_scale * acc[r][c] + ((lowrank && i == j) ? T(1) : T(0));
                            } else if (i == dim) {
                                if (lowrank) {
                                    v = j < dim ? xvec[j] : T(0);
                                } else {
                                    v = acc[r][c];
                                    if (prior_row != nullptr && j < k) v += p.lambda * prior_row[j];
                                }
                            } else {
                                v = T(0);
                            }
                            acc[r][c] = v;
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < TS; ++r)
#pragma unroll
                    for (int c = 0; c < TS; ++c) A[(size_t)(r * TS + c) * NT + q] = acc[r][c];
            }
            __syncthreads();
        }
        }
        HYPREC_PROF(0)   // gather + rank-1 accumulation
        if (!lowrank) {
#pragma unroll 8
            for (int e = tid; e < TS * TS * NT; e += nthreads) A[e] += p.base[e];
            __syncthreads();
        }
        HYPREC_PROF(1)   // base add
        if (warp == 0) factor_diag_tile_warp<T, TS, GA>(A, NT, 0, 0, dim, zvec, lane);
        __syncthreads();
        HYPREC_PROF(2)   // first diagonal tile

        // ---- right-looking blocked Cholesky, 8-column panels ---------------------------------
        for (int pnl = 0; pnl < nbk; ++pnl) {
            // panel solve: rows below the diagonal tile (system rows and the rhs row `dim`) times
            // the transposed inverse of the diagonal tile's factor
            const int tdiag = tile_id(pnl, pnl);
            for (int i = TS * (pnl + 1) + tid; i <= dim; i += nthreads) {
                const int r = i % TS;
                const int t = tile_id(i / TS, pnl);
                T x[TS], y[TS];
#pragma unroll
                for (int c = 0; c < TS; ++c) x[c] = A[(size_t)(r * TS + c) * NT + t];
#pragma unroll
                for (int c = 0; c < TS; ++c) {
                    T s = T(0);
#pragma unroll
                    for (int j = 0; j <= c; ++j) s += x[j] * A[(size_t)(c * TS + j) * NT + tdiag];
                    y[c] = s;
                }
#pragma unroll
                for (int c = 0; c < TS; ++c) {
                    A[(size_t)(r * TS + c) * NT + t] = y[c];
                    ybuf[(size_t)c * ystride + r * nb + i / TS] = y[c];   // panel, row 8*b+r at [r*nb+b]
                }
                if (i == dim) {
#pragma unroll
                    for (int c = 0; c < TS; ++c) zvec[pnl * TS + c] = y[c];
                }
            }
            __syncthreads();
            HYPREC_PROF(3)   // panel solves
            // trailing update.  Warp 0 updates only the next diagonal tile (spread over its lanes) and
            // factors it straight away; the other warps update every other trailing tile meanwhile, so
            // the factorisation of the next panel's diagonal runs beside the update instead of after it.
            const int ntr = nb - 1 - pnl;
            if (ntr > 0) {
                const int ntiles = tri(ntr);
                if (warp == 0) {
                    factor_diag_tile_warp<T, TS, GA>(A, NT, tile_id(pnl + 1, pnl + 1), pnl + 1, dim, zvec, lane, ybuf, ystride, nb);
                    HYPREC_PROF(5)   // warp 0: next diagonal tile (update + factor + inverse)
                    if (nwarps == 1) {
#pragma unroll 1
                        for (int q = 1 + lane; q < ntiles; q += 32)
                            trailing_update_tile<T, TS>(A, NT, ybuf, ystride, nb, pnl + 1 + tmap[2 * q], pnl + 1 + tmap[2 * q + 1]);
                    }
                } else {
#pragma unroll 1
                    for (int q = 1 + tid - 32; q < ntiles; q += nthreads - 32)
                        trailing_update_tile<T, TS>(A, NT, ybuf, ystride, nb, pnl + 1 + tmap[2 * q], pnl + 1 + tmap[2 * q + 1]);
                }
            }
            __syncthreads();
            HYPREC_PROF(6)   // warp 0 waiting for the rest of the trailing update
        }

        if (!lowrank && p.export_tiles != nullptr) {
            for (int e = tid; e < (TS * TS) * NT; e += nthreads) p.export_tiles[e] = A[e];
        }

        if (p.cross_out != nullptr) {
            // z = L^-1 b is complete: the row's share of sum_nz r (u . v)
            double part = 0.0;
            for (int i = tid; i < dim; i += nthreads) part += (double)zvec[i] * (double)zvec[i];
            const double zz = block_sum_fixed(part, s_red);
            if (tid == 0) p.cross_out[row] = lowrank ? (rr - zz) / (double)pos_scale : zz;
        }

        // ---- back substitution L^T x = z, one block column at a time -------------------------
        for (int bj = nbk - 1; bj >= 0; --bj) {
            int nv = dim - TS * bj;
            if (nv > TS) nv = TS;
            const int td = tile_id(bj, bj);
            if (tid < nv) {
                T s = T(0);
                for (int j = tid; j < nv; ++j) s += A[(size_t)(j * TS + tid) * NT + td] * zvec[TS * bj + j];
                xvec[TS * bj + tid] = s;
            }
            __syncthreads();
            for (int jp = tid; jp < TS * bj; jp += nthreads) {
                const int t = tile_id(bj, jp / TS), c = jp % TS;
                T s = zvec[jp];
                for (int i = 0; i < nv; ++i) s -= A[(size_t)(i * TS + c) * NT + t] * xvec[TS * bj + i];
                zvec[jp] = s;
            }
            __syncthreads();
        }
        HYPREC_PROF(7)   // back substitution
#ifndef HYPREC_EMU
        if (p.prof != nullptr && blockIdx.x == 0 && tid == 0) {
            p.prof[8] += 1;
            p.prof[9] += deg;
        }
#endif
        if (!lowrank) {
            for (int c = tid; c < k; c += nthreads) {
                const T v = xvec[c];
                if (!(v - v == T(0))) *p.status = 1;   // inf or nan: singular system (lambda == 0)
                out[c] = v;
            }
        } else {
            // y = Q_r^T t + p  (second, L2-resident pass over the row's Q rows; coalesced along a)
            for (int a = tid; a < k; a += nthreads) {
                T s = prior_row ? p.lambda * prior_row[a] : T(0);
                int i = 0;
                for (; i + 4 <= dim; i += 4) {   // four loads in flight per thread
                    const T q0 = p.fixed[(int64_t)idxs[i] * k + a], q1 = p.fixed[(int64_t)idxs[i + 1] * k + a];
                    const T q2 = p.fixed[(int64_t)idxs[i + 2] * k + a], q3 = p.fixed[(int64_t)idxs[i + 3] * k + a];
                    s += q0 * xvec[i];
                    s += q1 * xvec[i + 1];
                    s += q2 * xvec[i + 2];
                    s += q3 * xvec[i + 3];
                }
                for (; i < dim; ++i) s += p.fixed[(int64_t)idxs[i] * k + a] * xvec[i];
                if (!(s - s == T(0))) *p.status = 1;
                out[a] = s;
            }
        }
        HYPREC_PROF(10)   // solution write / low-rank y = Q_r^T t + p
    }
#ifndef HYPREC_EMU
    if (TC) {
        tc::tc_fence_before();
        __syncthreads();
        if (warp == 0) tc::tmem_dealloc(tcs.tmem_base, (uint32_t)p.tmem_cols);
    }
#endif
}

// ------------------------------------------------------------------------------------------------
// L^-1 (dense, lower triangular) and its transpose from the factored tiles of a k x k matrix
// (als_solve_rows_kernel with export_tiles: off-diagonal tiles hold L, diagonal tiles hold the
// inverse of their 8x8 factor).  Block column J of X = L^-1:
//     X_JJ = D_J^-1,   X_IJ = -D_I^-1 * sum_{Q=J}^{I-1} L_IQ X_QJ   (I > J)
// grid = block columns, 64 threads = the 8x8 elements of the tile being formed; the exported
// tiles are staged in shared memory when they fit (tiles_in_smem), else read through L2.
// Outputs are written with leading dimension ld (>= k): a diagonal block of a larger matrix.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(64) tri_inverse_kernel(const T* __restrict__ tiles, int k, int tiles_in_smem,
                                                         T* __restrict__ linv, T* __restrict__ linv_t, int ld) {
    HYPREC_DYN_SMEM(smem_raw);
    const int nb_src = aug_blocks(k), NT = tri(nb_src);
    const int nbk = (k + kTile - 1) / kTile;
    const int J = blockIdx.x;
    const int tid = threadIdx.x, r = tid / kTile, c = tid % kTile;
    T* xcol = (T*)smem_raw;                       // [nbk][64] tiles of block column J
    T* sbuf = xcol + (size_t)nbk * kTileElems;    // [64]
    T* ls = sbuf + kTileElems;                    // [64 * NT] optional copy of the tiles
    const T* L = tiles;
    if (tiles_in_smem) {
        for (int e = tid; e < kTileElems * NT; e += 64) ls[e] = tiles[e];
        L = ls;
    }
    __syncthreads();
    for (int I = J; I < nbk; ++I) {
        T v;
        if (I == J) {
            v = (c <= r) ? L[(size_t)(r * kTile + c) * NT + tile_id(J, J)] : T(0);
        } else {
            T s = T(0);
            for (int Q = J; Q < I; ++Q) {
                const int t = tile_id(I, Q);
#pragma unroll
                for (int kk = 0; kk < kTile; ++kk)
                    s += L[(size_t)(r * kTile + kk) * NT + t] * xcol[(size_t)Q * kTileElems + kk * kTile + c];
            }
            sbuf[r * kTile + c] = s;
            __syncthreads();
            const int td = tile_id(I, I);
            T acc = T(0);
            for (int rr = 0; rr <= r; ++rr) acc += L[(size_t)(r * kTile + rr) * NT + td] * sbuf[rr * kTile + c];
            v = -acc;
        }
        const int gi = I * kTile + r, gj = J * kTile + c;
        if (gi >= k || gj >= k) v = T(0);
        xcol[(size_t)I * kTileElems + r * kTile + c] = v;
        if (gi < k && gj < k) {
            linv[(size_t)gi * ld + gj] = v;
            linv_t[(size_t)gj * ld + gi] = v;
            if (I == J && c > r) {   // strictly upper part of the diagonal block is zero in L^-1
                linv[(size_t)gi * ld + gj] = T(0);
                linv_t[(size_t)gj * ld + gi] = T(0);
            }
        }
        __syncthreads();
    }
    // zero the strictly upper block triangle (tiles with I < J) of linv / lower of linv_t
    for (int I = 0; I < J; ++I) {
        const int gi = I * kTile + r, gj = J * kTile + c;
        if (gi < k && gj < k) {
            linv[(size_t)gi * ld + gj] = T(0);
            linv_t[(size_t)gj * ld + gi] = T(0);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Dense helpers for the blocked factorisation of B = 0.1 G + lambda I when k x k float64 tiles do
// not fit one SM's shared memory (k > ~230): B is split into r x r blocks of <= 192, diagonal blocks
// go through als_solve_rows_kernel (export) + tri_inverse_kernel, the rest is small GEMMs.
// ------------------------------------------------------------------------------------------------

// B = 0.1 * G + lambda * I as a dense float64 matrix.
template <typename T>
__global__ void build_base_dense_kernel(const T* __restrict__ g, int k, double lambda, double* __restrict__ b) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= k * k) return;
    const int i = idx / k, j = idx % k;
    b[idx] = kConfNeg * (double)g[idx] + (i == j ? lambda : 0.0);
}

// Tiled (8x8, structure-of-arrays) copy of the diagonal block B[off:off+bs, off:off+bs] with a zero
// right-hand-side row: the `base` input of a direct-mode factorisation with k := bs.
__global__ void tiles_from_dense_kernel(const double* __restrict__ b, int ld, int off, int bs, double* __restrict__ tiles) {
    const int nb = aug_blocks(bs), n_tiles = tri(nb);
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= kTileElems * n_tiles) return;
    const int e = idx / n_tiles, t = idx % n_tiles;
    int bi, bj;
    tri_coords(t, &bi, &bj);
    const int i = bi * kTile + e / kTile, j = bj * kTile + e % kTile;
    tiles[idx] = (i < bs && j < bs) ? b[(size_t)(off + i) * ld + off + j] : 0.0;
}

// C[M x N] = alpha * A[M x K] * op(B) + beta * C, row-major with leading dimensions; op(B) = B^T for
// B stored [N x K] (trans_b) or B stored [K x N].  Small matrices (<= 300): 16 x 16 tiles.
__global__ void dgemm_small_kernel(int M, int N, int K, const double* __restrict__ A, int lda,
                                   const double* __restrict__ B, int ldb, int trans_b, double* __restrict__ C, int ldc,
                                   double alpha, double beta) {
    __shared__ double as[16][17], bs[16][17];
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const int row = blockIdx.y * 16 + ty, col = blockIdx.x * 16 + tx;
    double acc = 0.0;
    for (int k0 = 0; k0 < K; k0 += 16) {
        as[ty][tx] = (row < M && k0 + tx < K) ? A[(size_t)row * lda + k0 + tx] : 0.0;
        // bs[kk][n]
        const int n_load = blockIdx.x * 16 + ty;   // for trans_b: row of B = output column
        if (trans_b) bs[tx][ty] = (n_load < N && k0 + tx < K) ? B[(size_t)n_load * ldb + k0 + tx] : 0.0;
        else bs[ty][tx] = (k0 + ty < K && col < N) ? B[(size_t)(k0 + ty) * ldb + col] : 0.0;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) acc += as[ty][kk] * bs[kk][tx];
        __syncthreads();
    }
    if (row < M && col < N) {
        double* c = C + (size_t)row * ldc + col;
        *c = alpha * acc + (beta != 0.0 ? beta * *c : 0.0);
    }
}

// element-wise conversion between the factor dtype and float64 / float32 work copies
template <typename TIn, typename TOut>
__global__ void cast_kernel(const TIn* __restrict__ src, int64_t count, TOut* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) dst[i] = (TOut)src[i];
}

// dst[j][i] = src[i][j] for a k x k matrix (L^-T from L^-1).
__global__ void transpose_dense_kernel(const double* __restrict__ src, int k, double* __restrict__ dst) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= k * k) return;
    const int i = idx / k, j = idx % k;
    dst[(size_t)j * k + i] = src[idx];
}

}  // namespace hyprec
