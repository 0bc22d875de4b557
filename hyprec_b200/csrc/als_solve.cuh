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
#pragma once
#include "common.cuh"

namespace hyprec {

constexpr int kSolveThreads = 256;

struct SolveLayout {
    int nb, n_tiles, kpad, chunk_rows;
    size_t off_a, off_y, off_z, off_x, off_map, total;
};

HYPREC_HD SolveLayout solve_layout(int k, int chunk_rows, bool a_in_smem, size_t elem) {
    SolveLayout l;
    l.nb = aug_blocks(k);
    l.n_tiles = tri(l.nb);
    l.kpad = l.nb * kTile;
    l.chunk_rows = chunk_rows;
    size_t o = 0;
    l.off_a = o;
    if (a_in_smem) o += elem * (size_t)kTileElems * l.n_tiles;
    l.off_y = o;
    o += elem * (size_t)chunk_rows * l.kpad;
    l.off_z = o;
    o += elem * (size_t)l.kpad;
    l.off_x = o;
    o += elem * (size_t)l.kpad;
    l.off_map = o;
    o += 2 * (size_t)l.n_tiles;
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
};

__device__ __forceinline__ double inv_sqrt(double x) { return rsqrt(x); }
__device__ __forceinline__ float inv_sqrt(float x) { return rsqrtf(x); }

// Cholesky of the leading nvalid x nvalid block of an 8x8 tile held in registers (lower
// triangle); the rows below it are carried along as in a panel factorisation (that is where the
// right-hand-side row picks up its part of z).  The diagonal is left INVERTED.
template <typename T>
__device__ __forceinline__ void chol8_factor(T (&a)[kTile][kTile], int nvalid) {
#pragma unroll
    for (int j = 0; j < kTile; ++j) {
        if (j < nvalid) {
            const T inv = inv_sqrt(a[j][j]);
            a[j][j] = inv;
#pragma unroll
            for (int i = j + 1; i < kTile; ++i) a[i][j] *= inv;
#pragma unroll
            for (int l = j + 1; l < kTile; ++l)
#pragma unroll
                for (int i = l; i < kTile; ++i) a[i][l] -= a[i][j] * a[l][j];
        }
    }
}

// In-place inverse of the leading lower-triangular nvalid x nvalid block (diagonal already
// inverted by chol8_factor): column j of L^-1 is -L22^-1 L[j+1:, j] / L[j][j].
template <typename T>
__device__ __forceinline__ void chol8_invert(T (&a)[kTile][kTile], int nvalid) {
#pragma unroll
    for (int j = kTile - 1; j >= 0; --j) {
        if (j < nvalid) {
            const T njj = -a[j][j];
#pragma unroll
            for (int i = kTile - 1; i > j; --i) {
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

// Factor diagonal tile `p` (tile id `t` of the tiled matrix A) in place: Cholesky of its leading
// block, then that block's inverse.  If the right-hand-side row sits in this tile (k not a
// multiple of 8 and p == k / 8) its z entries are published first.  Runs in ONE thread; kept out
// of line so that its register needs do not spill the callers' accumulator tiles.
#ifndef HYPREC_EMU
template <typename T>
__device__ __noinline__ void factor_diag_tile(T* A, int NT, int t, int p, int k, T* zvec) {
#else
template <typename T>
void factor_diag_tile(T* A, int NT, int t, int p, int k, T* zvec) {
#endif
    int nvalid = k - p * kTile;
    if (nvalid > kTile) nvalid = kTile;
    if (nvalid <= 0) return;
    T a[kTile][kTile];
#pragma unroll
    for (int r = 0; r < kTile; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) a[r][c] = A[(size_t)(r * kTile + c) * NT + t];
    chol8_factor(a, nvalid);
    if (nvalid < kTile) {
#pragma unroll
        for (int r = 1; r < kTile; ++r)
            if (r == nvalid) {
#pragma unroll
                for (int j = 0; j < r; ++j) zvec[p * kTile + j] = a[r][j];
            }
    }
    chol8_invert(a, nvalid);
#pragma unroll
    for (int r = 0; r < kTile; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) A[(size_t)(r * kTile + c) * NT + t] = a[r][c];
}

template <typename T>
__global__ void __launch_bounds__(kSolveThreads, 1) als_solve_rows_kernel(SolveArgs<T> p) {
    HYPREC_DYN_SMEM(smem_raw);
    __shared__ long long s_row;
    const int tid = threadIdx.x;
    const int lane = tid % 32, warp = tid / 32;
    const int k = p.k;
    const SolveLayout lay = solve_layout(k, p.chunk_rows, p.a_scratch == nullptr, sizeof(T));
    const int nb = lay.nb, NT = lay.n_tiles, kpad = lay.kpad, CH = lay.chunk_rows;
    T* A = p.a_scratch ? p.a_scratch + (size_t)blockIdx.x * kTileElems * NT : (T*)(smem_raw + lay.off_a);
    T* ybuf = (T*)(smem_raw + lay.off_y);
    T* zvec = (T*)(smem_raw + lay.off_z);
    T* xvec = (T*)(smem_raw + lay.off_x);
    unsigned char* tmap = smem_raw + lay.off_map;
    const int nbk = (k + kTile - 1) / kTile;   // block columns that hold factor columns
    const T pos_scale = T(kConfPos - kConfNeg);

    for (int t = tid; t < NT; t += kSolveThreads) {
        int bi, bj;
        tri_coords(t, &bi, &bj);
        tmap[2 * t] = (unsigned char)bi;
        tmap[2 * t + 1] = (unsigned char)bj;
    }

    while (true) {
        __syncthreads();   // previous row finished with A / zvec / xvec / s_row; tmap visible
        if (tid == 0) s_row = p.row_lo + (long long)atomicAdd(p.work_counter, 1);
        __syncthreads();
        const int64_t row = s_row;
        if (row >= p.row_hi) break;
        const int64_t lo = p.indptr[row];
        const int deg = (int)(p.indptr[row + 1] - lo);
        T* out = p.latent + row * k;
        if (deg == 0 && p.prior == nullptr) {
            // b == 0 => the solve returns exactly 0
            for (int c = tid; c < k; c += kSolveThreads) out[c] = T(0);
            continue;
        }

        // ---- gather + rank-1 accumulation, CH gathered rows at a time ------------------------
        const int nchunks = deg == 0 ? 1 : (deg + CH - 1) / CH;
        for (int ch = 0; ch < nchunks; ++ch) {
            const int c0 = ch * CH;
            int cr = deg - c0;
            if (cr > CH) cr = CH;
            // one warp per gathered row: coalesced read of the factor row, stored with column
            // 8*b+c at [c * nb + b] so tile owners read consecutive words
            for (int t = warp; t < cr; t += kSolveThreads / 32) {
                const int64_t src = (int64_t)p.indices[lo + c0 + t] * k;
                const T rating = p.values[lo + c0 + t];
                T* dst = ybuf + (size_t)t * kpad;
                for (int col = lane; col < kpad; col += 32) {
                    T v = T(0);
                    if (col < k) v = p.fixed[src + col];
                    else if (col == k) v = rating;
                    dst[(col % kTile) * nb + col / kTile] = v;
                }
            }
            __syncthreads();
            const bool last = (ch == nchunks - 1);
#pragma unroll 1
            for (int q = tid; q < NT; q += kSolveThreads) {
                const int bi = tmap[2 * q], bj = tmap[2 * q + 1];
                T acc[kTile][kTile];
                if (ch == 0) {
#pragma unroll
                    for (int r = 0; r < kTile; ++r)
#pragma unroll
                        for (int c = 0; c < kTile; ++c) acc[r][c] = T(0);
                } else {
#pragma unroll
                    for (int r = 0; r < kTile; ++r)
#pragma unroll
                        for (int c = 0; c < kTile; ++c) acc[r][c] = A[(size_t)(r * kTile + c) * NT + q];
                }
#pragma unroll 2
                for (int t = 0; t < cr; ++t) {
                    const T* yrow = ybuf + (size_t)t * kpad;
                    T vi[kTile], vj[kTile];
#pragma unroll
                    for (int r = 0; r < kTile; ++r) vi[r] = yrow[r * nb + bi];
#pragma unroll
                    for (int c = 0; c < kTile; ++c) vj[c] = yrow[c * nb + bj];
#pragma unroll
                    for (int r = 0; r < kTile; ++r)
#pragma unroll
                        for (int c = 0; c < kTile; ++c) acc[r][c] += vi[r] * vj[c];
                }
                if (last) {
                    // A = base + 0.9 * sum y y^T on factor rows; the rhs row keeps sum r y
#pragma unroll
                    for (int r = 0; r < kTile; ++r) {
                        const int i = bi * kTile + r;
#pragma unroll
                        for (int c = 0; c < kTile; ++c) {
                            const int j = bj * kTile + c;
                            T v;
                            if (i < k) {
                                v = p.base[(size_t)(r * kTile + c) * NT + q] + pos_scale * acc[r][c];
                            } else if (i == k) {
                                v = acc[r][c];
                                if (p.prior != nullptr && j < k) v += p.lambda * p.prior[row * k + j];
                            } else {
                                v = T(0);
                            }
                            acc[r][c] = v;
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < kTile; ++r)
#pragma unroll
                    for (int c = 0; c < kTile; ++c) A[(size_t)(r * kTile + c) * NT + q] = acc[r][c];
                if (last && q == 0) factor_diag_tile(A, NT, 0, 0, k, zvec);
            }
            __syncthreads();
        }

        // ---- right-looking blocked Cholesky, 8-column panels ---------------------------------
        for (int pnl = 0; pnl < nbk; ++pnl) {
            // panel solve: rows below the diagonal tile (factor rows and the rhs row k) times
            // the transposed inverse of the diagonal tile's factor
            const int tdiag = tile_id(pnl, pnl);
            for (int i = kTile * (pnl + 1) + tid; i <= k; i += kSolveThreads) {
                const int r = i % kTile;
                const int t = tile_id(i / kTile, pnl);
                T x[kTile], y[kTile];
#pragma unroll
                for (int c = 0; c < kTile; ++c) x[c] = A[(size_t)(r * kTile + c) * NT + t];
#pragma unroll
                for (int c = 0; c < kTile; ++c) {
                    T s = T(0);
#pragma unroll
                    for (int j = 0; j <= c; ++j) s += x[j] * A[(size_t)(c * kTile + j) * NT + tdiag];
                    y[c] = s;
                }
#pragma unroll
                for (int c = 0; c < kTile; ++c) {
                    A[(size_t)(r * kTile + c) * NT + t] = y[c];
                    ybuf[(size_t)c * kpad + r * nb + i / kTile] = y[c];   // panel, row 8*b+r at [r*nb+b]
                }
                if (i == k) {
#pragma unroll
                    for (int c = 0; c < kTile; ++c) zvec[pnl * kTile + c] = y[c];
                }
            }
            __syncthreads();
            // trailing update; the owner of the next diagonal tile factors it straight away
            const int ntr = nb - 1 - pnl;
            const int ntiles = tri(ntr);
#pragma unroll 1
            for (int q = tid; q < ntiles; q += kSolveThreads) {
                const int bi = pnl + 1 + tmap[2 * q], bj = pnl + 1 + tmap[2 * q + 1];
                const int t = tile_id(bi, bj);
                T acc[kTile][kTile];
#pragma unroll
                for (int r = 0; r < kTile; ++r)
#pragma unroll
                    for (int c = 0; c < kTile; ++c) acc[r][c] = A[(size_t)(r * kTile + c) * NT + t];
#pragma unroll 2
                for (int kk = 0; kk < kTile; ++kk) {
                    const T* prow = ybuf + (size_t)kk * kpad;
                    T li[kTile], lj[kTile];
#pragma unroll
                    for (int r = 0; r < kTile; ++r) li[r] = prow[r * nb + bi];
#pragma unroll
                    for (int c = 0; c < kTile; ++c) lj[c] = prow[c * nb + bj];
#pragma unroll
                    for (int r = 0; r < kTile; ++r)
#pragma unroll
                        for (int c = 0; c < kTile; ++c) acc[r][c] -= li[r] * lj[c];
                }
#pragma unroll
                for (int r = 0; r < kTile; ++r)
#pragma unroll
                    for (int c = 0; c < kTile; ++c) A[(size_t)(r * kTile + c) * NT + t] = acc[r][c];
                if (q == 0) factor_diag_tile(A, NT, t, pnl + 1, k, zvec);
            }
            __syncthreads();
        }

        // ---- back substitution L^T x = z, one block column at a time -------------------------
        for (int bj = nbk - 1; bj >= 0; --bj) {
            int nv = k - kTile * bj;
            if (nv > kTile) nv = kTile;
            const int td = tile_id(bj, bj);
            if (tid < nv) {
                T s = T(0);
                for (int j = tid; j < nv; ++j) s += A[(size_t)(j * kTile + tid) * NT + td] * zvec[kTile * bj + j];
                xvec[kTile * bj + tid] = s;
            }
            __syncthreads();
            for (int jp = tid; jp < kTile * bj; jp += kSolveThreads) {
                const int t = tile_id(bj, jp / kTile), c = jp % kTile;
                T s = zvec[jp];
                for (int i = 0; i < nv; ++i) s -= A[(size_t)(i * kTile + c) * NT + t] * xvec[kTile * bj + i];
                zvec[jp] = s;
            }
            __syncthreads();
        }
        for (int c = tid; c < k; c += kSolveThreads) {
            const T v = xvec[c];
            if (!(v - v == T(0))) *p.status = 1;   // inf or nan: singular system (lambda == 0)
            out[c] = v;
        }
    }
}

}  // namespace hyprec
