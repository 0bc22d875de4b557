// Gram matrix G = Y^T Y of a rows x k factor matrix, column sums, and the tiled base matrix
// 0.1 * G + lambda * I the row solver starts every system from.
//
// Replaces the dense part of `(fixed_vecs.T * confidence).dot(fixed_vecs)`
// (/root/reference/lib/collaborative_filtering.py:84, :92): every row's normal matrix is
// 0.1 * Y^T Y plus a sparse correction, so Y^T Y is computed once per half-step.
//
// This is the SIMT (float64 / float32) implementation; deterministic: per-split partial sums
// are reduced in a fixed order, no floating-point atomics.
#pragma once
#include "common.cuh"

namespace hyprec {

constexpr int kGramTile = 64;      // G tile edge per CTA
constexpr int kGramRows = 32;      // factor rows staged per step
constexpr int kGramThreads = 256;  // 16 x 16 threads, 4 x 4 interleaved outputs each

// grid = (lower-triangle tiles of G, row splits). partial[split][k][k] receives the tile's sum
// over the split's rows (only tiles with ti >= tj are written).
template <typename T>
__global__ void __launch_bounds__(kGramThreads)
gram_partial_kernel(const T* __restrict__ y, int64_t rows, int k, int64_t rows_per_split,
                    T* __restrict__ partial) {
    __shared__ T yi[kGramRows][kGramTile + 1];
    __shared__ T yj[kGramRows][kGramTile + 1];
    int ti, tj;
    tri_coords((int)blockIdx.x, &ti, &tj);
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int64_t r_begin = (int64_t)blockIdx.y * rows_per_split;
    int64_t r_end = r_begin + rows_per_split;
    if (r_end > rows) r_end = rows;

    T acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = T(0);

    for (int64_t r0 = r_begin; r0 < r_end; r0 += kGramRows) {
        // stage kGramRows x 64 columns of both column blocks (coalesced along columns)
        for (int e = tid; e < kGramRows * kGramTile; e += kGramThreads) {
            int rr = e / kGramTile, cc = e % kGramTile;
            int64_t row = r0 + rr;
            int ci = ti * kGramTile + cc, cj = tj * kGramTile + cc;
            T vi = T(0), vj = T(0);
            if (row < r_end) {
                if (ci < k) vi = y[row * k + ci];
                if (cj < k) vj = y[row * k + cj];
            }
            yi[rr][cc] = vi;
            yj[rr][cc] = vj;
        }
        __syncthreads();
#pragma unroll 4
        for (int rr = 0; rr < kGramRows; ++rr) {
            T va[4], vb[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) va[a] = yi[rr][ty + 16 * a];
#pragma unroll
            for (int b = 0; b < 4; ++b) vb[b] = yj[rr][tx + 16 * b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] += va[a] * vb[b];
        }
        __syncthreads();
    }
    T* out = partial + (int64_t)blockIdx.y * k * k;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            int i = ti * kGramTile + ty + 16 * a, j = tj * kGramTile + tx + 16 * b;
            if (i < k && j < k) out[(int64_t)i * k + j] = acc[a][b];
        }
}

// G[i][j] = sum over splits (fixed order), mirrored from the lower tile triangle.
template <typename T>
__global__ void gram_reduce_kernel(const T* __restrict__ partial, int nsplit, int k, T* __restrict__ g) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= k * k) return;
    int i = idx / k, j = idx % k;
    int si = i, sj = j;
    if (i / kGramTile < j / kGramTile) { si = j; sj = i; }
    T s = T(0);
    for (int p = 0; p < nsplit; ++p) s += partial[(int64_t)p * k * k + (int64_t)si * k + sj];
    g[idx] = s;
}

// Column sums in float64: partial[split][k], then a fixed-order reduce.  Feeds the row
// thresholds mean_i(u . v_i) = u . (sum_i v_i) / n (abstract_recommender.py:183).
template <typename T>
__global__ void colsum_partial_kernel(const T* __restrict__ y, int64_t rows, int k,
                                      int64_t rows_per_split, double* __restrict__ partial) {
    int64_t r_begin = (int64_t)blockIdx.x * rows_per_split;
    int64_t r_end = r_begin + rows_per_split;
    if (r_end > rows) r_end = rows;
    for (int c = threadIdx.x; c < k; c += blockDim.x) {
        double s = 0.0;
        for (int64_t r = r_begin; r < r_end; ++r) s += (double)y[r * k + c];
        partial[(int64_t)blockIdx.x * k + c] = s;
    }
}

__global__ void colsum_reduce_kernel(const double* __restrict__ partial, int nsplit, int k,
                                     double* __restrict__ out) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= k) return;
    double s = 0.0;
    for (int p = 0; p < nsplit; ++p) s += partial[(int64_t)p * k + c];
    out[c] = s;
}

// base[(r*8+c) * n_tiles + t] for tile t = (bi, bj): 0.1 * G + lambda on the diagonal for
// factor rows (i < k, j < k), zero for the right-hand-side row (i == k) and the padding.
template <typename T>
__global__ void build_base_tiles_kernel(const T* __restrict__ g, int k, T lambda, T* __restrict__ base) {
    const int nb = aug_blocks(k);
    const int n_tiles = tri(nb);
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= kTileElems * n_tiles) return;
    int e = idx / n_tiles, t = idx % n_tiles;
    int bi, bj;
    tri_coords(t, &bi, &bj);
    int i = bi * kTile + e / kTile, j = bj * kTile + e % kTile;
    T v = T(0);
    if (i < k && j < k) {
        v = T(kConfNeg) * g[(int64_t)i * k + j];
        if (i == j) v += lambda;
    }
    base[idx] = v;
}

// <A, B>_F of two k x k matrices, float64, single block, fixed order.  For
// ||U V^T||_F^2 = <U^T U, V^T V>_F in the RMSE.
template <typename T>
__global__ void frob_inner_kernel(const T* __restrict__ a, const T* __restrict__ b, int count,
                                  double* __restrict__ out) {
    __shared__ double red[1024];
    // four independent partial sums per thread: the loads of four strides are in flight together
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    const int step = blockDim.x;
    int i = threadIdx.x;
    for (; i + 3 * step < count; i += 4 * step) {
        s0 += (double)a[i] * (double)b[i];
        s1 += (double)a[i + step] * (double)b[i + step];
        s2 += (double)a[i + 2 * step] * (double)b[i + 2 * step];
        s3 += (double)a[i + 3 * step] * (double)b[i + 3 * step];
    }
    for (; i < count; i += step) s0 += (double)a[i] * (double)b[i];
    red[threadIdx.x] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    for (int w = blockDim.x / 2; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = red[0];
}

}  // namespace hyprec
