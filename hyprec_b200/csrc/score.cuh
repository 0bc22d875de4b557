// Dense score sweep S = U V^T in 128 x 128 tiles with a fused epilogue, and the small
// per-row kernels around it (row thresholds, sparse dot products, rounded flags).
//
// Replaces (all without keeping the users x items matrix):
//   predictions = user_vecs.dot(item_vecs.T)              collaborative_filtering.py:270
//   avg = sum(predictions[user]) / n; rounded = score>=avg abstract_recommender.py:183-186
//   sum(sum(rounded_predictions))  (the `ratio` numerator) abstract_recommender.py:115
//   rounded[ratings.nonzero()]     (calculate_recall)      evaluator.py:264-266
//   sum_nz r * (u . v)             (RMSE cross term)       evaluator.py:249 via the Gram identity
// SIMT float64 / float32 implementation.
#pragma once
#include "common.cuh"

namespace hyprec {

constexpr int kSweepTile = 128;
constexpr int kSweepKC = 16;
constexpr int kSweepThreads = 256;
constexpr int kSweepStore = 0;   // write the tile to a dense float64 matrix
constexpr int kSweepCount = 1;   // count score >= threshold per row
constexpr int kSweepStorePeers = 2;   // kSweepStore + the same stores into the replicas on the other GPUs

template <typename T>
struct SweepArgs {
    const T* user_vecs;   // [n_users][k]
    const T* item_vecs;   // [n_items][k]
    int k;
    int64_t user_lo, user_hi, n_items;
    double* dense_out;               // [user_hi-user_lo][n_items]          (kSweepStore)
    float* dense_out_f32;            // (kSweepStore) when non-null the tile is stored here as float32 instead
    const int32_t* row_list;         // optional (kSweepStore): row r of the product is stored at
                                     // dense_out[row_list[r] * n_items] instead of row r - user_lo
    const double* thresholds;        // [n_users]                           (kSweepCount)
    unsigned long long* counts;      // [user_hi-user_lo], zeroed by caller (kSweepCount)
    // Multi-GPU (comm.cuh, kSweepStore): the tile is also stored into these replicas of the output on
    // the other GPUs (NVLink peer memory); same element type as the local output.
    int n_peer = 0;
    void* peer_out[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

// grid = (item tiles, user tiles); 16 x 16 threads, each an 8 x 8 set of outputs interleaved
// by 16 so that shared-memory reads are conflict-free / broadcast.
template <typename T, int MODE>
__global__ void __launch_bounds__(kSweepThreads) score_sweep_kernel(SweepArgs<T> p) {
    __shared__ T us[kSweepKC][kSweepTile + 1];
    __shared__ T vs[kSweepKC][kSweepTile + 1];
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int64_t i0 = p.user_lo + (int64_t)blockIdx.y * kSweepTile;
    const int64_t j0 = (int64_t)blockIdx.x * kSweepTile;
    const int k = p.k;

    T acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = T(0);

    for (int c0 = 0; c0 < k; c0 += kSweepKC) {
        // all of the step's global loads first (16 in flight per thread), then the shared-memory stores: left to
        // the compiler the loop has come out as load -> store pairs, i.e. one exposed latency per element
        constexpr int kStage = kSweepTile * kSweepKC / kSweepThreads;
        T ureg[kStage], vreg[kStage];
#pragma unroll
        for (int i = 0; i < kStage; ++i) {
            const int e = tid + i * kSweepThreads;
            const int rr = e / kSweepKC, cc = e % kSweepKC;
            const int col = c0 + cc;
            const int64_t ui = i0 + rr, vj = j0 + rr;
            ureg[i] = (ui < p.user_hi && col < k) ? p.user_vecs[ui * k + col] : T(0);
            vreg[i] = (vj < p.n_items && col < k) ? p.item_vecs[vj * k + col] : T(0);
        }
#pragma unroll
        for (int i = 0; i < kStage; ++i) {
            const int e = tid + i * kSweepThreads;
            const int rr = e / kSweepKC, cc = e % kSweepKC;
            us[cc][rr] = ureg[i];
            vs[cc][rr] = vreg[i];
        }
        __syncthreads();
#pragma unroll
        for (int cc = 0; cc < kSweepKC; ++cc) {
            T va[8], vb[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) va[a] = us[cc][ty + 16 * a];
#pragma unroll
            for (int b = 0; b < 8; ++b) vb[b] = vs[cc][tx + 16 * b];
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) acc[a][b] += va[a] * vb[b];
        }
        __syncthreads();
    }

#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int64_t row = i0 + ty + 16 * a;
        const bool row_ok = row < p.user_hi;
        if (MODE == kSweepStore || MODE == kSweepStorePeers) {
            if (row_ok) {
                const int64_t dst = p.row_list ? (int64_t)p.row_list[row] : row - p.user_lo;
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const int64_t col = j0 + tx + 16 * b;
                    if (col < p.n_items) {
                        if (p.dense_out_f32) p.dense_out_f32[dst * p.n_items + col] = (float)acc[a][b];
                        else p.dense_out[dst * p.n_items + col] = (double)acc[a][b];
                        if (MODE == kSweepStorePeers) {
                            for (int rep = 0; rep < p.n_peer; ++rep) {
                                if (p.dense_out_f32) static_cast<float*>(p.peer_out[rep])[dst * p.n_items + col] = (float)acc[a][b];
                                else static_cast<double*>(p.peer_out[rep])[dst * p.n_items + col] = (double)acc[a][b];
                            }
                        }
                    }
                }
            }
        } else {
            const double thr = row_ok ? p.thresholds[row] : 0.0;
            unsigned cnt = 0;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const int64_t col = j0 + tx + 16 * b;
                if (row_ok && col < p.n_items && (double)acc[a][b] >= thr) ++cnt;
            }
            // the 16 lanes that share ty hold the same row
            cnt += __shfl_xor_sync(0xffffffffu, cnt, 8);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, 4);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
            if (tx == 0 && row_ok && cnt) atomicAdd(p.counts + (row - p.user_lo), (unsigned long long)cnt);
        }
    }
}

template <typename TA, typename TB>
__device__ __forceinline__ double warp_dot(const TA* __restrict__ a, const TB* __restrict__ b, int k, int lane) {
    double s = 0.0;
    for (int l = lane; l < k; l += 32) s += (double)a[l] * (double)b[l];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    return s;
}

// thresholds[u] = u . colsum / n_items, one warp per user (float64).
template <typename T>
__global__ void row_threshold_kernel(const T* __restrict__ user_vecs, const double* __restrict__ colsum,
                                     int k, int64_t n_users, int64_t n_items, double* __restrict__ thresholds) {
    const int lane = threadIdx.x % 32;
    const int64_t u = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    if (u >= n_users) return;   // whole warps leave together
    const double s = warp_dot(user_vecs + u * k, colsum, k, lane);
    if (lane == 0) thresholds[u] = s / (double)n_items;
}

// Per CSR row: s1 = sum_nz r * (u . v_j), s2 = sum_nz r^2 (float64), one warp per row.
// s1 = sum_l u_l * (sum_e r_e v_{e,l}): every lane keeps its slice of u in registers and accumulates
// its slice of the r-weighted sum of the row's item vectors (in the factor dtype: float32 mode stays on
// the float32 pipe) -- no shuffles inside the loop, four gathered rows in flight -- and the warp
// reduces once per row in float64 (fixed order: deterministic).
template <typename T>
__global__ void sddmm_row_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                 const T* __restrict__ values, const T* __restrict__ user_vecs,
                                 const T* __restrict__ item_vecs, int k, int64_t n_rows,
                                 double* __restrict__ s1, double* __restrict__ s2) {
    constexpr int kSlice = 16;   // register slice per lane: k <= 512
    const int lane = threadIdx.x % 32;
    const int64_t u = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    if (u >= n_rows) return;
    double a1 = 0.0, a2 = 0.0;
    const int64_t e0 = indptr[u], e1 = indptr[u + 1];
    if (k <= 32 * kSlice) {
        T acc[kSlice];
#pragma unroll
        for (int i = 0; i < kSlice; ++i) acc[i] = T(0);
        int64_t e = e0;
        for (; e + 4 <= e1; e += 4) {
            const T* v0 = item_vecs + (int64_t)indices[e] * k;
            const T* v1 = item_vecs + (int64_t)indices[e + 1] * k;
            const T* v2 = item_vecs + (int64_t)indices[e + 2] * k;
            const T* v3 = item_vecs + (int64_t)indices[e + 3] * k;
            const T r0 = values[e], r1 = values[e + 1], r2 = values[e + 2], r3 = values[e + 3];
            a2 += (double)r0 * (double)r0 + (double)r1 * (double)r1 + (double)r2 * (double)r2 + (double)r3 * (double)r3;
#pragma unroll
            for (int i = 0; i < kSlice; ++i) {
                const int l = lane + 32 * i;
                if (l < k) acc[i] += (r0 * v0[l] + r1 * v1[l]) + (r2 * v2[l] + r3 * v3[l]);
            }
        }
        for (; e < e1; ++e) {
            const T* v0 = item_vecs + (int64_t)indices[e] * k;
            const T r0 = values[e];
            a2 += (double)r0 * (double)r0;
#pragma unroll
            for (int i = 0; i < kSlice; ++i) {
                const int l = lane + 32 * i;
                if (l < k) acc[i] += r0 * v0[l];
            }
        }
#pragma unroll
        for (int i = 0; i < kSlice; ++i) {
            const int l = lane + 32 * i;
            if (l < k) a1 += (double)acc[i] * (double)user_vecs[u * k + l];
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) a1 += __shfl_xor_sync(0xffffffffu, a1, d);
    } else {
        for (int64_t e = e0; e < e1; ++e) {
            const double r = (double)values[e];
            const double d = warp_dot(user_vecs + u * k, item_vecs + (int64_t)indices[e] * k, k, lane);
            a1 += r * d;
            a2 += r * r;
        }
    }
    if (lane == 0) {
        s1[u] = a1;
        s2[u] = a2;
    }
}

// partial[blockIdx.x] = sum of squares of a slice of x (float64, fixed order)
template <typename T>
__global__ void sumsq_partial_kernel(const T* __restrict__ x, int64_t count, double* __restrict__ partial) {
    __shared__ double red[256];
    const int64_t per = (count + gridDim.x - 1) / gridDim.x;
    const int64_t lo = per * blockIdx.x, hi = lo + per < count ? lo + per : count;
    double s = 0.0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) s += (double)x[i] * (double)x[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = blockDim.x / 2; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

// out[0] = sum in[0..n) in a fixed order (single block).
__global__ void reduce_sum_kernel(const double* __restrict__ in, int64_t n, double* __restrict__ out) {
    __shared__ double red[1024];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += in[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = blockDim.x / 2; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = red[0];
}

// flags[e - indptr[row_lo]] = (u . v_j >= thresholds[u]) for every non-zero e of rows
// [row_lo, row_hi), one warp per positive.
template <typename T>
__global__ void nz_flags_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                const T* __restrict__ user_vecs, const T* __restrict__ item_vecs, int k,
                                const double* __restrict__ thresholds, int64_t row_lo, int64_t row_hi,
                                uint8_t* __restrict__ flags) {
    // one warp per positive (rows differ in length by orders of magnitude); its row by bisection
    const int lane = threadIdx.x % 32;
    const int64_t base = indptr[row_lo];
    const int64_t e = base + (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    if (e >= indptr[row_hi]) return;
    int64_t lo = row_lo, hi = row_hi;          // invariant: indptr[lo] <= e < indptr[hi]
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) / 2;
        if (indptr[mid] <= e) lo = mid;
        else hi = mid;
    }
    const int64_t u = lo;
    const double d = warp_dot(user_vecs + u * k, item_vecs + (int64_t)indices[e] * k, k, lane);
    if (lane == 0) flags[e - base] = d >= thresholds[u] ? 1 : 0;
}

}  // namespace hyprec
