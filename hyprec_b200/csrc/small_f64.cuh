// Small low-rank ALS systems, arithmetic in float64: ONE WARP per row, several rows in flight per CTA.
//
// The deg x deg form of a row update (als_solve.cuh: S = I + 0.9 Q_r Q_r^T on whitened factors, rhs = r,
// y = Q_r^T t; same reference lines, /root/reference/lib/collaborative_filtering.py:83-99) for rows with at most
// 32 positives -- 95 % of the items and a third of the users at the citeulike shapes.  The tile kernel spends
// ~70 k cycles of barriers and latency on such a row with a 64-thread CTA; here a warp
//   1. builds S with FP64 tensor-core MMAs (mma.sync.m8n8k4.f64): the 8-row blocks of the gathered rows of Q
//      are A and B fragments as they lie in memory -- lane (g, c) of a block loads Q[idx[8 b + g]][4 kk + c], one
//      32-byte sector per row and k-step, no staging -- and only the lower block triangle is formed;
//   2. factors it with one row per lane (column broadcast through shared memory), solves forward with shuffles and
//      backward against the factor kept in shared memory;
//   3. forms y = Q_r^T t with lanes along the factor index (coalesced second pass over the L2-resident rows).
// No __syncthreads anywhere: warps pull rows from an atomic counter independently.  The RMSE cross term
// (|r|^2 - |z|^2) / 0.9 is a by-product, as in the other row kernels.
// T = double: the float64 mode's <= 32 bucket.  T = float: the float32 mode's <= 32 bucket -- factors and ratings
// are read and written as float32, everything in between is the same float64 arithmetic (a 32 x 32 system is too
// small for a tensor-memory pass to pay: tc_chol_multi_kernel<32> spends ~40 us per pass of 8 rows on its serial
// chain, 4.8 us per row and SM at the scaled shape's 240 k small item rows).
#pragma once
#include "common.cuh"

namespace hyprec {

constexpr int kSmallDim = 32;            // largest system (positives of the row)
constexpr int kSmallWarps = 8;           // warps (rows in flight) per CTA
constexpr int kSmallLd = kSmallDim + 1;  // row stride of the per-warp matrix (conflict-free column walks)

template <typename T>
struct SmallArgsT {
    const int64_t* indptr;     // CSR of the side being solved, indexed by global row id
    const int32_t* indices;    // ids into `fixed`
    const T* values;           // ratings
    const T* fixed;            // whitened factors Q [n_fixed][k]
    T* ymat;                   // compact output: work item w writes row ymat_row0 + w
    int64_t ymat_row0;
    const int32_t* row_list;   // rows of the launch, each with 1 .. kSmallDim positives (0 positives: y = 0)
    int64_t n_list;
    int k;
    int* work_counter;         // zeroed before launch
    int* status;               // set to 1 when a solution is not finite
    double* cross_out;         // optional [n_rows]: the row's share of the RMSE cross term
};
typedef SmallArgsT<double> SmallArgs;

HYPREC_HD size_t small_smem_bytes() {
    return (size_t)kSmallWarps * (sizeof(double) * (kSmallDim * kSmallLd + 3 * kSmallDim) + sizeof(int) * kSmallDim);
}

// D (8x8, two entries per lane) += A (8x4, one entry per lane: row lane/4, column lane%4) * B (4x8: row lane%4, column lane/4)
__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
#ifdef HYPREC_EMU
    const int lane = (int)(emu::linear_tid() % 32);
    double s0 = 0.0, s1 = 0.0;
    for (int kk = 0; kk < 4; ++kk) {
        const double av = __shfl_sync(0xffffffffu, a, (lane / 4) * 4 + kk);
        const double b0 = __shfl_sync(0xffffffffu, b, (2 * (lane % 4)) * 4 + kk);
        const double b1 = __shfl_sync(0xffffffffu, b, (2 * (lane % 4) + 1) * 4 + kk);
        s0 += av * b0;
        s1 += av * b1;
    }
    c0 += s0;
    c1 += s1;
#else
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
#endif
}

__device__ __forceinline__ double shfl_f64(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

template <typename T>
__global__ void __launch_bounds__(kSmallWarps * 32) lowrank_small_kernel(SmallArgsT<T> p) {
    HYPREC_DYN_SMEM(small_raw);
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    const int g = lane / 4, c = lane % 4;
    unsigned char* mine = small_raw + (size_t)warp * (sizeof(double) * (kSmallDim * kSmallLd + 3 * kSmallDim) + sizeof(int) * kSmallDim);
    double* S = (double*)mine;                               // [32][33]: S, then its Cholesky factor (lower)
    double* col = S + kSmallDim * kSmallLd;                  // [32] the column being eliminated
    double* tvec = col + kSmallDim;                          // [32] solution t of S t = r
    double* dinv = tvec + kSmallDim;                         // [32] 1 / L[j][j]
    int* idx = (int*)(dinv + kSmallDim);                     // [32] gathered row ids (-1: padding)
    const int k = p.k;
    const double pos = kConfPos - kConfNeg;

    while (true) {
        long long w = 0;
        if (lane == 0) w = (long long)atomicAdd(p.work_counter, 1);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= p.n_list) break;
        const int64_t row = (int64_t)p.row_list[w];
        const int64_t lo = p.indptr[row];
        const int deg = (int)(p.indptr[row + 1] - lo);
        T* out = p.ymat + (p.ymat_row0 + w) * k;
        if (deg == 0) {
            for (int a = lane; a < k; a += 32) out[a] = T(0);
            continue;
        }
        const int nb = (deg + 7) / 8;
        __syncwarp();
        idx[lane] = lane < deg ? p.indices[lo + lane] : -1;
        const double r = lane < deg ? (double)p.values[lo + lane] : 0.0;
        __syncwarp();

        // ---- 1. D = Q_r Q_r^T, lower 8 x 8 blocks, by FP64 MMAs over k-steps of 4 -------------------------
        double acc[10][2];
#pragma unroll
        for (int t = 0; t < 10; ++t) acc[t][0] = acc[t][1] = 0.0;
        const T* rowp[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int id = b < nb ? idx[8 * b + g] : -1;
            rowp[b] = id >= 0 ? p.fixed + (int64_t)id * k : nullptr;
        }
        // (warp-uniform conditions below: nb is the same on every lane)
        auto mma_blocks = [&](const double (&f)[4]) {
            dmma_8x8x4(acc[0][0], acc[0][1], f[0], f[0]);
            if (nb > 1) {
                dmma_8x8x4(acc[1][0], acc[1][1], f[1], f[0]);
                dmma_8x8x4(acc[2][0], acc[2][1], f[1], f[1]);
            }
            if (nb > 2) {
                dmma_8x8x4(acc[3][0], acc[3][1], f[2], f[0]);
                dmma_8x8x4(acc[4][0], acc[4][1], f[2], f[1]);
                dmma_8x8x4(acc[5][0], acc[5][1], f[2], f[2]);
            }
            if (nb > 3) {
                dmma_8x8x4(acc[6][0], acc[6][1], f[3], f[0]);
                dmma_8x8x4(acc[7][0], acc[7][1], f[3], f[1]);
                dmma_8x8x4(acc[8][0], acc[8][1], f[3], f[2]);
                dmma_8x8x4(acc[9][0], acc[9][1], f[3], f[3]);
            }
        };
        constexpr int kVec = 16 / (int)sizeof(T);            // factor entries per 16-byte load
        if (k % kVec == 0) {
            // One 16-byte load per lane, row block and group of kVec k-steps: lane (g, c) fetches columns
            // col .. col + kVec - 1 of row g, col = k0 + 4 kVec h + kVec c, and k-step s of the group multiplies entry s
            // of every lane's vector -- the four lanes of a row then cover columns {col(c) + s}, a different set of
            // four per step, the same one in the A and the B fragment (a sum over k does not care about its order).
            // Two groups per iteration: 8 x 16 bytes in flight per lane, a quarter (float) / half (double) of the
            // round trips of the entry-wise loop below -- the rows come from HBM at the scaled shapes.
            struct __align__(16) Vec { T v[kVec]; };
            for (int k0 = 0; k0 < k; k0 += 8 * kVec) {
                Vec ld[2][4];
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int colk = k0 + 4 * kVec * h + kVec * c;
                        if (rowp[b] != nullptr && colk < k) {
                            ld[h][b] = *reinterpret_cast<const Vec*>(rowp[b] + colk);
                        } else {
#pragma unroll
                            for (int s = 0; s < kVec; ++s) ld[h][b].v[s] = T(0);
                        }
                    }
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int s = 0; s < kVec; ++s) {
                        const double f[4] = {(double)ld[h][0].v[s], (double)ld[h][1].v[s], (double)ld[h][2].v[s], (double)ld[h][3].v[s]};
                        mma_blocks(f);
                    }
            }
        } else {
            for (int k0 = 0; k0 < k; k0 += 8) {              // two k-steps per iteration: their loads go out together
                double f[2][4];
#pragma unroll
                for (int s = 0; s < 2; ++s)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int colk = k0 + 4 * s + c;
                        f[s][b] = (rowp[b] != nullptr && colk < k) ? (double)rowp[b][colk] : 0.0;
                    }
#pragma unroll
                for (int s = 0; s < 2; ++s) mma_blocks(f[s]);
            }
        }
        // S = I + 0.9 D into shared memory: tile (bi, bj) entry (row g, columns 2c, 2c+1)
#pragma unroll
        for (int bi = 0; bi < 4; ++bi)
#pragma unroll
            for (int bj = 0; bj <= bi; ++bj) {
                if (bi < nb) {
                    const int t = bi * (bi + 1) / 2 + bj;
                    const int i = 8 * bi + g, j = 8 * bj + 2 * c;
                    S[i * kSmallLd + j] = pos * acc[t][0] + (i == j ? 1.0 : 0.0);
                    S[i * kSmallLd + j + 1] = pos * acc[t][1] + (i == j + 1 ? 1.0 : 0.0);
                }
            }
        __syncwarp();

        // ---- 2. Cholesky, one row per lane; padding rows (lane >= deg) are identity rows ------------------
        double a[kSmallDim];
#pragma unroll
        for (int j = 0; j < kSmallDim; ++j) a[j] = (lane < deg && j <= lane && j < deg) ? S[lane * kSmallLd + j] : (j == lane ? 1.0 : 0.0);
        double dmine = 1.0;
#pragma unroll
        for (int j = 0; j < kSmallDim; ++j) {
            if (j < deg) {                                  // (warp-uniform)
                const double piv = shfl_f64(a[j], j);
                const double d = rsqrt(piv);
                const double l = a[j] * d;                  // L[lane][j] for lane >= j
                if (lane == j) dmine = d;
                a[j] = l;
                col[lane] = l;
                __syncwarp();
#pragma unroll
                for (int q = j + 1; q < kSmallDim; ++q)
                    if (q < deg && q <= lane) a[q] -= l * col[q];
                __syncwarp();
            }
        }
        dinv[lane] = dmine;
#pragma unroll
        for (int j = 0; j < kSmallDim; ++j)
            if (j <= lane) S[lane * kSmallLd + j] = a[j];   // the factor, for the back substitution
        __syncwarp();

        // ---- 3. forward substitution L z = r (shuffles), then L^T t = z against the stored factor ---------
        double rhs = r, zmine = 0.0;
#pragma unroll
        for (int j = 0; j < kSmallDim; ++j) {
            if (j < deg) {
                const double zj = shfl_f64(rhs, j) * shfl_f64(dmine, j);
                if (lane == j) zmine = zj;
                if (lane > j) rhs -= a[j] * zj;
            }
        }
        if (p.cross_out != nullptr) {
            double rr = r * r, zz = zmine * zmine;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                rr += __shfl_xor_sync(0xffffffffu, rr, s);
                zz += __shfl_xor_sync(0xffffffffu, zz, s);
            }
            if (lane == 0) p.cross_out[row] = (rr - zz) / pos;
        }
        double back = zmine, tmine = 0.0;
        for (int j = deg - 1; j >= 0; --j) {
            const double tj = shfl_f64(back, j) * dinv[j];
            if (lane == j) tmine = tj;
            if (lane < j) back -= S[j * kSmallLd + lane] * tj;
        }
        tvec[lane] = lane < deg ? tmine : 0.0;
        __syncwarp();

        // ---- 4. y = Q_r^T t: lanes along the factor index -----------------------------------------------------
        bool finite = true;
        if (k % kVec == 0) {
            // 16-byte loads, eight rows in flight per lane: the rows are L2-resident after step 1, so a row costs one
            // round trip per 32 kVec columns and eight of them overlap
            struct __align__(16) Vec { T v[kVec]; };
            for (int col0 = kVec * lane; col0 < k; col0 += 32 * kVec) {
                double s[kVec];
#pragma unroll
                for (int e = 0; e < kVec; ++e) s[e] = 0.0;
                for (int i0 = 0; i0 < deg; i0 += 8) {
                    Vec q[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = i0 + u < deg ? i0 + u : deg - 1;       // (the repeats get weight 0 below)
                        q[u] = *reinterpret_cast<const Vec*>(p.fixed + (int64_t)idx[i] * k + col0);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const double w = i0 + u < deg ? tvec[i0 + u] : 0.0;
#pragma unroll
                        for (int e = 0; e < kVec; ++e) s[e] += (double)q[u].v[e] * w;
                    }
                }
                Vec o;
#pragma unroll
                for (int e = 0; e < kVec; ++e) {
                    finite = finite && (s[e] - s[e] == 0.0);
                    o.v[e] = (T)s[e];
                }
                *reinterpret_cast<Vec*>(out + col0) = o;
            }
        } else {
            for (int col0 = lane; col0 < k; col0 += 32) {
                double s = 0.0;
                int i = 0;
                for (; i + 4 <= deg; i += 4) {
                    const double q0 = (double)p.fixed[(int64_t)idx[i] * k + col0], q1 = (double)p.fixed[(int64_t)idx[i + 1] * k + col0];
                    const double q2 = (double)p.fixed[(int64_t)idx[i + 2] * k + col0], q3 = (double)p.fixed[(int64_t)idx[i + 3] * k + col0];
                    s += q0 * tvec[i];
                    s += q1 * tvec[i + 1];
                    s += q2 * tvec[i + 2];
                    s += q3 * tvec[i + 3];
                }
                for (; i < deg; ++i) s += (double)p.fixed[(int64_t)idx[i] * k + col0] * tvec[i];
                finite = finite && (s - s == 0.0);
                out[col0] = (T)s;
            }
        }
        if (!finite) *p.status = 1;
    }
}

}  // namespace hyprec
