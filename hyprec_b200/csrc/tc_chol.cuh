// Fused per-row ALS update with the row's normal matrix resident in tensor memory (float32 mode,
// systems of up to 256 unknowns, sm_100a).  Same mathematics and the same reference lines as
// als_solve.cuh (/root/reference/lib/collaborative_filtering.py:83-99: confidence-weighted normal
// matrix, right-hand side, numpy.linalg.solve), different machine mapping:
//
//   1. the rank-k accumulation (tc_syrk.cuh) leaves D = 0.9 X X^T in TMEM and never copies it out;
//   2. blocked right-looking Cholesky with 32-column panels.  One thread owns one row of the system:
//      it pulls its 32 panel entries from TMEM (tcgen05.ld, lane = row), the warp that owns the 32
//      diagonal rows factors the 32 x 32 block in registers (shuffle broadcasts), every row below solves
//      its 32 entries against it, writes them back to TMEM (tcgen05.st: the finished columns of D become
//      L) and, split hi/lo, into the operand stage, and ONE tcgen05.mma chain applies the trailing update
//      D22 -= L21 L21^T to the whole remaining matrix (3xTF32, negated A operand);
//   3. the right-hand side rides along (z = L^-1 b is finished panel by panel), and the back
//      substitution re-reads L from TMEM, reducing each panel's 32 dot products with a
//      transpose-reduce butterfly (31 shuffles per warp).
//
// The trailing update -- the O(n^3) part -- runs on the tensor cores; the SIMT cores only see
// O(n^2) work per system plus the 32 x 32 diagonal blocks.
//
// System forms (see als_solve.cuh for the derivation):
//   low-rank  S = I + 0.9 Q_r Q_r^T (deg x deg),  rhs = r - 0.9 Q_r p,  output y = Q_r^T t + p
//   direct    A = 0.1 G + lambda I + 0.9 Y_r^T Y_r (k x k),  rhs = sum r y + lambda theta,  output x
// Padding up to the next multiple of 32 is an identity block, so the factorisation never sees it.
#pragma once
#include "tc_syrk.cuh"

namespace hyprec {
namespace tc {

constexpr int kCholThreads = 256;   // one thread per system row
constexpr int kCholPanel = 32;

struct CholRowsArgs {
    const int64_t* indptr;    // CSR of the side being solved, indexed by global row id
    const int32_t* indices;   // ids into `fixed`
    const float* values;      // ratings
    const float* fixed;       // low-rank: whitened factors Q [n_fixed][k]; direct: factors Y
    const float* gram;        // direct: dense k x k Gram of the fixed side (0.1 G + lambda I is formed on the fly)
    const float* prior;       // low-rank: rows of theta L^-T; direct: theta; or null
    float* out;               // low-rank: ymat, compact (work item w writes row out_row0 + w); direct: latent [n_rows][k]
    int64_t out_row0;
    int* work_counter;        // zeroed before launch
    int* status;              // set to 1 when a solution is not finite
    float lambda;
    int k;
    int64_t row_lo, row_hi;
    const int32_t* row_list;  // optional row ids instead of [row_lo, row_hi)
    int64_t n_list;
    int max_dim;              // largest system of the launch (<= 256)
    int lowrank;
    int tmem_cols;
    double* cross_out;        // optional [n_rows]: sum_j r_j (fixed_j . x) of the row -- the RMSE cross term, a by-product
                              // of the solve (b.x = |L^-1 b|^2; low-rank form: (|r|^2 - |z|^2) / 0.9); no prior
    long long* prof;          // optional clock64() phase totals of block 0
    // Direct rows whose positives are split over several CTAs (tc_syrk.cuh split_partial_kernel): work items
    // w < n_split start their own accumulation at positive split_lo[w] (absolute offset into `indices`) and add
    // what the partial kernel accumulated over the positives before it:
    //   split_gram[w] = G + 9 sum_rest y y^T  (dense k x k, takes the place of `gram`: 0.1 * it = 0.1 G + 0.9 sum_rest)
    //   split_b[w]    = sum_rest r y
    int n_split = 0;
    const int64_t* split_lo = nullptr;
    const float* split_gram = nullptr;
    const float* split_b = nullptr;
};

struct CholLayout {
    size_t off_stage, off_l11, off_l11t, off_dinv, off_z, off_x, off_b, off_part, off_idx, off_rst, off_red, off_bar, total;
};

__host__ __device__ __forceinline__ CholLayout chol_layout(int max_dim) {
    CholLayout l;
    size_t o = 0;
    l.off_stage = o;                       // 2 stages x (hi + lo) x rows_pad x 128 B, 1024-byte aligned
    o += 2 * syrk_stage_bytes(max_dim);
    l.off_l11 = o;  o += sizeof(float) * 32 * 33;   // L11 row-major, padded (column reads)
    l.off_l11t = o; o += sizeof(float) * 32 * 32;   // L11 transposed (broadcast vector reads)
    l.off_dinv = o; o += sizeof(float) * 256;       // 1 / L[j][j]
    l.off_z = o;    o += sizeof(float) * 256;       // z = L^-1 b
    l.off_x = o;    o += sizeof(float) * 256;       // solution
    l.off_b = o;    o += sizeof(float) * 256;       // right-hand side
    l.off_part = o; o += sizeof(float) * 8 * 32;    // per-warp partial dot products (back substitution)
    l.off_idx = o;  o += sizeof(int) * 256;         // positives' ids
    l.off_rst = o;  o += sizeof(float) * 256;       // ratings of the current batch (direct gather)
    l.off_red = o;  o += sizeof(float) * 16;        // per-warp partial sums of |r|^2 and |z|^2
    l.off_bar = (o + 7) & ~(size_t)7;
    o = l.off_bar + 32;
    l.total = o + 1024;                    // alignment slack
    return l;
}

__device__ __forceinline__ void tmem_store_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
          "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
          "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// D[rows of the M-block][n_cols columns from TMEM column `tcol`] -= A B^T, K = 32 (one staged panel):
// A = stage rows [a_row0, a_row0+128), B = stage rows [b_row0, b_row0+n_cols); 3xTF32, A negated.
__device__ __forceinline__ void chol_issue_update(uint32_t tmem_d, unsigned char* hi_base, size_t half_bytes, int a_row0,
                                                  int b_row0, int n_cols) {
    const uint32_t sa = smem_u32(hi_base);
    const uint64_t a_hi = umma_desc(sa + (uint32_t)a_row0 * 128), a_lo = umma_desc(sa + (uint32_t)half_bytes + (uint32_t)a_row0 * 128);
    const uint64_t b_hi = umma_desc(sa + (uint32_t)b_row0 * 128), b_lo = umma_desc(sa + (uint32_t)half_bytes + (uint32_t)b_row0 * 128);
    const uint32_t idesc = umma_idesc_tf32(128, n_cols) | (1u << 13);   // a_negate
#pragma unroll
    for (int ks = 0; ks < kCholPanel / 8; ++ks) {
        const uint64_t adv = (uint64_t)((ks * 8 * 4) >> 4);
        umma_tf32(tmem_d, a_hi + adv, b_hi + adv, idesc, 1);
        umma_tf32(tmem_d, a_hi + adv, b_lo + adv, idesc, 1);
        umma_tf32(tmem_d, a_lo + adv, b_hi + adv, idesc, 1);
    }
}


// ---- building blocks shared by the single-system kernel below and tc_chol_multi.cuh ---------------

// Cholesky of a 32 x 32 diagonal block held one row per lane (a[j] = entry (lane, j)); on return a
// holds L11 (row `lane`), column j of L11 sits at l11t[j * 32 + lane] and the return value is
// 1 / L[lane][lane].  Column j (still unscaled) goes through shared memory: every lane reads the
// entries of the rows after j as broadcast vector loads; the next pivot only needs lane j+1's own
// entries, so it goes out first with one shuffle on the critical path.  colbuf: 2 x 32 floats.
__device__ __forceinline__ float chol_factor_diag(float (&a)[32], int lane, float* colbuf, float* l11t) {
    float dmine = 0.f;
    float piv = __shfl_sync(0xffffffffu, a[0], 0);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        float* col = colbuf + (j & 1) * 32;
        col[lane] = a[j];
        const float d = rsqrt_approx(piv);
        const float w = a[j] * d * d;      // a[j] / pivot
        if (lane == j) dmine = d;
        if (j < 31) piv = __shfl_sync(0xffffffffu, fmaf(-w, a[j], a[j + 1]), j + 1);
        __syncwarp();
        const float4* col4 = reinterpret_cast<const float4*>(col);
#pragma unroll
        for (int q = (j + 1) / 4; q < 8; ++q) {
            const float4 t = col4[q];
            if (4 * q > j) a[4 * q] = fmaf(-w, t.x, a[4 * q]);
            if (4 * q + 1 > j) a[4 * q + 1] = fmaf(-w, t.y, a[4 * q + 1]);
            if (4 * q + 2 > j) a[4 * q + 2] = fmaf(-w, t.z, a[4 * q + 2]);
            if (4 * q + 3 > j) a[4 * q + 3] = fmaf(-w, t.w, a[4 * q + 3]);
        }
        a[j] *= d;
        l11t[j * 32 + lane] = a[j];        // column j of L11, for the rows below
    }
    return dmine;
}

// x = a L11^-T for one row below the diagonal block (its 32 entries of L), column by column.
// dinv32: the 32 reciprocal diagonal entries of the block (16-byte aligned).
__device__ __forceinline__ void chol_panel_solve(float (&a)[32], const float* l11t, const float* dinv32) {
    const float4* l4 = reinterpret_cast<const float4*>(l11t);
    const float4* dinv4 = reinterpret_cast<const float4*>(dinv32);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const float4 dv = dinv4[j / 4];
        a[j] *= (j % 4 == 0) ? dv.x : (j % 4 == 1) ? dv.y : (j % 4 == 2) ? dv.z : dv.w;
#pragma unroll
        for (int q = (j + 1) / 4; q < 8; ++q) {
            const float4 t = l4[j * 8 + q];
            if (4 * q > j) a[4 * q] = fmaf(-a[j], t.x, a[4 * q]);
            if (4 * q + 1 > j) a[4 * q + 1] = fmaf(-a[j], t.y, a[4 * q + 1]);
            if (4 * q + 2 > j) a[4 * q + 2] = fmaf(-a[j], t.z, a[4 * q + 2]);
            if (4 * q + 3 > j) a[4 * q + 3] = fmaf(-a[j], t.w, a[4 * q + 3]);
        }
    }
}

// the 32 entries of one row as the hi / lo operand row `r` of a K = 32 panel stage
__device__ __forceinline__ void chol_stage_row(const float (&a)[32], unsigned char* hi_base, size_t half_bytes, int r) {
    const size_t row_off = (size_t)(r >> 3) * 1024 + (size_t)(r & 7) * 128;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float4 h, l;
        split_tf32(a[4 * c], h.x, l.x);
        split_tf32(a[4 * c + 1], h.y, l.y);
        split_tf32(a[4 * c + 2], h.z, l.z);
        split_tf32(a[4 * c + 3], h.w, l.w);
        const size_t off = row_off + (size_t)((c ^ (r & 7)) << 4);
        *reinterpret_cast<float4*>(hi_base + off) = h;
        *reinterpret_cast<float4*>(hi_base + half_bytes + off) = l;
    }
    fence_async_smem();
}

__device__ __forceinline__ void tmem_store_row(uint32_t taddr, const float (&a)[32]) {
    uint32_t raw[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) raw[j] = __float_as_uint(a[j]);
    tmem_store_32x32(taddr, raw);
}

// forward substitution of one panel's right-hand side by one warp: returns z[lane] of L11 z = b
__device__ __forceinline__ float chol_forward_rhs(float bv, int lane, const float* l11t, const float* dinv32) {
    float zmine = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const float zj = __shfl_sync(0xffffffffu, bv, j) * dinv32[j];
        if (lane == j) zmine = zj;
        if (lane > j) bv = fmaf(-l11t[j * 32 + lane], zj, bv);
    }
    return zmine;
}

// sum over the warp's 32 rows of v[j] for every j: afterwards lane l holds the total of entry l
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float keep = up ? v[i + s] : v[i];
            const float send = up ? v[i] : v[i + s];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

// L11^T x = rhs by one warp (l11: row-major, row stride 33): returns x[lane]
__device__ __forceinline__ float chol_backward_diag(float rhs, int lane, const float* l11, const float* dinv32) {
    float xmine = 0.f;
#pragma unroll
    for (int j = 31; j >= 0; --j) {
        const float xj = __shfl_sync(0xffffffffu, rhs, j) * dinv32[j];
        if (lane == j) xmine = xj;
        if (lane < j) rhs = fmaf(-l11[j * 33 + lane], xj, rhs);
    }
    return xmine;
}

#define HYPREC_TCPROF(slot)                                                          \
    if (p.prof != nullptr && blockIdx.x == 0 && tid == 0) {                          \
        const long long now_ = clock64();                                            \
        p.prof[slot] += now_ - prof_t;                                               \
        prof_t = now_;                                                               \
    }

__global__ void __launch_bounds__(kCholThreads, 1) tc_chol_rows_kernel(CholRowsArgs p) {
    extern __shared__ __align__(16) unsigned char chol_smem_dyn[];
    unsigned char* smem = chol_smem_dyn + ((1024u - (smem_u32(chol_smem_dyn) & 1023u)) & 1023u);
    const CholLayout lay = chol_layout(p.max_dim);
    float* l11 = (float*)(smem + lay.off_l11);
    float* l11t = (float*)(smem + lay.off_l11t);
    float* dinv = (float*)(smem + lay.off_dinv);
    float* zs = (float*)(smem + lay.off_z);
    float* xs = (float*)(smem + lay.off_x);
    float* bs = (float*)(smem + lay.off_b);
    float* part = (float*)(smem + lay.off_part);
    int* idxs = (int*)(smem + lay.off_idx);
    float* rst = (float*)(smem + lay.off_rst);
    float* red = (float*)(smem + lay.off_red);
    uint64_t* bars = (uint64_t*)(smem + lay.off_bar);       // [0], [1] staging stages, [2] panel updates
    uint32_t* tmem_slot = (uint32_t*)(bars + 3);
    __shared__ long long s_row;

    const int tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
    const int k = p.k;
    const bool lowrank = p.lowrank != 0;
    const float sqrt_pos = 0.9486832980505138f;             // sqrt(1.0 - 0.1): D accumulates 0.9 X X^T
    const float pos_scale = (float)(kConfPos - kConfNeg);
    const int64_t n_work = p.row_list ? p.n_list : p.row_hi - p.row_lo;
    long long prof_t = 0;

    SyrkState st;
    st.bars = bars;
    st.phase[0] = st.phase[1] = 0;
    st.stage = smem + lay.off_stage;
    st.rows_pad = syrk_rows_pad(p.max_dim);
    uint32_t upd_phase = 0;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    st.tmem_base = *tmem_slot;
    const size_t half_bytes = (size_t)st.rows_pad * 128;
    // this thread's row r = tid lives in TMEM lane tid % 128 of block tid / 128
    const uint32_t my_taddr = st.tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(tid >= 128 ? 128 : 0);

    while (true) {
        __syncthreads();
        if (tid == 0) s_row = (long long)atomicAdd(p.work_counter, 1);
        __syncthreads();
        if (s_row >= n_work) break;
        const int64_t row = p.row_list ? (int64_t)p.row_list[s_row] : p.row_lo + s_row;
        if (p.prof != nullptr && blockIdx.x == 0 && tid == 0) prof_t = clock64();
        const bool split = s_row < p.n_split;          // (direct mode only: the host plans it)
        const int64_t lo = split ? p.split_lo[s_row] : p.indptr[row];
        const int deg = (int)(p.indptr[row + 1] - lo);
        const float* gram_mat = split ? p.split_gram + (size_t)s_row * k * k : p.gram;
        float* out = p.out + (lowrank ? p.out_row0 + s_row : row) * k;
        const float* prior_row = p.prior ? p.prior + row * k : nullptr;
        if (deg == 0 && (lowrank || prior_row == nullptr)) {
            // direct: b == 0 => the solve returns exactly 0; low-rank: y = p
            for (int c = tid; c < k; c += kCholThreads) out[c] = (lowrank && prior_row) ? p.lambda * prior_row[c] : 0.f;
            continue;
        }
        const int dim = lowrank ? deg : k;
        const int dimp = (dim + kCholPanel - 1) / kCholPanel * kCholPanel;
        const int n_panels = dimp / kCholPanel;
        const bool two = dimp > 128;

        // ---- 1. D = 0.9 X X^T into TMEM; right-hand side into bs ------------------------------
        if (lowrank) {
            for (int i = tid; i < dim; i += kCholThreads) idxs[i] = p.indices[lo + i];
            __syncthreads();
            if (prior_row == nullptr) {
                for (int i = tid; i < dim; i += kCholThreads) bs[i] = p.values[lo + i];
            } else {
                for (int i = warp; i < dim; i += kCholThreads / 32) {
                    const float* q = p.fixed + (int64_t)idxs[i] * k;
                    double d = 0.0;
                    for (int l = lane; l < k; l += 32) d += (double)q[l] * (double)prior_row[l];
#pragma unroll
                    for (int s = 16; s > 0; s >>= 1) d += __shfl_xor_sync(0xffffffffu, d, s);
                    if (lane == 0) bs[i] = p.values[lo + i] - pos_scale * p.lambda * (float)d;
                }
            }
            __syncthreads();
            syrk_lowrank_accumulate(p.fixed, idxs, dim, k, st, sqrt_pos, dimp);
        } else {
            if (deg > 0) {
                syrk_direct_accumulate(p.fixed, p.indices, p.values, lo, deg, k, st, sqrt_pos, dimp, idxs, rst, 256, bs);
            } else {
                // no positives (only reached with a prior): D = 0 through one k-block of zero operands
                if (tid < k) bs[tid] = 0.f;
                float4* z4 = reinterpret_cast<float4*>(st.stage);
                for (size_t e = tid; e < 2 * half_bytes / sizeof(float4); e += kCholThreads) z4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
                fence_async_smem();
                __syncthreads();
                if (tid == 0) syrk_issue(st, st.stage, half_bytes, dimp, two, false, &bars[0]);
                mbar_wait(&bars[0], st.phase[0]);
                st.phase[0] ^= 1;
                tc_fence_after();
            }
            if (prior_row != nullptr && tid < k) bs[tid] += p.lambda * prior_row[tid];
            if (split && tid < k) bs[tid] += p.split_b[(size_t)s_row * k + tid];
        }
        HYPREC_TCPROF(0)
        if (tid >= dim && tid < dimp) bs[tid] = 0.f;   // identity padding solves to zero
        if (p.cross_out != nullptr && lowrank) {       // |r|^2 while bs still holds the ratings
            float q = tid < dim ? bs[tid] * bs[tid] : 0.f;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) q += __shfl_xor_sync(0xffffffffu, q, s);
            if (lane == 0) red[8 + warp] = q;
        }
        // direct mode: this row's 32 entries of the shared Gram for the coming panel, fetched one panel
        // ahead (behind the trailing update's wait)
        float4 g4[8];
        auto fetch_gram = [&](int c0) {
            if (!lowrank && tid < dim && tid >= c0) {
                const float* grow = gram_mat + (int64_t)tid * k + c0;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    g4[c] = (c0 + 4 * c < dim) ? __ldg(reinterpret_cast<const float4*>(grow + 4 * c)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        fetch_gram(0);

        // ---- 2. blocked Cholesky, 32-column panels --------------------------------------------
        for (int pn = 0; pn < n_panels; ++pn) {
            const int c0 = pn * kCholPanel;
            const int r = tid;
            float a[32];
            if (warp >= pn && r < dimp) {              // (whole warps: dimp is a multiple of 32)
                uint32_t raw[32];
                tmem_load_32x32(my_taddr + (uint32_t)c0, raw);
                const float* gv = reinterpret_cast<const float*>(g4);
                const int dj = r - c0;                 // this row's diagonal entry sits at a[dj] when 0 <= dj < 32
                if (r < dim && c0 + kCholPanel <= dim) {
                    // interior panel (the common case): branch-free
                    const float diag_add = lowrank ? 1.f : p.lambda;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float v = __uint_as_float(raw[j]);
                        if (!lowrank) v = fmaf((float)kConfNeg, gv[j], v);
                        a[j] = v + ((j == dj) ? diag_add : 0.f);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int col = c0 + j;
                        float v;
                        if (r < dim && col < dim) {
                            v = __uint_as_float(raw[j]);
                            if (lowrank) v += (j == dj) ? 1.f : 0.f;
                            else v += (float)kConfNeg * gv[j] + ((j == dj) ? p.lambda : 0.f);
                        } else {
                            v = (j == dj) ? 1.f : 0.f;     // identity padding
                        }
                        a[j] = v;
                    }
                }
            }
            if (warp == pn) {
                // the 32 x 32 diagonal block, one row per lane (part: free until the back substitution)
                dinv[c0 + lane] = chol_factor_diag(a, lane, part, l11t);
            }
            __syncthreads();
            HYPREC_TCPROF(2)
            const int rem0 = c0 + kCholPanel;
            unsigned char* hi_base = st.stage;         // panel updates use stage 0
            if (warp == pn) {
                // (off the other warps' critical path) L11 goes back to TMEM for the back substitution
                tmem_store_row(my_taddr + (uint32_t)c0, a);
            }
            if (warp == 0) {
                // forward substitution of this panel's right-hand side, z_p = L11^-1 b_p, by warp 0: it is
                // the diagonal warp of panel 0 and has no rows left in the later panels, so the chain runs
                // beside the other warps' panel solve
                zs[c0 + lane] = chol_forward_rhs(bs[c0 + lane], lane, l11t, dinv + c0);
            }
            if (warp > pn && r < dimp) {
                // x = a L11^-T (this row's 32 entries of L), column by column
                chol_panel_solve(a, l11t, dinv + c0);
                tmem_store_row(my_taddr + (uint32_t)c0, a);
                chol_stage_row(a, hi_base, half_bytes, r);
            }
            tc_fence_before();
            __syncthreads();
            HYPREC_TCPROF(3)
            if (rem0 < dimp) {
                if (tid == 0) {
                    tc_fence_after();
                    if (rem0 < 128) {
                        const int n0 = (two ? 128 : dimp) - rem0;
                        chol_issue_update(st.tmem_base + (uint32_t)rem0, hi_base, half_bytes, 0, rem0, n0);
                    }
                    if (two) chol_issue_update(st.tmem_base + 128u + (uint32_t)rem0, hi_base, half_bytes, 128, rem0, dimp - rem0);
                    umma_commit(&bars[2]);
                }
                if (warp > pn && r < dimp) {
                    // the rows below take this panel's z out of their right-hand side
                    float acc = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc = fmaf(a[j], zs[c0 + j], acc);
                    bs[r] -= acc;
                }
                fetch_gram(rem0);
                mbar_wait(&bars[2], upd_phase);
                upd_phase ^= 1;
                tc_fence_after();
            }
            HYPREC_TCPROF(5)
        }

        if (p.cross_out != nullptr) {                  // |z|^2, z = L^-1 b complete
            float q = tid < dim ? zs[tid] * zs[tid] : 0.f;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) q += __shfl_xor_sync(0xffffffffu, q, s);
            if (lane == 0) red[warp] = q;
        }

        // ---- 3. back substitution L^T x = z, last panel first ---------------------------------
        float xmine = 0.f;
        for (int pn = n_panels - 1; pn >= 0; --pn) {
            const int c0 = pn * kCholPanel;
            const int r = tid;
            if (warp > pn && r < dimp) {
                uint32_t raw[32];
                tmem_load_32x32(my_taddr + (uint32_t)c0, raw);
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]) * xmine;
                // transpose-reduce: afterwards lane l holds sum over the warp's rows of entry l
                warp_transpose_reduce(v, lane);
                part[warp * 32 + lane] = v[0];
            } else if (warp == pn) {
                uint32_t raw[32];
                tmem_load_32x32(my_taddr + (uint32_t)c0, raw);
#pragma unroll
                for (int j = 0; j < 32; ++j) l11[lane * 33 + j] = __uint_as_float(raw[j]);
            }
            __syncthreads();
            if (warp == pn) {
                float rhs = zs[c0 + lane];
                const int last_warp = dimp / 32 - 1;
                for (int w = pn + 1; w <= last_warp; ++w) rhs -= part[w * 32 + lane];
                xmine = chol_backward_diag(rhs, lane, l11, dinv + c0);
                xs[c0 + lane] = xmine;
            }
            __syncthreads();
        }
        HYPREC_TCPROF(7)
        if (p.prof != nullptr && blockIdx.x == 0 && tid == 0) {
            p.prof[8] += 1;
            p.prof[9] += deg;
        }
        if (p.cross_out != nullptr && tid == 0) {      // (the back substitution's barriers published red[])
            double zz = 0.0, rr = 0.0;
            for (int w = 0; w < kCholThreads / 32; ++w) {
                zz += (double)red[w];
                if (lowrank) rr += (double)red[8 + w];
            }
            p.cross_out[row] = lowrank ? (rr - zz) / (double)pos_scale : zz;
        }

        // ---- 4. output ------------------------------------------------------------------------
        if (!lowrank) {
            if (tid < k) {
                const float v = xs[tid];
                if (!(v - v == 0.f)) *p.status = 1;    // inf or nan: singular system (lambda == 0)
                out[tid] = v;
            }
        } else {
            // y = Q_r^T t + p  (second, L2-resident pass over the row's Q rows).  Four thread groups of 64 take every
            // fourth row each, a thread its 4 columns with 16-byte loads, eight rows in flight; the groups' partial
            // sums meet in shared memory (l11 is free after the back substitution).
            // (k may exceed 256 -- only the system is limited to 256 unknowns: 256 columns per round)
            const int grp = tid >> 6;
            for (int cb = 0; cb < k; cb += 256) {
                const int c4 = cb + (tid & 63) * 4;
                float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c4 < k) {
                    for (int i0 = grp; i0 < dim; i0 += 32) {
                        float4 q[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int i = i0 + 4 * u < dim ? i0 + 4 * u : i0;     // (the repeats get weight 0 below)
                            q[u] = __ldg(reinterpret_cast<const float4*>(p.fixed + (int64_t)idxs[i] * k + c4));
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const float w = i0 + 4 * u < dim ? xs[i0 + 4 * u] : 0.f;
                            acc4.x = fmaf(q[u].x, w, acc4.x);
                            acc4.y = fmaf(q[u].y, w, acc4.y);
                            acc4.z = fmaf(q[u].z, w, acc4.z);
                            acc4.w = fmaf(q[u].w, w, acc4.w);
                        }
                    }
                }
                __syncthreads();                           // (l11: its last readers -- back substitution, previous round -- are done)
                if (c4 < k) *reinterpret_cast<float4*>(l11 + grp * 256 + (c4 - cb)) = acc4;
                __syncthreads();
                const int a = cb + tid;
                if (a < k) {
                    float s = prior_row ? p.lambda * prior_row[a] : 0.f;
                    s += (l11[tid] + l11[256 + tid]) + (l11[512 + tid] + l11[768 + tid]);
                    if (!(s - s == 0.f)) *p.status = 1;
                    out[a] = s;
                }
            }
        }
        HYPREC_TCPROF(10)
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(st.tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace tc
}  // namespace hyprec
