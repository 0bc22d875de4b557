// Tensor-core construction of one row's normal matrix inside the fused row-solve kernel
// (als_solve.cuh), float32 mode, sm_100a.
//
// Replaces the arithmetic of `(Y.T * confidence).dot(Y)` restricted to the row's positives
// (/root/reference/lib/collaborative_filtering.py:84-85, :92), in the two forms the row kernel uses:
//   low-rank rows   S = I + 0.9 Q_r Q_r^T      Q_r = the deg gathered rows of the whitened factors;
//                   a [deg x k] x [k x deg] product, both operands K-major (a gathered row is
//                   contiguous along k)
//   direct rows     A = base + 0.9 Y_r^T Y_r   Y_r = the deg gathered factor rows; a [k x deg] x [deg x k]
//                   product whose contraction runs over the gathered rows, so each thread owns one
//                   factor index, reads it from 32 gathered rows (coalesced across the CTA) and
//                   stores 4 consecutive rows' values as one 16-byte chunk: the transpose into the
//                   K-major operand layout happens in registers.  b = sum r y falls out of the same
//                   loads.
// The gathered rows are split x = hi + lo (both TF32) on their way into shared memory and the product is
// accumulated as hi*hi + hi*lo + lo*hi in float32 (3xTF32, error ~2^-22 per term: float32-level),
// by tcgen05.mma with the accumulator in tensor memory.  Only the gather (coalesced 16-byte loads,
// 8 in flight per thread), the split and the final TMEM -> tile copy run on the SIMT cores; the
// gather of k-block b+1 overlaps the MMAs of k-block b (two staging stages, mbarrier hand-over).
//
// Staging layout per stage: `rows_pad` rows x 128 bytes (32 floats of one k-block), SWIZZLE_128B
// K-major (16-byte chunk c of row r at (r/8)*1024 + (r%8)*128 + ((c ^ (r%8)) * 16)), hi rows first,
// then the lo rows.  The stages alias the tile storage of the system being built: the tiles are
// written only after the last MMA has retired.
//
// Accumulators: D00 = rows 0..127 x cols 0..127 at TMEM column 0; for systems above 128 rows
// D1 = rows 128..255 x cols 0..255 at TMEM column 128 (the upper-right block is never formed).
#pragma once
#include "tc_score.cuh"

namespace hyprec {
namespace tc {

constexpr int kSyrkKB = 32;   // factors per k-block = one 128-byte swizzle row

__host__ __device__ __forceinline__ int syrk_rows_pad(int max_dim) { return max_dim <= 128 ? 128 : 256; }
// both stages (hi + lo each)
__host__ __device__ __forceinline__ size_t syrk_stage_bytes(int max_dim) { return (size_t)syrk_rows_pad(max_dim) * 128 * 2; }
__host__ __device__ __forceinline__ int syrk_tmem_cols(int max_dim) { return max_dim <= 128 ? 128 : 512; }

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    uint32_t h, l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = __uint_as_float(h);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - hi));
    lo = __uint_as_float(l);
}

struct SyrkState {
    uint64_t* bars;       // [2] one per staging stage, count 1 (tcgen05.commit arrives)
    uint32_t phase[2];    // parity of the next completion to wait for
    uint32_t tmem_base;
    unsigned char* stage; // 1024-byte aligned, 2 * syrk_stage_bytes
    int rows_pad;
};


// One k-block (32 contraction steps) of D += X X^T from a filled stage, three TF32 products per K=8 step,
// then a commit that arrives on `bar` when they have retired.  Row block 0 (stage rows 0..127) against
// the n0 stage rows from 0 goes to TMEM column 0; row block 1 (rows 128..255) against the n1 rows from
// b1_row0 goes to TMEM column 128 (n == 0: block not formed).
__device__ __forceinline__ void syrk_issue_blocks(const SyrkState& st, unsigned char* hi_base, size_t half_bytes, int n0, int n1,
                                                  int b1_row0, bool accumulate, uint64_t* bar) {
    tc_fence_after();
    const uint32_t sa = smem_u32(hi_base);
    const uint64_t a0_hi = umma_desc(sa), a0_lo = umma_desc(sa + (uint32_t)half_bytes);
    const uint64_t a1_hi = umma_desc(sa + 128 * 128), a1_lo = umma_desc(sa + (uint32_t)half_bytes + 128 * 128);
    const uint64_t b1_hi = umma_desc(sa + (uint32_t)b1_row0 * 128), b1_lo = umma_desc(sa + (uint32_t)half_bytes + (uint32_t)b1_row0 * 128);
    const uint32_t idesc0 = umma_idesc_tf32(128, n0 > 0 ? n0 : 16);
    const uint32_t idesc1 = umma_idesc_tf32(128, n1 > 0 ? n1 : 16);
#pragma unroll
    for (int ks = 0; ks < kSyrkKB / 8; ++ks) {
        const uint64_t adv = (uint64_t)((ks * 8 * 4) >> 4);
        const uint32_t accum = (accumulate || ks != 0) ? 1u : 0u;
        if (n0 > 0) {
            umma_tf32(st.tmem_base, a0_hi + adv, a0_hi + adv, idesc0, accum);
            umma_tf32(st.tmem_base, a0_hi + adv, a0_lo + adv, idesc0, 1);
            umma_tf32(st.tmem_base, a0_lo + adv, a0_hi + adv, idesc0, 1);
        }
        if (n1 > 0) {
            umma_tf32(st.tmem_base + 128, a1_hi + adv, b1_hi + adv, idesc1, accum);
            umma_tf32(st.tmem_base + 128, a1_hi + adv, b1_lo + adv, idesc1, 1);
            umma_tf32(st.tmem_base + 128, a1_lo + adv, b1_hi + adv, idesc1, 1);
        }
    }
    umma_commit(bar);
}

// single system: D00 (and D1 = rows 128.. x all columns for > 128 rows)
__device__ __forceinline__ void syrk_issue(const SyrkState& st, unsigned char* hi_base, size_t half_bytes, int npad, bool two,
                                           bool accumulate, uint64_t* bar) {
    syrk_issue_blocks(st, hi_base, half_bytes, two ? 128 : npad, two ? npad : 0, 0, accumulate, bar);
}

// TMEM -> tiles: lane = system row, 32 columns per load; only tiles on or below the diagonal.
// A[i][j] = scale * D[i][j] + diag * (i == j) for i, j < dim.
__device__ __forceinline__ void syrk_epilogue(const SyrkState& st, float* A, int NT, int dim, float scale, float diag) {
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nwarps = blockDim.x / 32;
    const bool two = dim > 128;
    const int quad = warp & 3, half = warp >> 2, nhalf = nwarps >> 2;
    for (int mb = 0; mb < (two ? 2 : 1); ++mb) {
        const int row_base = mb * 128 + quad * 32;
        if (row_base >= dim) continue;
        const int i = row_base + lane;
        const int jmax = min(dim, row_base + 32);
        const int bi = i >> 3, ri = i & 7;
        for (int c = 0; c < jmax; c += 32) {
            if (((c >> 5) % nhalf) != half) continue;
            uint32_t v[32];
            const uint32_t taddr = st.tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mb ? 128 : 0) + (uint32_t)c;
            tmem_load_32x32(taddr, v);
            if (i < dim) {
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) {
                    const int j = c + jj;
                    const int bj = j >> 3;
                    if (j < dim && bj <= bi)
                        A[(size_t)(ri * 8 + (j & 7)) * NT + (bi * (bi + 1) / 2 + bj)] =
                            scale * __uint_as_float(v[jj]) + (i == j ? diag : 0.f);
                }
            }
        }
    }
}

// D (TMEM) = pre_scale^2 * Q_r Q_r^T for the `dim` rows idxs[0..dim) of `fixed` ([.][k], k % 4 == 0);
// `npad` (a multiple of 16, >= dim) columns are formed.  All threads of the CTA call it (blockDim.x a
// multiple of 128).
// diag_blocks: several independent systems side by side (tc_chol_multi.cuh) -- rows without a positive
// carry id -1 (zero rows) and the two row blocks only meet their own 128 columns.
__device__ __forceinline__ void syrk_lowrank_accumulate(const float* __restrict__ fixed, const int* idxs, int dim, int k,
                                                        SyrkState& st, float pre_scale, int npad, bool diag_blocks = false) {
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int chunk = tid & 7, r0 = tid >> 3, rpp = nthreads >> 3;
    const int nkb = (k + kSyrkKB - 1) / kSyrkKB;
    const size_t half_bytes = (size_t)st.rows_pad * 128;          // hi rows, then lo rows
    const size_t stage_bytes = 2 * half_bytes;
    const bool two = dim > 128;
    bool pending[2] = {false, false};

    // the gather runs two k-blocks ahead of the stores (two register buffers): a k-block's load latency
    // is covered by the split / store / MMA issue of the two blocks before it
    auto load = [&](int kb, float4 (&v)[8]) {
        const int col = kb * kSyrkKB + chunk * 4;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int r = r0 + j * rpp;
            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < dim && col < k) {
                const int id = idxs[r];
                if (id >= 0) v[j] = __ldg(reinterpret_cast<const float4*>(fixed + (int64_t)id * k + col));
            }
        }
    };
    auto stage_and_issue = [&](int kb, float4 (&v)[8]) {
        const int s = kb & 1;
        if (pending[s]) {     // the MMAs that read this stage two k-blocks ago have retired
            mbar_wait(&st.bars[s], st.phase[s]);
            st.phase[s] ^= 1;
            pending[s] = false;
        }
        unsigned char* hi_base = st.stage + (size_t)s * stage_bytes;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int r = r0 + j * rpp;
            if (r < dim) {
                float4 h, l;
                split_tf32(v[j].x * pre_scale, h.x, l.x);
                split_tf32(v[j].y * pre_scale, h.y, l.y);
                split_tf32(v[j].z * pre_scale, h.z, l.z);
                split_tf32(v[j].w * pre_scale, h.w, l.w);
                const size_t off = (size_t)(r >> 3) * 1024 + (size_t)(r & 7) * 128 + (size_t)((chunk ^ (r & 7)) << 4);
                *reinterpret_cast<float4*>(hi_base + off) = h;
                *reinterpret_cast<float4*>(hi_base + half_bytes + off) = l;
            }
        }
        fence_async_smem();      // generic-proxy stores -> visible to the tensor core's async proxy
        __syncthreads();
        if (tid == 0) {
            if (diag_blocks) syrk_issue_blocks(st, hi_base, half_bytes, 128, 128, 128, kb != 0, &st.bars[s]);
            else syrk_issue(st, hi_base, half_bytes, npad, two, kb != 0, &st.bars[s]);
        }
        pending[s] = true;
    };
    float4 va[8], vb[8];
    load(0, va);
    if (nkb > 1) load(1, vb);
    for (int kb = 0; kb < nkb; kb += 2) {
        stage_and_issue(kb, va);
        if (kb + 2 < nkb) load(kb + 2, va);
        if (kb + 1 < nkb) {
            stage_and_issue(kb + 1, vb);
            if (kb + 3 < nkb) load(kb + 3, vb);
        }
    }
#pragma unroll
    for (int s = 0; s < 2; ++s)
        if (pending[s]) {
            mbar_wait(&st.bars[s], st.phase[s]);
            st.phase[s] ^= 1;
        }
    tc_fence_after();
    __syncthreads();          // every thread is past its last staging access: the stages may be reused
}

// S = I + scale * Q_r Q_r^T written into the TS=8 tile storage A (lower-triangle tiles, full diagonal
// tiles); A aliases the staging stages.
__device__ __forceinline__ void syrk_lowrank(const float* __restrict__ fixed, const int* idxs, int dim, int k,
                                             SyrkState& st, float* A, int NT, float scale) {
    syrk_lowrank_accumulate(fixed, idxs, dim, k, st, 1.f, (dim + 15) & ~15);
    syrk_epilogue(st, A, NT, dim, scale, 1.f);
    tc_fence_before();
}


// D (TMEM) = pre_scale^2 * Y_r^T Y_r (k x k) for the deg rows indices[lo .. lo+deg) of `fixed`, and
// bvec[a] = sum_j values[lo+j] * Y[j][a] (k even, k <= 256).
// idxs / rstage: shared scratch of `batch` entries (a multiple of 32).
__device__ __forceinline__ void syrk_direct_accumulate(const float* __restrict__ fixed, const int32_t* __restrict__ indices,
                                                       const float* __restrict__ values, int64_t lo, int deg, int k,
                                                       SyrkState& st, float pre_scale, int npad, int* idxs, float* rstage,
                                                       int batch, float* bvec) {
    // blockDim.x == 256: thread t owns the factor pair a = 2 (t % 128), a + 1 for the 16 gathered rows
    // 16 (t / 128) .. +15 of every k-block (8-byte loads, one address computation per row and pair)
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int half = tid >> 7, a = (tid & 127) * 2;
    const int nkb = (deg + kSyrkKB - 1) / kSyrkKB;
    const size_t half_bytes = (size_t)st.rows_pad * 128;
    const size_t stage_bytes = 2 * half_bytes;
    const bool two = k > 128;
    const int kb_per_batch = batch / kSyrkKB;
    bool pending[2] = {false, false};
    float bsum0 = 0.f, bsum1 = 0.f;

    // ids and ratings of a whole batch of k-blocks are staged at once; the factor loads run one k-block
    // ahead of the stores (two register buffers)
    auto stage_ids = [&](int kb) {
        for (int i = tid; i < batch; i += nthreads) {
            const int j = kb * kSyrkKB + i;
            idxs[i] = j < deg ? (indices ? indices[lo + j] : (int)(lo + j)) : -1;   // null: consecutive rows
            rstage[i] = (values && j < deg) ? values[lo + j] : 0.f;
        }
        __syncthreads();
    };
    auto load = [&](int kb, float2 (&v)[16]) {
        const int slot = (kb % kb_per_batch) * kSyrkKB + 16 * half;
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
            const int id = idxs[slot + jj];
            v[jj] = (a < k && id >= 0) ? __ldg(reinterpret_cast<const float2*>(fixed + (int64_t)id * k + a)) : make_float2(0.f, 0.f);
        }
    };
    auto stage_and_issue = [&](int kb, float2 (&v)[16]) {
        const int s = kb & 1;
        const int slot = (kb % kb_per_batch) * kSyrkKB + 16 * half;
        if (pending[s]) {
            mbar_wait(&st.bars[s], st.phase[s]);
            st.phase[s] ^= 1;
            pending[s] = false;
        }
        unsigned char* hi_base = st.stage + (size_t)s * stage_bytes;
        if (a < st.rows_pad) {
            const size_t row0 = (size_t)(a >> 3) * 1024 + (size_t)(a & 7) * 128;          // row a (a is even: a + 1 is
            const size_t row1 = row0 + 128;                                                 // in the same 8-row group)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float4 h0, l0, h1, l1;
                split_tf32(v[4 * c].x * pre_scale, h0.x, l0.x);
                split_tf32(v[4 * c + 1].x * pre_scale, h0.y, l0.y);
                split_tf32(v[4 * c + 2].x * pre_scale, h0.z, l0.z);
                split_tf32(v[4 * c + 3].x * pre_scale, h0.w, l0.w);
                split_tf32(v[4 * c].y * pre_scale, h1.x, l1.x);
                split_tf32(v[4 * c + 1].y * pre_scale, h1.y, l1.y);
                split_tf32(v[4 * c + 2].y * pre_scale, h1.z, l1.z);
                split_tf32(v[4 * c + 3].y * pre_scale, h1.w, l1.w);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float rt = rstage[slot + 4 * c + e];
                    bsum0 = fmaf(rt, v[4 * c + e].x, bsum0);
                    bsum1 = fmaf(rt, v[4 * c + e].y, bsum1);
                }
                const int chunk = 4 * half + c;
                const size_t off0 = row0 + (size_t)((chunk ^ (a & 7)) << 4);
                const size_t off1 = row1 + (size_t)((chunk ^ ((a + 1) & 7)) << 4);
                *reinterpret_cast<float4*>(hi_base + off0) = h0;
                *reinterpret_cast<float4*>(hi_base + half_bytes + off0) = l0;
                *reinterpret_cast<float4*>(hi_base + off1) = h1;
                *reinterpret_cast<float4*>(hi_base + half_bytes + off1) = l1;
            }
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) syrk_issue(st, hi_base, half_bytes, npad, two, kb != 0, &st.bars[s]);
        pending[s] = true;
    };
    // a batch's ids may only be replaced once no load of the batch is outstanding: the look-ahead
    // stops at batch boundaries
    float2 va[16], vb[16];
    for (int kb0 = 0; kb0 < nkb; kb0 += kb_per_batch) {
        const int kb1 = min(nkb, kb0 + kb_per_batch);
        stage_ids(kb0);
        load(kb0, va);
        for (int kb = kb0; kb < kb1; kb += 2) {
            if (kb + 1 < kb1) load(kb + 1, vb);
            stage_and_issue(kb, va);
            if (kb + 1 < kb1) {
                if (kb + 2 < kb1) load(kb + 2, va);
                stage_and_issue(kb + 1, vb);
            }
        }
    }
#pragma unroll
    for (int s = 0; s < 2; ++s)
        if (pending[s]) {
            mbar_wait(&st.bars[s], st.phase[s]);
            st.phase[s] ^= 1;
        }
    tc_fence_after();
    if (half == 0 && a < k) {
        bvec[a] = bsum0;
        bvec[a + 1] = bsum1;
    }
    __syncthreads();
    if (half == 1 && a < k) {
        bvec[a] += bsum0;
        bvec[a + 1] += bsum1;
    }
    __syncthreads();
}

// A = scale * Y_r^T Y_r as tiles (no diagonal term: the caller adds the shared base); bvec = sum r y.
__device__ __forceinline__ void syrk_direct(const float* __restrict__ fixed, const int32_t* __restrict__ indices,
                                            const float* __restrict__ values, int64_t lo, int deg, int k, SyrkState& st,
                                            float* A, int NT, float scale, int* idxs, float* rstage, int batch, float* bvec) {
    syrk_direct_accumulate(fixed, indices, values, lo, deg, k, st, 1.f, (k + 15) & ~15, idxs, rstage, batch, bvec);
    syrk_epilogue(st, A, NT, k, scale, 0.f);
    tc_fence_before();
}


// ------------------------------------------------------------------------------------------------
// Gram matrix G = Y^T Y of a factor matrix on the tensor cores (float32 mode): the same transposing
// accumulation as the k x k row systems, over consecutive rows.  Replaces the shared
// `fixed_vecs.T.dot(fixed_vecs)` part of (Y.T * confidence).dot(Y)
// (/root/reference/lib/collaborative_filtering.py:84, :92).
// One CTA per SM, each over a contiguous slice of rows; the k x k partial (lower blocks) goes to
// `partial[blockIdx.x]`, gram_tc_reduce_kernel sums the slices in float64 and mirrors the result.
// ------------------------------------------------------------------------------------------------
struct GramTcLayout {
    size_t off_stage, off_idx, off_rst, off_b, off_bar, total;
};
__host__ __device__ __forceinline__ GramTcLayout gram_tc_layout(int k) {
    GramTcLayout l;
    size_t o = 0;
    l.off_stage = o; o += 2 * syrk_stage_bytes(k);
    l.off_idx = o;   o += sizeof(int) * 256;
    l.off_rst = o;   o += sizeof(float) * 256;
    l.off_b = o;     o += sizeof(float) * 256;
    l.off_bar = o;
    o += 32;
    l.total = o + 1024;
    return l;
}

__global__ void __launch_bounds__(256, 1) gram_tc_kernel(const float* __restrict__ y, int64_t n_rows, int k,
                                                         int64_t rows_per_cta, float* __restrict__ partial) {
    extern __shared__ __align__(16) unsigned char gram_smem_dyn[];
    unsigned char* smem = gram_smem_dyn + ((1024u - (smem_u32(gram_smem_dyn) & 1023u)) & 1023u);
    const GramTcLayout lay = gram_tc_layout(k);
    uint64_t* bars = (uint64_t*)(smem + lay.off_bar);
    uint32_t* tmem_slot = (uint32_t*)(bars + 2);
    const int tid = threadIdx.x, warp = tid / 32;
    const uint32_t tmem_cols = (uint32_t)syrk_tmem_cols(k);
    SyrkState st;
    st.bars = bars;
    st.phase[0] = st.phase[1] = 0;
    st.stage = smem + lay.off_stage;
    st.rows_pad = syrk_rows_pad(k);
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    st.tmem_base = *tmem_slot;

    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r1 = r0 + rows_per_cta < n_rows ? r0 + rows_per_cta : n_rows;
    const int deg = r1 > r0 ? (int)(r1 - r0) : 0;
    float* out = partial + (size_t)blockIdx.x * k * k;
    const int npad = (k + 15) & ~15;
    if (deg > 0)
        syrk_direct_accumulate(y, nullptr, nullptr, r0, deg, k, st, 1.f, npad, (int*)(smem + lay.off_idx),
                               (float*)(smem + lay.off_rst), 256, (float*)(smem + lay.off_b));
    // row i = tid: columns 0 .. 32 * (i / 32) + 31 (the row block's part on or below the diagonal blocks)
    const int i = tid;
    if (i < ((k + 31) & ~31)) {     // whole warps
        const uint32_t taddr = st.tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(tid >= 128 ? 128 : 0);
        const int jmax = min(k, (i / 32) * 32 + 32);
        for (int c = 0; c < jmax; c += 32) {
            uint32_t v[32];
            if (deg > 0) tmem_load_32x32(taddr + (uint32_t)c, v);
            if (i < k) {
#pragma unroll
                for (int jj = 0; jj < 32; ++jj)
                    if (c + jj < k) out[(size_t)i * k + c + jj] = deg > 0 ? __uint_as_float(v[jj]) : 0.f;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(st.tmem_base, tmem_cols);
}

// G[i][j] = G[j][i] = sum over slices of partial[.][max(i,j)][min(i,j)]
__global__ void gram_tc_reduce_kernel(const float* __restrict__ partial, int n_slices, int k, float* __restrict__ g) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= k * k) return;
    const int i = idx / k, j = idx % k;
    if (j > i) return;
    // entries right of the diagonal 32-block of row i were not written: (i, j) with j <= i is always inside
    double s = 0.0;
    for (int c = 0; c < n_slices; ++c) s += (double)partial[(size_t)c * k * k + idx];
    g[(size_t)i * k + j] = (float)s;
    g[(size_t)j * k + i] = (float)s;
}


// ------------------------------------------------------------------------------------------------
// Rows with tens of thousands of positives (popular items of the scaled workload) split over several
// CTAs: every segment of such a row's positives is accumulated like a Gram slice -- the same
// transposing tcgen05 accumulation, over the gathered factor rows indices[seg_lo .. seg_lo+seg_len) --
// into partial[seg] (k x k, lower blocks) and partial_b[seg] = sum r y; split_reduce_kernel folds a
// row's segments into the dense matrix tc_chol_rows_kernel adds in place of the shared Gram.  The
// arithmetic is that of (Y.T * confidence).dot(Y) restricted to the row's positives
// (/root/reference/lib/collaborative_filtering.py:84-86, :92), only distributed differently.
// ------------------------------------------------------------------------------------------------
struct SplitArgs {
    const float* fixed;        // factors of the fixed side [.][k]
    const int32_t* indices;    // CSR column ids of the side being solved
    const float* values;       // ratings
    const int64_t* seg_lo;     // [n_seg] first positive of the segment (absolute offset into indices)
    const int32_t* seg_len;    // [n_seg]
    int n_seg;
    int k;
    float* partial;            // [n_seg][k][k]
    float* partial_b;          // [n_seg][k]
    int* work_counter;         // zeroed before launch
};

__global__ void __launch_bounds__(256, 1) split_partial_kernel(SplitArgs p) {
    extern __shared__ __align__(16) unsigned char split_smem_dyn[];
    unsigned char* smem = split_smem_dyn + ((1024u - (smem_u32(split_smem_dyn) & 1023u)) & 1023u);
    const int k = p.k;
    const GramTcLayout lay = gram_tc_layout(k);
    uint64_t* bars = (uint64_t*)(smem + lay.off_bar);
    uint32_t* tmem_slot = (uint32_t*)(bars + 2);
    float* bvec = (float*)(smem + lay.off_b);
    __shared__ int s_seg;
    const int tid = threadIdx.x, warp = tid / 32;
    const uint32_t tmem_cols = (uint32_t)syrk_tmem_cols(k);
    SyrkState st;
    st.bars = bars;
    st.phase[0] = st.phase[1] = 0;
    st.stage = smem + lay.off_stage;
    st.rows_pad = syrk_rows_pad(k);
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    st.tmem_base = *tmem_slot;
    const int npad = (k + 15) & ~15;
    while (true) {
        __syncthreads();
        if (tid == 0) s_seg = atomicAdd(p.work_counter, 1);
        __syncthreads();
        const int seg = s_seg;
        if (seg >= p.n_seg) break;
        const int len = p.seg_len[seg];
        syrk_direct_accumulate(p.fixed, p.indices, p.values, p.seg_lo[seg], len, k, st, 1.f, npad, (int*)(smem + lay.off_idx),
                               (float*)(smem + lay.off_rst), 256, bvec);
        float* out = p.partial + (size_t)seg * k * k;
        const int i = tid;
        if (i < ((k + 31) & ~31)) {     // whole warps; row i: columns up to the end of its diagonal 32-block
            const uint32_t taddr = st.tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(tid >= 128 ? 128 : 0);
            const int jmax = min(k, (i / 32) * 32 + 32);
            for (int c = 0; c < jmax; c += 32) {
                uint32_t v[32];
                tmem_load_32x32(taddr + (uint32_t)c, v);
                if (i < k) {
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj)
                        if (c + jj < k) out[(size_t)i * k + c + jj] = __uint_as_float(v[jj]);
                }
            }
        }
        if (tid < k) p.partial_b[(size_t)seg * k + tid] = bvec[tid];
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(st.tmem_base, tmem_cols);
}

// split_gram[s][i][j] = gram[i][j] + ratio * sum_{segments of s} partial[.][max(i,j)][min(i,j)]  (ratio = 0.9 / 0.1),
// split_b[s][a] = sum partial_b; grid = (ceil(k*k / 256), n_split).
__global__ void split_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ partial_b,
                                    const int32_t* __restrict__ seg_ptr, const float* __restrict__ gram, int k, float ratio,
                                    float* __restrict__ split_gram, float* __restrict__ split_b) {
    const int s = blockIdx.y;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int s0 = seg_ptr[s], s1 = seg_ptr[s + 1];
    if (idx < k) {
        double b = 0.0;
        for (int c = s0; c < s1; ++c) b += (double)partial_b[(size_t)c * k + idx];
        split_b[(size_t)s * k + idx] = (float)b;
    }
    if (idx >= k * k) return;
    const int i = idx / k, j = idx % k;
    if (j > i) return;
    double acc = 0.0;
    for (int c = s0; c < s1; ++c) acc += (double)partial[(size_t)c * k * k + idx];
    const float v = (float)((double)gram[idx] + (double)ratio * acc);
    split_gram[(size_t)s * k * k + (size_t)i * k + j] = v;
    split_gram[(size_t)s * k * k + (size_t)j * k + i] = v;
}

}  // namespace tc
}  // namespace hyprec
