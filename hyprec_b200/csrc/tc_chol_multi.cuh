// Several small low-rank row systems per CTA pass, all resident in tensor memory (float32 mode).
//
// Rows with few positives give small S = I + 0.9 Q_r Q_r^T systems (the median
// citeulike row has ~10-30 positives).  One 256-row pass of tc_chol.cuh would leave most of the CTA idle
// on them, so here 256 / SLOT systems of up to SLOT = 32, 64 or 128 unknowns sit side by side:
// system s owns the operand-stage rows / TMEM lanes [s * SLOT, (s+1) * SLOT) and the SLOT TMEM columns
// at the same offset inside its 128-row block, the two 128 x 128 accumulator blocks are built by one
// tcgen05.mma chain per k-block (the cross-system entries are computed and never read), and every
// system runs the panel algorithm of tc_chol.cuh at the same time -- its own diagonal warp, its own
// L11, its own trailing-update MMA restricted to its own columns.  Thread t still owns system row t.
//
// Reference arithmetic: as in als_solve.cuh (low-rank form of collaborative_filtering.py:83-99).
#pragma once
#include "tc_chol.cuh"

namespace hyprec {
namespace tc {

struct CholMultiLayout {
    size_t off_stage, off_l11, off_dinv, off_z, off_x, off_b, off_part, off_col, off_idx, off_meta, off_red, off_bar, total;
};

__host__ __device__ __forceinline__ CholMultiLayout chol_multi_layout() {
    CholMultiLayout l;
    size_t o = 0;
    l.off_stage = o; o += 2 * syrk_stage_bytes(256);
    l.off_l11 = o;   o += sizeof(float) * 8 * 33 * 32;   // per system: L11 (transposed going forward, row-major + pad going back)
    l.off_dinv = o;  o += sizeof(float) * 256;
    l.off_z = o;     o += sizeof(float) * 256;
    l.off_x = o;     o += sizeof(float) * 256;
    l.off_b = o;     o += sizeof(float) * 256;
    l.off_part = o;  o += sizeof(float) * 8 * 32;
    l.off_col = o;   o += sizeof(float) * 8 * 64;        // per warp: column exchange of the diagonal factor
    l.off_idx = o;   o += sizeof(int) * 256;
    l.off_meta = o;  o += 8 * (8 + 8 + 8);               // per system: row id, CSR offset, dimension
    l.off_red = o;   o += sizeof(float) * 16;            // per-warp partial sums of |r|^2 and |z|^2
    l.off_bar = (o + 7) & ~(size_t)7;
    o = l.off_bar + 32;
    l.total = o + 1024;
    return l;
}

template <int SLOT>
__global__ void __launch_bounds__(kCholThreads, 1) tc_chol_multi_kernel(CholRowsArgs p) {
    constexpr int NSYS = 256 / SLOT;          // systems per pass
    constexpr int WPS = SLOT / 32;            // warps (= panels at most) per system
    extern __shared__ __align__(16) unsigned char cholm_smem_dyn[];
    unsigned char* smem = cholm_smem_dyn + ((1024u - (smem_u32(cholm_smem_dyn) & 1023u)) & 1023u);
    const CholMultiLayout lay = chol_multi_layout();
    float* l11_all = (float*)(smem + lay.off_l11);
    float* dinv = (float*)(smem + lay.off_dinv);
    float* zs = (float*)(smem + lay.off_z);
    float* xs = (float*)(smem + lay.off_x);
    float* bs = (float*)(smem + lay.off_b);
    float* part = (float*)(smem + lay.off_part);
    float* colbufs = (float*)(smem + lay.off_col);
    int* idxs = (int*)(smem + lay.off_idx);
    long long* sys_row = (long long*)(smem + lay.off_meta);
    long long* sys_lo = sys_row + 8;
    long long* sys_dim = sys_lo + 8;
    float* red = (float*)(smem + lay.off_red);
    uint64_t* bars = (uint64_t*)(smem + lay.off_bar);
    uint32_t* tmem_slot = (uint32_t*)(bars + 3);
    __shared__ long long s_first;

    const int tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
    const int sys = tid / SLOT, lr = tid % SLOT, lw = lr / 32;     // my system, my row in it, my panel index
    const int k = p.k;
    const float sqrt_pos = 0.9486832980505138f;
    const float pos_scale = (float)(kConfPos - kConfNeg);
    const int64_t n_work = p.row_list ? p.n_list : p.row_hi - p.row_lo;
    long long prof_t = 0;

    SyrkState st;
    st.bars = bars;
    st.phase[0] = st.phase[1] = 0;
    st.stage = smem + lay.off_stage;
    st.rows_pad = 256;
    uint32_t upd_phase = 0;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(tmem_slot, 256u);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    st.tmem_base = *tmem_slot;
    const size_t half_bytes = (size_t)256 * 128;
    // row t: TMEM lane t % 128 of block t / 128; the system's columns start at its row offset in the block
    const uint32_t my_taddr = st.tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(tid >= 128 ? 128 : 0) +
                              (uint32_t)(((tid % 128) / SLOT) * SLOT);
    float* l11s = l11_all + (size_t)sys * 33 * 32;       // this system's L11
    float* colbuf = colbufs + warp * 64;

    while (true) {
        __syncthreads();
        if (tid == 0) s_first = (long long)atomicAdd(p.work_counter, NSYS);
        __syncthreads();
        if (s_first >= n_work) break;
        if (p.prof != nullptr && blockIdx.x == 0 && tid == 0) prof_t = clock64();
        if (tid < NSYS) {
            const long long w = s_first + tid;
            long long row = -1, lo = 0, deg = 0;
            if (w < n_work) {
                row = p.row_list ? (long long)p.row_list[w] : p.row_lo + w;
                lo = p.indptr[row];
                deg = p.indptr[row + 1] - lo;
            }
            sys_row[tid] = row;
            sys_lo[tid] = lo;
            sys_dim[tid] = deg;
        }
        __syncthreads();
        const int dim = (int)sys_dim[sys];
        const int dimp = (dim + kCholPanel - 1) / kCholPanel * kCholPanel;
        int max_pan = 0;
#pragma unroll
        for (int s = 0; s < NSYS; ++s) max_pan = max(max_pan, ((int)sys_dim[s] + kCholPanel - 1) / kCholPanel);

        // ---- 1. ids, right-hand sides r - 0.9 Q_r p, D = 0.9 Q_r Q_r^T for all systems ----------
        idxs[tid] = lr < dim ? p.indices[sys_lo[sys] + lr] : -1;
        __syncthreads();
        if (p.prior == nullptr) {
            bs[tid] = lr < dim ? p.values[sys_lo[sys] + lr] : 0.f;
        } else {
            for (int i = warp; i < 256; i += kCholThreads / 32) {
                const int si = i / SLOT, li = i % SLOT;
                float v = 0.f;
                if (li < (int)sys_dim[si]) {       // (warp-uniform)
                    const float* q = p.fixed + (int64_t)idxs[i] * k;
                    const float* pr = p.prior + sys_row[si] * k;
                    double d = 0.0;
                    for (int l = lane; l < k; l += 32) d += (double)q[l] * (double)pr[l];
#pragma unroll
                    for (int s = 16; s > 0; s >>= 1) d += __shfl_xor_sync(0xffffffffu, d, s);
                    v = p.values[sys_lo[si] + li] - pos_scale * p.lambda * (float)d;
                }
                if (lane == 0) bs[i] = v;
            }
        }
        __syncthreads();
        if (p.cross_out != nullptr) {                   // |r|^2 per warp while bs still holds the ratings
            float q = bs[tid] * bs[tid];
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) q += __shfl_xor_sync(0xffffffffu, q, s);
            if (lane == 0) red[8 + warp] = q;
        }
        if (max_pan > 0) syrk_lowrank_accumulate(p.fixed, idxs, 256, k, st, sqrt_pos, 128, true);
        HYPREC_TCPROF(0)

        // ---- 2. every system's blocked Cholesky, panel by panel ---------------------------------
        for (int pn = 0; pn < max_pan; ++pn) {
            const int c0 = pn * kCholPanel;
            const bool active = lw >= pn && lr < dimp;     // warp-uniform: a warp lies inside one system
            const bool is_diag = active && lw == pn;
            float a[32];
            if (active) {
                uint32_t raw[32];
                tmem_load_32x32(my_taddr + (uint32_t)c0, raw);
                const int dj = lr - c0;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float v = (j == dj) ? 1.f : 0.f;       // I of S = I + 0.9 Q Q^T, and the identity padding
                    if (lr < dim && c0 + j < dim) v += __uint_as_float(raw[j]);
                    a[j] = v;
                }
            }
            if (is_diag) dinv[tid] = chol_factor_diag(a, lane, colbuf, l11s);
            __syncthreads();
            HYPREC_TCPROF(2)
            const int rem0 = c0 + kCholPanel;
            if (is_diag) tmem_store_row(my_taddr + (uint32_t)c0, a);
            if (lw == 0 && c0 < dimp) {
                // the system's first warp (its diagonal warp in panel 0, idle afterwards) takes the panel's
                // right-hand side forward: z_p = L11^-1 b_p
                const int base = sys * SLOT + c0;
                zs[base + lane] = chol_forward_rhs(bs[base + lane], lane, l11s, dinv + base);
            }
            if (active && lw > pn) {
                chol_panel_solve(a, l11s, dinv + sys * SLOT + c0);
                tmem_store_row(my_taddr + (uint32_t)c0, a);
                chol_stage_row(a, st.stage, half_bytes, tid);
            }
            tc_fence_before();
            __syncthreads();
            HYPREC_TCPROF(3)
            if (pn + 1 < max_pan) {
                if (tid == 0) {
                    tc_fence_after();
#pragma unroll
                    for (int s = 0; s < NSYS; ++s) {
                        const int dp = ((int)sys_dim[s] + kCholPanel - 1) / kCholPanel * kCholPanel;
                        if (dp > rem0) {
                            const int row0 = s * SLOT;                 // first stage row / lane of the system
                            const int blk = row0 / 128;
                            chol_issue_update(st.tmem_base + (uint32_t)(blk * 128 + row0 % 128 + rem0), st.stage, half_bytes,
                                              blk * 128, row0 + rem0, dp - rem0);
                        }
                    }
                    umma_commit(&bars[2]);
                }
                if (active && lw > pn) {
                    float acc = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc = fmaf(a[j], zs[sys * SLOT + c0 + j], acc);
                    bs[tid] -= acc;
                }
                mbar_wait(&bars[2], upd_phase);
                upd_phase ^= 1;
                tc_fence_after();
            }
            HYPREC_TCPROF(5)
        }

        if (p.cross_out != nullptr) {                   // |z|^2 per warp (z = 0 on the padding)
            float q = lr < dimp ? zs[tid] * zs[tid] : 0.f;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) q += __shfl_xor_sync(0xffffffffu, q, s);
            if (lane == 0) red[warp] = q;
        }

        // ---- 3. back substitution, last panel first ---------------------------------------------
        float xmine = 0.f;
        for (int pn = max_pan - 1; pn >= 0; --pn) {
            const int c0 = pn * kCholPanel;
            const bool in_sys = lr < dimp && c0 < dimp;
            if (in_sys && lw > pn) {
                uint32_t raw[32];
                tmem_load_32x32(my_taddr + (uint32_t)c0, raw);
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]) * xmine;
                warp_transpose_reduce(v, lane);
                part[warp * 32 + lane] = v[0];
            } else if (in_sys && lw == pn) {
                uint32_t raw[32];
                tmem_load_32x32(my_taddr + (uint32_t)c0, raw);
#pragma unroll
                for (int j = 0; j < 32; ++j) l11s[lane * 33 + j] = __uint_as_float(raw[j]);
            }
            __syncthreads();
            if (in_sys && lw == pn) {
                float rhs = zs[tid];
                for (int w = pn + 1; w < dimp / 32; ++w) rhs -= part[(sys * WPS + w) * 32 + lane];
                xmine = chol_backward_diag(rhs, lane, l11s, dinv + sys * SLOT + c0);
                xs[tid] = xmine;
            }
            __syncthreads();
        }
        HYPREC_TCPROF(7)
        if (p.prof != nullptr && blockIdx.x == 0 && tid == 0) {
            p.prof[8] += NSYS;
            for (int s = 0; s < NSYS; ++s) p.prof[9] += sys_dim[s];
        }

        __syncthreads();
        if (p.cross_out != nullptr && tid < NSYS && sys_row[tid] >= 0) {
            // RMSE cross term of the row: sum_j r_j (q_j . y) = (|r|^2 - |z|^2) / 0.9 (S t = r, y = Q^T t)
            double zz = 0.0, rr = 0.0;
            for (int w = 0; w < WPS; ++w) {
                zz += (double)red[tid * WPS + w];
                rr += (double)red[8 + tid * WPS + w];
            }
            p.cross_out[sys_row[tid]] = (rr - zz) / (double)pos_scale;
        }

        // ---- 4. y = Q_r^T t + p for every system ------------------------------------------------
        // 16-byte loads: a unit is (system, 4 columns, row group) -- 256 columns per round, eight rows in flight per
        // thread.  Two systems per pass leave room for two row groups each (their halves meet in shared memory; the
        // L11 area is free after the back substitution); with 4 / 8 systems a thread walks all rows of its unit.
        constexpr int RG = NSYS == 2 ? 2 : 1;
        constexpr int UNITS = NSYS * 64 * RG;                // per round of 256 columns: 256, 256, 512
        float4* halves = reinterpret_cast<float4*>(l11_all);
        for (int cb = 0; cb < k; cb += 256) {
            for (int u0 = 0; u0 < UNITS; u0 += kCholThreads) {
                const int u = u0 + tid;
                const int s = u / (64 * RG), rem = u % (64 * RG), g = rem / 64, c4 = cb + 4 * (rem % 64);
                const long long row = sys_row[s];
                const bool live = row >= 0 && c4 < k;
                float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live) {
                    const int d = (int)sys_dim[s];
                    const int* ids = idxs + s * SLOT;
                    const float* t = xs + s * SLOT;
                    for (int i0 = g; i0 < d; i0 += 8 * RG) {
                        float4 q[8];
#pragma unroll
                        for (int v = 0; v < 8; ++v) {
                            const int i = i0 + RG * v < d ? i0 + RG * v : i0;      // (the repeats get weight 0 below)
                            q[v] = __ldg(reinterpret_cast<const float4*>(p.fixed + (int64_t)ids[i] * k + c4));
                        }
#pragma unroll
                        for (int v = 0; v < 8; ++v) {
                            const float w = i0 + RG * v < d ? t[i0 + RG * v] : 0.f;
                            acc4.x = fmaf(q[v].x, w, acc4.x);
                            acc4.y = fmaf(q[v].y, w, acc4.y);
                            acc4.z = fmaf(q[v].z, w, acc4.z);
                            acc4.w = fmaf(q[v].w, w, acc4.w);
                        }
                    }
                }
                if (RG == 2) {
                    __syncthreads();                       // (the area's last readers -- back substitution, previous round -- are done)
                    if (g == 1) halves[tid] = acc4;
                    __syncthreads();
                    if (g == 0) {
                        const float4 o = halves[tid + 64];
                        acc4.x += o.x; acc4.y += o.y; acc4.z += o.z; acc4.w += o.w;
                    }
                }
                if (live && g == 0) {
                    if (p.prior) {
                        const float4 pr = *reinterpret_cast<const float4*>(p.prior + row * k + c4);
                        acc4.x = fmaf(p.lambda, pr.x, acc4.x); acc4.y = fmaf(p.lambda, pr.y, acc4.y);
                        acc4.z = fmaf(p.lambda, pr.z, acc4.z); acc4.w = fmaf(p.lambda, pr.w, acc4.w);
                    }
                    const float chk = (acc4.x + acc4.y) + (acc4.z + acc4.w);
                    if (!(chk - chk == 0.f)) *p.status = 1;
                    *reinterpret_cast<float4*>(p.out + (p.out_row0 + s_first + s) * k + c4) = acc4;
                }
            }
        }
        HYPREC_TCPROF(10)
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(st.tmem_base, 256u);
}

}  // namespace tc
}  // namespace hyprec
