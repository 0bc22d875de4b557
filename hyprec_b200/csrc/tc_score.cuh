// Tensor-core score sweep for sm_100a: S~ = U V^T on tcgen05 (kind::tf32, 3xTF32 split),
// operands staged by TMA into 128B-swizzled shared memory, accumulators in TMEM, fused epilogue.
//
// Replaces the arithmetic of `predictions = user_vecs.dot(item_vecs.T)`
// (/root/reference/lib/collaborative_filtering.py:270) where it is only needed approximately:
//   * rounded-prediction counts (abstract_recommender.py:183-186): a cell is decided by the
//     tensor-core score when |s~ - threshold| exceeds a rigorous error margin
//     c * ||u|| * ||v||; the few cells inside the margin are appended to a list and re-evaluated
//     in float64 by recheck_cells_kernel, so the counts are exactly those of the float64 path;
//   * a dense float32 score matrix that the top-k kernel uses as a pre-filter before exact
//     float64 re-scoring (topk.cuh).
//
// 3xTF32: every float64/float32 factor x is split as x = hi + lo (+ ~2^-22 |x|), hi and lo
// representable in TF32, and S~ accumulates hi*hi + hi*lo + lo*hi in float32.
//
// Kernel structure (persistent, one CTA per SM, 320 threads):
//   warp 0      TMA producer: 4 boxes (U_hi, U_lo, V_hi, V_lo; 128 rows x 32 floats, SWIZZLE_128B)
//               per k-block into a 3-stage ring guarded by full/empty mbarriers
//   warp 1      TMEM allocation (256 columns = 2 accumulator stages of 128 x 128 float32) and
//               MMA issue: per k-block 4 k-steps x 3 tcgen05.mma (M128 N128 K8), tcgen05.commit
//               releases the smem stage / publishes the accumulator
//   warps 2-9   epilogue: tcgen05.ld 32x32b.x32 (lane = row; two warps per TMEM lane quadrant, one
//               per column half), float32 margin test and/or float32 store, then hand the TMEM
//               stage back
// All mbarrier waits are bounded: a protocol error traps instead of hanging the GPU.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace hyprec {
namespace tc {

constexpr int kBM = 128, kBN = 128, kBK = 32;       // tile: users x items x factors per k-block
constexpr int kStages = 3;
constexpr int kTileBytes = kBM * kBK * 4;            // one operand box: 16 KB
constexpr int kStageBytes = 4 * kTileBytes;          // U_hi, U_lo, V_hi, V_lo
constexpr int kEpilogueWarps = 8;                    // two per TMEM lane quadrant (column halves)
constexpr int kThreads = 64 + 32 * kEpilogueWarps;
constexpr int kAccCols = kBN;                        // TMEM columns per accumulator stage
constexpr int kTmemCols = 2 * kAccCols;              // power of two >= 32
constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + 4 * kBN;

struct ScoreArgs {
    int64_t m, n;
    int k;
    const double* thresholds;     // [m]            (count mode)
    const float* unorm;           // [m] ||u|| rounded up
    const float* vnorm;           // [n] ||v|| rounded up
    float margin_coef;            // margin = margin_coef * ||u|| * ||v||
    unsigned long long* counts;   // [m] definite score >= threshold (count mode), zeroed by caller
    unsigned int* border_count;   // number of borderline cells appended
    uint2* border_cells;          // (row, col) pairs
    unsigned int border_capacity;
    float* dense;                 // [m][dense_ld] float32 scores, or null
    int64_t dense_ld;
    const int32_t* row_map;       // optional: product row r is stored at dense row row_map[r]
    // Multi-GPU (comm.cuh): the same rows also go into these replicas of the dense matrix on the other
    // GPUs (NVLink peer memory) -- the all-gather of a freshly solved / whitened shard is this epilogue.
    int n_dense_peer = 0;
    float* dense_peer[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // Survivors for the top-k (topk.cuh, "fused score -> mask -> top-k"): every cell of the sweep whose item is a
    // candidate of its user (bit set in the user's bitmap; null: every item) and whose score reaches the user's cut
    // is appended to the user's list -- the only trace the users x items scores leave in memory.
    const float* surv_cut = nullptr;          // [m] cut - 2E per user, or null: nothing is emitted
    const unsigned int* cand_bits = nullptr;  // [m][bits_ld]
    int64_t bits_ld = 0;
    unsigned int* surv_cnt = nullptr;         // [m], zeroed by the caller; counts past surv_cap mean overflow
    uint2* surv = nullptr;                    // [m][surv_cap]: (item id, float32 score bits)
    int surv_cap = 0;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: ~1 s of polling, then trap (a broken pipeline must not hang the device).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// rows are 128 B apart inside an 8-row group, groups are 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);       // start address, 16-byte units
    d |= (uint64_t)1 << 16;                           // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset
    d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
    return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): float32 accumulate, TF32 x TF32, K-major A and B.
__device__ __forceinline__ uint32_t umma_idesc_tf32(int mdim, int ndim) {
    uint32_t d = 0;
    d |= 1u << 4;                        // c_format = F32
    d |= 2u << 7;                        // a_format = TF32
    d |= 2u << 10;                       // b_format = TF32
    d |= (uint32_t)(ndim >> 3) << 17;    // N / 8
    d |= (uint32_t)(mdim >> 4) << 24;    // M / 16
    return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_load_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
tc_score_kernel(const __grid_constant__ CUtensorMap map_u_hi, const __grid_constant__ CUtensorMap map_u_lo,
                const __grid_constant__ CUtensorMap map_v_hi, const __grid_constant__ CUtensorMap map_v_lo, ScoreArgs p) {
    extern __shared__ uint8_t tc_smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B needs 1024 B
    uint64_t* bars = (uint64_t*)(smem + (size_t)kStages * kStageBytes);
    uint64_t* full_bar = bars;                       // [kStages]
    uint64_t* empty_bar = bars + kStages;            // [kStages]
    uint64_t* acc_full = bars + 2 * kStages;         // [2]
    uint64_t* acc_empty = bars + 2 * kStages + 2;    // [2]
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * kStages + 4);
    float* vnorm_s = (float*)(smem + (size_t)kStages * kStageBytes + 256);   // [kBN]

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int tiles_m = (int)((p.m + kBM - 1) / kBM), tiles_n = (int)((p.n + kBN - 1) / kBN);
    const int n_tiles = tiles_m * tiles_n;
    const int k_blocks = (p.k + kBK - 1) / kBK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], kEpilogueWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int tm = tile / tiles_n, tn = tile % tiles_n;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* st = smem + (size_t)stage * kStageBytes;
                    mbar_expect_tx(&full_bar[stage], kStageBytes);
                    tma_load_2d(st, &map_u_hi, &full_bar[stage], kb * kBK, tm * kBM);
                    tma_load_2d(st + kTileBytes, &map_u_lo, &full_bar[stage], kb * kBK, tm * kBM);
                    tma_load_2d(st + 2 * kTileBytes, &map_v_hi, &full_bar[stage], kb * kBK, tn * kBN);
                    tma_load_2d(st + 3 * kTileBytes, &map_v_lo, &full_bar[stage], kb * kBK, tn * kBN);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(kBM, kBN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(&acc_empty[acc], acc_phase ^ 1);          // epilogue drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_c = tmem_base + (uint32_t)(acc * kAccCols);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = smem_u32(smem + (size_t)stage * kStageBytes);
                    const uint64_t a_hi = umma_desc(st), a_lo = umma_desc(st + kTileBytes);
                    const uint64_t b_hi = umma_desc(st + 2 * kTileBytes), b_lo = umma_desc(st + 3 * kTileBytes);
#pragma unroll
                    for (int ks = 0; ks < kBK / 8; ++ks) {
                        const uint64_t adv = (uint64_t)((ks * 8 * 4) >> 4);   // 32 bytes per K=8 step
                        umma_tf32(tmem_c, a_hi + adv, b_hi + adv, idesc, (kb | ks) != 0);
                        umma_tf32(tmem_c, a_hi + adv, b_lo + adv, idesc, 1);
                        umma_tf32(tmem_c, a_lo + adv, b_hi + adv, idesc, 1);
                    }
                    umma_commit(&empty_bar[stage]);                  // smem stage free once these MMAs retire
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&acc_full[acc]);                         // accumulator complete
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        const int quad = warp % 4;                  // TMEM lane quadrant this warp may access
        const int half = (warp - 2) / 4;            // which 64 columns of the tile
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int tm = tile / tiles_n, tn = tile % tiles_n;
            const int64_t row = (int64_t)tm * kBM + quad * 32 + lane;
            const int64_t col0 = (int64_t)tn * kBN;
            const bool row_ok = row < p.m;
            // stage the tile's column norms (all epilogue warps take part)
            const int et = threadIdx.x - 64;         // 0..255
            asm volatile("bar.sync 1, 256;" ::: "memory");       // previous tile's readers are done
            if (et < kBN) vnorm_s[et] = (col0 + et < p.n && p.vnorm) ? p.vnorm[col0 + et] : 0.f;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // float32 test: s~ - thr_f against margin = c ||u|| ||v|| (+0.1 % for the float32 products
            // of the margin itself) + the rounding of the threshold to float32
            float thr_f = 0.f, thr_err = 0.f, un = 0.f;
            if (row_ok && p.counts) {
                const double thr = p.thresholds[row];
                thr_f = (float)thr;
                thr_err = fabsf(thr_f) * 1.1920929e-7f;
                un = p.unorm[row] * p.margin_coef * 1.001f;
            }
            // (fetched before the wait for the accumulator: the bitmap words of this warp's two 32-column chunks)
            float cut_f = 0.f;
            unsigned int cand_w[kBN / 64] = {0xffffffffu, 0xffffffffu};
            const bool emit = row_ok && p.surv_cut != nullptr;
            if (emit) {
                cut_f = p.surv_cut[row];
#pragma unroll
                for (int q = 0; q < kBN / 64; ++q) {
                    const int64_t cq = col0 + half * (kBN / 2) + 32 * q;
                    if (p.cand_bits) cand_w[q] = cq < p.n ? p.cand_bits[row * p.bits_ld + (cq >> 5)] : 0u;
                    if (cq + 32 > p.n) cand_w[q] &= (cq >= p.n) ? 0u : (0xffffffffu >> (32 - (int)(p.n - cq)));
                }
            }
            mbar_wait(&acc_full[acc], acc_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            unsigned int definite = 0;
#pragma unroll 1
            for (int c = half * (kBN / 2); c < (half + 1) * (kBN / 2); c += 32) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kAccCols + c);
                tmem_load_32x32(taddr, v);
                if (row_ok) {
                    if (p.dense) {
                        const int64_t drow = p.row_map ? (int64_t)p.row_map[row] : row;
                        const int64_t doff = drow * p.dense_ld + col0 + c;
                        for (int rep = 0; rep <= p.n_dense_peer; ++rep) {
                            float* dst = (rep == 0 ? p.dense : p.dense_peer[rep - 1]) + doff;
                            if (col0 + c + 32 <= p.dense_ld) {
#pragma unroll
                                for (int j = 0; j < 32; j += 4)
                                    *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                                        __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                            } else {
                                for (int j = 0; j < 32; ++j)
                                    if (col0 + c + j < p.dense_ld) dst[j] = __uint_as_float(v[j]);
                            }
                        }
                    }
                    if (emit) {
                        const unsigned int cand = cand_w[(c - half * (kBN / 2)) >> 5];
                        if (cand) {
                            // one compare + one predicated OR per cell; the (rare) appends run off the mask, with ONE
                            // reservation in the user's list per chunk (a returning atomic is a round trip to L2)
                            unsigned int pass = 0;
#pragma unroll
                            for (int j = 0; j < 32; ++j) pass |= (__uint_as_float(v[j]) >= cut_f) ? (1u << j) : 0u;
                            pass &= cand;
                            if (pass) {
                                unsigned int at = atomicAdd(p.surv_cnt + row, (unsigned int)__popc(pass));
                                uint2* list = p.surv + (size_t)row * p.surv_cap;
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    if (pass & (1u << j)) {
                                        if (at < (unsigned int)p.surv_cap) list[at] = make_uint2((unsigned int)(col0 + c + j), v[j]);
                                        ++at;
                                    }
                                }
                            }
                        }
                    }
                    if (p.counts) {
                        const bool full_chunk = col0 + c + 32 <= p.n;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (full_chunk || col0 + c + j < p.n) {
                                const float diff = __uint_as_float(v[j]) - thr_f;
                                const float margin = fmaf(un, vnorm_s[c + j], thr_err);
                                if (diff > margin) {
                                    ++definite;
                                } else if (diff >= -margin) {
                                    const unsigned int at = atomicAdd(p.border_count, 1u);
                                    if (at < p.border_capacity)
                                        p.border_cells[at] = make_uint2((unsigned)row, (unsigned)(col0 + c + j));
                                }
                            }
                        }
                    }
                }
            }
            if (row_ok && p.counts && definite) atomicAdd(p.counts + row, (unsigned long long)definite);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    }
}

// x = hi + lo (+ O(2^-22 |x|)), both representable in TF32 (cvt.rna keeps 10 mantissa bits).
template <typename T>
__global__ void split_tf32_kernel(const T* __restrict__ x, int64_t count, float* __restrict__ hi, float* __restrict__ lo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double v = (double)x[i];
    uint32_t h, l;
    const float vf = (float)v;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(vf));
    const float hf = __uint_as_float(h);
    const float rest = (float)(v - (double)hf);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(rest));
    hi[i] = hf;
    lo[i] = __uint_as_float(l);
}

// ||row||_2 rounded up to float32, one warp per row.
template <typename T>
__global__ void row_norm_kernel(const T* __restrict__ x, int64_t rows, int k, float* __restrict__ out) {
    const int lane = threadIdx.x % 32;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    if (r >= rows) return;
    double s = 0.0;
    for (int l = lane; l < k; l += 32) s += (double)x[r * k + l] * (double)x[r * k + l];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) out[r] = __double2float_ru(sqrt(s) * (1.0 + 1e-6));
}

// Exact float64 re-evaluation of the cells the tensor-core pass could not decide.
template <typename T>
__global__ void recheck_cells_kernel(const uint2* __restrict__ cells, const unsigned int* __restrict__ n_cells,
                                     unsigned int capacity, const T* __restrict__ user_vecs,
                                     const T* __restrict__ item_vecs, int k, const double* __restrict__ thresholds,
                                     unsigned long long* __restrict__ counts) {
    const int lane = threadIdx.x % 32;
    unsigned int total = *n_cells;
    if (total > capacity) total = capacity;
    for (unsigned int c = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32; c < total; c += gridDim.x * (blockDim.x / 32)) {
        const uint2 cell = cells[c];
        double s = 0.0;
        const T* u = user_vecs + (int64_t)cell.x * k;
        const T* v = item_vecs + (int64_t)cell.y * k;
        for (int l = lane; l < k; l += 32) s += (double)u[l] * (double)v[l];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (lane == 0 && s >= thresholds[cell.x]) atomicAdd(counts + cell.x, 1ull);
    }
}

}  // namespace tc
}  // namespace hyprec
