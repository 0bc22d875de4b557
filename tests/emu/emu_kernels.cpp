// TEST INFRASTRUCTURE: runs the SIMT kernels of hyprec_b200/csrc/*.cuh on the CPU through
// emu_cuda.h so tests/test_emu_kernels.py can check their indexing and synchronisation against
// the oracle without a GPU.  Built by tests/emu/build.py into tests/emu/libhyprec_emu.so.
#define HYPREC_EMU 1
#include "emu_cuda.h"

namespace emu {
thread_local emu_dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
thread_local Block* t_block = nullptr;
}  // namespace emu

#include "../../hyprec_b200/csrc/als_solve.cuh"
#include "../../hyprec_b200/csrc/gram.cuh"
#include "../../hyprec_b200/csrc/score.cuh"
#include "../../hyprec_b200/csrc/topk.cuh"
#include "../../hyprec_b200/csrc/comm.cuh"
#include "../../hyprec_b200/csrc/small_f64.cuh"

using namespace hyprec;

template <typename T>
static void gram_impl(const T* y, int64_t rows, int k, int nsplit, T* g) {
    const int ntile = (k + kGramTile - 1) / kGramTile, tiles = tri(ntile);
    int64_t per = (rows + nsplit - 1) / nsplit;
    per = (per + kGramRows - 1) / kGramRows * kGramRows;
    nsplit = (int)((rows + per - 1) / per);
    std::vector<T> partial((size_t)nsplit * k * k, T(0));
    emu::launch(dim3(tiles, nsplit), dim3(kGramThreads), 0,
                [&]() { gram_partial_kernel<T>(y, rows, k, per, partial.data()); });
    emu::launch(dim3((k * k + 255) / 256), dim3(256), 0,
                [&]() { gram_reduce_kernel<T>(partial.data(), nsplit, k, g); });
}

static double* g_cross_out = nullptr;     // emu_set_cross_out: where the row kernels put the RMSE cross term

template <typename T>
static int half_step_impl(const int64_t* indptr, const int32_t* indices, const T* values, const T* fixed,
                          int64_t n_fixed, T* latent, int64_t row_lo, int64_t row_hi, int k, T lambda,
                          const T* prior, int chunk_rows, int grid, int a_in_smem) {
    std::vector<T> g((size_t)k * k);
    gram_impl<T>(fixed, n_fixed, k, 3, g.data());
    const int nb = aug_blocks(k), n_tiles = tri(nb);
    std::vector<T> base((size_t)kTileElems * n_tiles);
    emu::launch(dim3((kTileElems * n_tiles + 255) / 256), dim3(256), 0,
                [&]() { build_base_tiles_kernel<T>(g.data(), k, lambda, base.data()); });
    SolveLayout lay = solve_layout(8, k, chunk_rows, a_in_smem != 0, sizeof(T));
    std::vector<T> scratch(a_in_smem ? 1 : (size_t)grid * kTileElems * n_tiles);
    int counter[2] = {0, 0};
    SolveArgs<T> a;
    a.indptr = indptr; a.indices = indices; a.values = values; a.fixed = fixed; a.latent = latent;
    a.base = base.data(); a.prior = prior; a.a_scratch = a_in_smem ? nullptr : scratch.data();
    a.work_counter = counter; a.status = counter + 1; a.lambda = lambda; a.k = k;
    a.row_lo = row_lo; a.row_hi = row_hi; a.chunk_rows = chunk_rows;
    a.row_list = nullptr; a.n_list = 0; a.max_dim = k; a.lowrank = 0; a.ymat = nullptr; a.ymat_row0 = 0; a.export_tiles = nullptr; a.prof = nullptr;
    a.cross_out = g_cross_out;
    if (a_in_smem) emu::launch(dim3(grid), dim3(kSolveThreads), lay.total, [&]() { als_solve_rows_kernel<T, 8>(a); });
    else emu::launch(dim3(grid), dim3(kSolveThreads), lay.total, [&]() { als_solve_rows_kernel<T, 8, false, true>(a); });
    return counter[1];
}

// C = A B^T through the sweep kernel (A: ra x k, B: rb x k, C: ra x rb float64)
static void gemm_nt(const double* a, int64_t ra, const double* b, int64_t rb, int k, double* c) {
    SweepArgs<double> s;
    s.user_vecs = a; s.item_vecs = b; s.k = k; s.user_lo = 0; s.user_hi = ra; s.n_items = rb;
    s.dense_out = c; s.dense_out_f32 = nullptr; s.row_list = nullptr; s.thresholds = nullptr; s.counts = nullptr;
    emu::launch(dim3((unsigned)((rb + kSweepTile - 1) / kSweepTile), (unsigned)((ra + kSweepTile - 1) / kSweepTile)),
                dim3(kSweepThreads), 0, [&]() { score_sweep_kernel<double, kSweepStore>(s); });
}

// Low-rank (Woodbury) half-step, the same pipeline hyprec_b200/csrc/api.cu runs on the device:
// Gram -> base tiles -> factor B (one dummy row, tiles exported) -> L^-1 -> Q = Y L^-T ->
// per-row deg x deg systems on Q (rows with deg <= dim_cap) -> X = Ymat L^-1; heavier rows go
// through the direct kernel.
static int half_step_lowrank_impl(const int64_t* indptr, const int32_t* indices, const double* values,
                                  const double* fixed, int64_t n_fixed, double* latent, int64_t n_rows, int k,
                                  double lambda, const double* prior, int dim_cap, int threads, int grid,
                                  int tiles_in_smem, int ts) {
    typedef double T;
    std::vector<T> g((size_t)k * k);
    gram_impl<T>(fixed, n_fixed, k, 3, g.data());
    const int nb = aug_blocks(k), n_tiles = tri(nb);
    std::vector<T> base((size_t)kTileElems * n_tiles), tiles((size_t)kTileElems * n_tiles);
    emu::launch(dim3((kTileElems * n_tiles + 255) / 256), dim3(256), 0,
                [&]() { build_base_tiles_kernel<T>(g.data(), k, lambda, base.data()); });
    int counter[2] = {0, 0};
    // 1. factor B: a single row without positives
    int64_t dummy_ptr[2] = {0, 0};
    std::vector<T> dummy_out(k);
    SolveArgs<T> a;
    a.indptr = dummy_ptr; a.indices = indices; a.values = values; a.fixed = fixed; a.latent = dummy_out.data();
    a.base = base.data(); a.prior = nullptr; a.a_scratch = nullptr; a.work_counter = counter; a.status = counter + 1;
    a.lambda = lambda; a.k = k; a.row_lo = 0; a.row_hi = 1; a.chunk_rows = 8; a.row_list = nullptr; a.n_list = 0;
    a.max_dim = k; a.lowrank = 0; a.ymat = nullptr; a.ymat_row0 = 0; a.export_tiles = tiles.data(); a.prof = nullptr;
    SolveLayout lay = solve_layout(8, k, 8, true, sizeof(T));
    emu::launch(dim3(1), dim3(kSolveThreads), lay.total, [&]() { als_solve_rows_kernel<T, 8>(a); });
    // 2. L^-1 and its transpose
    std::vector<T> linv((size_t)k * k, -777.0), linv_t((size_t)k * k, -777.0);
    const int nbk = (k + kTile - 1) / kTile;
    size_t inv_smem = sizeof(T) * ((size_t)nbk * kTileElems + kTileElems + (tiles_in_smem ? (size_t)kTileElems * n_tiles : 0));
    emu::launch(dim3(nbk), dim3(64), inv_smem,
                [&]() { tri_inverse_kernel<T>(tiles.data(), k, tiles_in_smem, linv.data(), linv_t.data(), k); });
    // 3. whitened factors Q = Y L^-T  (Q[i][a] = sum_b Y[i][b] Linv[a][b]) and prior rows p = lambda theta L^-T
    std::vector<T> q((size_t)n_fixed * k), pq;
    gemm_nt(fixed, n_fixed, linv.data(), k, k, q.data());
    if (prior) {
        pq.resize((size_t)n_rows * k);
        gemm_nt(prior, n_rows, linv.data(), k, k, pq.data());
    }
    // 4. rows by degree
    std::vector<int32_t> light, heavy;
    int max_deg = 1;
    for (int64_t r = 0; r < n_rows; ++r) {
        const int deg = (int)(indptr[r + 1] - indptr[r]);
        if (deg <= dim_cap) { light.push_back((int32_t)r); max_deg = std::max(max_deg, deg); }
        else heavy.push_back((int32_t)r);
    }
    std::vector<T> ymat((size_t)n_rows * k, 0.0);
    if (!light.empty()) {
        counter[0] = 0;
        SolveArgs<T> b = a;
        b.indptr = indptr; b.fixed = q.data(); b.prior = prior ? pq.data() : nullptr; b.latent = nullptr;
        b.row_list = light.data(); b.n_list = (int64_t)light.size(); b.max_dim = max_deg; b.lowrank = 1;
        b.ymat = ymat.data(); b.ymat_row0 = 0; b.export_tiles = nullptr; b.chunk_rows = 16;   // compact: list position i -> row i
        b.cross_out = g_cross_out;
        SolveLayout l2 = solve_layout(ts, max_deg, 16, true, sizeof(T));
        if (ts == 0) {
            // the warp-per-row kernel (small_f64.cuh): every light row must have <= 32 positives and no prior
            SmallArgs sa;
            sa.indptr = indptr; sa.indices = indices; sa.values = values; sa.fixed = q.data(); sa.ymat = ymat.data();
            sa.ymat_row0 = 0; sa.row_list = light.data(); sa.n_list = (int64_t)light.size(); sa.k = k;
            sa.work_counter = counter; sa.status = counter + 1; sa.cross_out = g_cross_out;
            emu::launch(dim3(grid), dim3(kSmallWarps * 32), small_smem_bytes(), [&]() { lowrank_small_kernel<double>(sa); });
        } else if (ts == 4) emu::launch(dim3(grid), dim3(threads), l2.total, [&]() { als_solve_rows_kernel<T, 4>(b); });
        else emu::launch(dim3(grid), dim3(threads), l2.total, [&]() { als_solve_rows_kernel<T, 8>(b); });
        // 5. X = Ymat L^-1  (X[i][c] = sum_a Ymat[i][a] LinvT[c][a]) -- light rows only are meaningful
        std::vector<T> x((size_t)n_rows * k);
        gemm_nt(ymat.data(), n_rows, linv_t.data(), k, k, x.data());
        for (size_t i = 0; i < light.size(); ++i) memcpy(latent + (size_t)light[i] * k, x.data() + i * k, sizeof(T) * k);
    }
    if (!heavy.empty()) {
        counter[0] = 0;
        SolveArgs<T> d = a;
        d.indptr = indptr; d.latent = latent; d.prior = prior; d.row_list = heavy.data(); d.n_list = (int64_t)heavy.size();
        d.export_tiles = nullptr; d.chunk_rows = 8;
        d.cross_out = g_cross_out;
        emu::launch(dim3(grid), dim3(kSolveThreads), lay.total, [&]() { als_solve_rows_kernel<T, 8>(d); });
    }
    return counter[1];
}

template <typename T>
static void eval_impl(const T* U, const T* V, int64_t m, int64_t n, int k, int64_t lo, int64_t hi,
                      const int64_t* cp, const int32_t* ci, int cap, const int64_t* fp, const int32_t* fi,
                      const double* fv, double* thr, int64_t* ones, int32_t* tidx, double* tval, double* thit,
                      int grid, double prefilter_coef = 0.0, bool spill = false, bool survivors = false,
                      unsigned int* fallbacks_out = nullptr, bool two_pass = false) {
    std::vector<double> part(2 * (size_t)k), colsum(k);
    const int64_t per = (n + 1) / 2;
    emu::launch(dim3(2), dim3(64), 0, [&]() { colsum_partial_kernel<T>(V, n, k, per, part.data()); });
    emu::launch(dim3((k + 255) / 256), dim3(256), 0, [&]() { colsum_reduce_kernel(part.data(), 2, k, colsum.data()); });
    emu::launch(dim3((unsigned)((m + 1) / 2)), dim3(64), 0,
                [&]() { row_threshold_kernel<T>(U, colsum.data(), k, m, n, thr); });
    const int64_t nu = hi - lo;
    if (ones) {
        std::vector<unsigned long long> counts(nu, 0);
        SweepArgs<T> a;
        a.user_vecs = U; a.item_vecs = V; a.k = k; a.user_lo = lo; a.user_hi = hi; a.n_items = n;
        a.dense_out = nullptr; a.dense_out_f32 = nullptr; a.row_list = nullptr; a.thresholds = thr; a.counts = counts.data();
        emu::launch(dim3((unsigned)((n + kSweepTile - 1) / kSweepTile), (unsigned)((nu + kSweepTile - 1) / kSweepTile)),
                    dim3(kSweepThreads), 0, [&]() { score_sweep_kernel<T, kSweepCount>(a); });
        for (int64_t u = 0; u < nu; ++u) ones[u] = (int64_t)counts[u];
    }
    int64_t max_cand = n;
    if (ci) {
        max_cand = 1;
        for (int64_t u = 0; u < nu; ++u) max_cand = std::max(max_cand, cp[u + 1] - cp[u]);
    }
    int sort_size = 2;
    while (sort_size < cap) sort_size <<= 1;
    // optional pre-filter: float32 scores as the tensor-core sweep would provide them (here: the
    // float64 product rounded to float32 and perturbed inside the error bound)
    std::vector<float> approx, unorm_v(nu, 0.f);
    float vmax[2] = {0.f, 0.f};
    const int64_t ld = n + 3;
    if (prefilter_coef > 0.0) {
        approx.assign((size_t)nu * ld, 0.f);
        for (int64_t j = 0; j < n; ++j) {
            double s = 0;
            for (int c = 0; c < k; ++c) s += (double)V[j * k + c] * (double)V[j * k + c];
            vmax[0] = std::max(vmax[0], (float)(std::sqrt(s) * 1.0001));
        }
        for (int64_t u = 0; u < nu; ++u) {
            double s = 0;
            for (int c = 0; c < k; ++c) s += (double)U[(lo + u) * k + c] * (double)U[(lo + u) * k + c];
            unorm_v[u] = (float)(std::sqrt(s) * 1.0001);
            for (int64_t j = 0; j < n; ++j) {
                double d = 0;
                for (int c = 0; c < k; ++c) d += (double)U[(lo + u) * k + c] * (double)V[j * k + c];
                const double wiggle = 0.9 * prefilter_coef * unorm_v[u] * vmax[0] * (((u * 131 + j * 7) % 21) - 10) / 10.0;
                approx[(size_t)u * ld + j] = (float)(d + wiggle);
            }
        }
    }
    // survivors mode: what api.cu does without the dense matrix -- candidate bitmap, sampled columns, per-user cut
    // (sample_cut_kernel), and the cells the sweep's epilogue would emit (candidate cells with s~ >= cut - 2E),
    // here taken from the perturbed `approx` scores in a scrambled order
    std::vector<unsigned int> bits, surv_cnt;
    std::vector<int32_t> distinct;
    std::vector<uint2> surv;
    std::vector<float> cut, cut_emit, sample;
    if (survivors) {
        const int64_t words = (n + 31) / 32;
        if (ci) {
            bits.assign((size_t)nu * words, 0u);
            distinct.assign(nu, 0);
            emu::launch(dim3(3), dim3(64), 0, [&]() { cand_bitmap_kernel(cp, ci, nu, words, bits.data(), distinct.data()); });
        }
        const int target = eval_cut_target(cap);
        const int64_t n_sample = eval_sample_columns(n, target, 16);
        sample.assign((size_t)nu * n_sample, 0.f);
        for (int64_t u = 0; u < nu; ++u)
            for (int64_t j = 0; j < n_sample; ++j) sample[(size_t)u * n_sample + j] = approx[(size_t)u * ld + sample_item(j, n_sample, n)];
        cut.assign(nu, 0.f);
        cut_emit.assign(nu, 0.f);
        SampleCutArgs sc;
        sc.scores = sample.data(); sc.scores_ld = n_sample; sc.n_sample = n_sample; sc.n_items = n;
        sc.cand_bits = ci ? bits.data() : nullptr; sc.bits_ld = words; sc.cand_indptr = ci ? cp : nullptr; sc.n_users = nu;
        sc.capacity = cap; sc.target = target; sc.unorm = unorm_v.data(); sc.vmax = vmax; sc.err_coef = (float)prefilter_coef;
        sc.cut = cut.data(); sc.cut_emit = cut_emit.data();
        // (room for about half of the sampled columns: users with longer lists take the re-reading path)
        sc.stage_cap = (int)std::max<int64_t>(8, n_sample / 2);
        emu::launch(dim3(2), dim3(kTopkThreads), sizeof(unsigned) * (size_t)sc.stage_cap, [&]() { sample_cut_kernel(sc); });
        surv.assign((size_t)nu * kSurvCap, make_uint2(0u, 0u));
        surv_cnt.assign(nu, 0u);
        for (int64_t u = 0; u < nu; ++u)
            for (int64_t jj = 0; jj < n; ++jj) {
                const int64_t j = (jj * 7919 + 13 * u) % n;          // any order (7919 is prime: a permutation for n != 7919)
                const bool cand = !ci || ((bits[(size_t)u * words + (j >> 5)] >> (j & 31)) & 1u);
                if (cand && approx[(size_t)u * ld + j] >= cut_emit[u]) {
                    const unsigned int at = surv_cnt[u]++;
                    if (at < (unsigned int)kSurvCap) surv[(size_t)u * kSurvCap + at] = make_uint2((unsigned)j, __float_as_uint(approx[(size_t)u * ld + j]));
                }
            }
    }
    TopkLayout lay = topk_layout((int)max_cand, k, sort_size, prefilter_coef > 0.0, survivors);
    TopkArgs<T> a;
    a.user_vecs = U; a.item_vecs = V; a.k = k; a.n_items = n; a.user_lo = lo; a.user_hi = hi;
    a.cand_indptr = ci ? cp : nullptr; a.cand_ids = ci; a.thresholds = thr;
    a.full_indptr = fp; a.full_indices = fi; a.full_values = fv;
    a.capacity = cap; a.sort_size = sort_size; a.max_cand = (int)max_cand;
    a.topk_idx = tidx; a.topk_val = tval; a.topk_hit = thit;
    a.approx = prefilter_coef > 0.0 ? approx.data() : nullptr; a.approx_ld = ld; a.unorm = unorm_v.data();
    a.err_coef = (float)prefilter_coef; a.vmax = vmax; a.fallbacks = reinterpret_cast<unsigned int*>(vmax + 1);
    a.spill = nullptr; a.spill_stride = 0;
    if (survivors) {
        a.approx = nullptr;
        a.surv = surv.data(); a.surv_cnt = surv_cnt.data(); a.surv_cap = kSurvCap; a.cut = cut.data();
        a.cand_distinct = ci ? distinct.data() : nullptr;
    }
    std::vector<unsigned char> redo(nu, 0);
    if (survivors && two_pass) {
        // as api.cu launches it: pass 1 on the emitted cells alone (arrays of kSurvCap entries), pass 2 for the flagged users
        TopkArgs<T> f = a;
        f.pass = 1; f.redo = redo.data(); f.max_cand = kSurvCap;
        const TopkLayout flay = topk_layout(kSurvCap, k, sort_size, true, true);
        emu::launch(dim3(grid), dim3(kTopkThreads), flay.total, [&]() { eval_topk_kernel<T>(f); });
        a.pass = 2; a.redo = redo.data();
    }
    if (!spill) {
        emu::launch(dim3(grid), dim3(kTopkThreads), lay.total, [&]() { eval_topk_kernel<T>(a); });
        if (fallbacks_out) memcpy(fallbacks_out, vmax + 1, sizeof(unsigned int));
        return;
    }
    // per-candidate arrays in a per-CTA "global" scratch, as api.cu sets it up for lists beyond shared memory
    const size_t stride = (lay.off_u + 255) & ~(size_t)255;
    std::vector<unsigned char> scratch(stride * (size_t)grid, 0xcd);
    a.spill = scratch.data();
    a.spill_stride = stride;
    lay = topk_layout(0, k, sort_size, prefilter_coef > 0.0, survivors);
    emu::launch(dim3(grid), dim3(kTopkThreads), lay.total, [&]() { eval_topk_kernel<T, true>(a); });
    if (fallbacks_out) memcpy(fallbacks_out, vmax + 1, sizeof(unsigned int));
}

extern "C" {

void emu_set_cross_out(double* p) { g_cross_out = p; }

// the warp-per-row kernel on float32 factors (float64 arithmetic inside): rows `list` of the CSR against whitened
// factors q, compact output ymat[i] for list position i, cross[row] = (|r|^2 - |z|^2) / 0.9
int emu_small_rows_f32(const int64_t* indptr, const int32_t* indices, const float* values, const float* q, int k,
                       const int32_t* list, int64_t n_list, float* ymat, double* cross, int grid) {
    int counter[2] = {0, 0};
    SmallArgsT<float> sa;
    sa.indptr = indptr; sa.indices = indices; sa.values = values; sa.fixed = q; sa.ymat = ymat; sa.ymat_row0 = 0;
    sa.row_list = list; sa.n_list = n_list; sa.k = k; sa.work_counter = counter; sa.status = counter + 1; sa.cross_out = cross;
    emu::launch(dim3(grid), dim3(kSmallWarps * 32), small_smem_bytes(), [&]() { lowrank_small_kernel<float>(sa); });
    return counter[1];
}

void emu_gram_f64(const double* y, int64_t rows, int k, int nsplit, double* g) { gram_impl<double>(y, rows, k, nsplit, g); }
void emu_gram_f32(const float* y, int64_t rows, int k, int nsplit, float* g) { gram_impl<float>(y, rows, k, nsplit, g); }

int emu_half_step_f64(const int64_t* indptr, const int32_t* indices, const double* values, const double* fixed,
                      int64_t n_fixed, double* latent, int64_t row_lo, int64_t row_hi, int k, double lambda,
                      const double* prior, int chunk_rows, int grid, int a_in_smem) {
    return half_step_impl<double>(indptr, indices, values, fixed, n_fixed, latent, row_lo, row_hi, k, lambda, prior,
                                  chunk_rows, grid, a_in_smem);
}
int emu_half_step_f32(const int64_t* indptr, const int32_t* indices, const float* values, const float* fixed,
                      int64_t n_fixed, float* latent, int64_t row_lo, int64_t row_hi, int k, float lambda,
                      const float* prior, int chunk_rows, int grid, int a_in_smem) {
    return half_step_impl<float>(indptr, indices, values, fixed, n_fixed, latent, row_lo, row_hi, k, lambda, prior,
                                 chunk_rows, grid, a_in_smem);
}

int emu_half_step_lowrank_f64(const int64_t* indptr, const int32_t* indices, const double* values, const double* fixed,
                              int64_t n_fixed, double* latent, int64_t n_rows, int k, double lambda, const double* prior,
                              int dim_cap, int threads, int grid, int tiles_in_smem, int ts) {
    return half_step_lowrank_impl(indptr, indices, values, fixed, n_fixed, latent, n_rows, k, lambda, prior, dim_cap,
                                  threads, grid, tiles_in_smem, ts);
}

void emu_eval_f64(const double* U, const double* V, int64_t m, int64_t n, int k, int64_t lo, int64_t hi,
                  const int64_t* cp, const int32_t* ci, int cap, const int64_t* fp, const int32_t* fi,
                  const double* fv, double* thr, int64_t* ones, int32_t* tidx, double* tval, double* thit, int grid) {
    eval_impl<double>(U, V, m, n, k, lo, hi, cp, ci, cap, fp, fi, fv, thr, ones, tidx, tval, thit, grid);
}

void emu_eval_prefilter_f64(const double* U, const double* V, int64_t m, int64_t n, int k, int64_t lo, int64_t hi,
                            const int64_t* cp, const int32_t* ci, int cap, const int64_t* fp, const int32_t* fi,
                            const double* fv, double* thr, int64_t* ones, int32_t* tidx, double* tval, double* thit,
                            int grid, double prefilter_coef) {
    eval_impl<double>(U, V, m, n, k, lo, hi, cp, ci, cap, fp, fi, fv, thr, ones, tidx, tval, thit, grid, prefilter_coef);
}

void emu_eval_spill_f64(const double* U, const double* V, int64_t m, int64_t n, int k, int64_t lo, int64_t hi,
                        const int64_t* cp, const int32_t* ci, int cap, const int64_t* fp, const int32_t* fi,
                        const double* fv, double* thr, int64_t* ones, int32_t* tidx, double* tval, double* thit,
                        int grid, double prefilter_coef) {
    eval_impl<double>(U, V, m, n, k, lo, hi, cp, ci, cap, fp, fi, fv, thr, ones, tidx, tval, thit, grid, prefilter_coef,
                      true);
}

// the top-k fed by sweep survivors instead of a dense score matrix (spill != 0: per-candidate arrays in global scratch)
void emu_eval_survivors_f64(const double* U, const double* V, int64_t m, int64_t n, int k, int64_t lo, int64_t hi,
                            const int64_t* cp, const int32_t* ci, int cap, const int64_t* fp, const int32_t* fi,
                            const double* fv, double* thr, int64_t* ones, int32_t* tidx, double* tval, double* thit,
                            int grid, double prefilter_coef, int spill, unsigned int* fallbacks) {
    // spill: bit 0 = per-candidate arrays in a global scratch, bit 1 = two launches (pass 1 / pass 2)
    eval_impl<double>(U, V, m, n, k, lo, hi, cp, ci, cap, fp, fi, fv, thr, ones, tidx, tval, thit, grid, prefilter_coef,
                      (spill & 1) != 0, true, fallbacks, (spill & 2) != 0);
}

// One rank's side of `rounds` exchange rounds over host-memory "arenas" (csrc/comm.cuh): write own rows of a
// replicated matrix, push them into every replica, push a mailbox slot, signal / wait, reduce the slots in rank
// order, barrier again (nobody overwrites a slot another rank still reads).  Called from one host thread per rank.
// Returns the status word (2: a barrier timed out).
int emu_comm_rounds(void* const* bases, int world, int rank, uint64_t off_matrix, uint64_t off_slots, uint64_t slot_doubles,
                    uint64_t off_flags, int k, const int64_t* bounds, int rounds, double* sums_out, uint64_t timeout_ns) {
    PeerList peers;
    memset(&peers, 0, sizeof(peers));
    for (int r = 0; r < world; ++r) peers.base[r] = bases[r];
    peers.world = world;
    peers.rank = rank;
    int status = 0;
    unsigned long long seq = 0;
    double* own = reinterpret_cast<double*>(static_cast<unsigned char*>(bases[rank]) + off_matrix);
    std::vector<double> slot(slot_doubles);
    auto barrier = [&]() {
        ++seq;
        emu::launch(dim3(1), dim3(32), 0, [&]() { comm_signal_kernel(peers, off_flags, seq); });
        emu::launch(dim3(1), dim3(32), 0, [&]() { comm_wait_kernel(peers, off_flags, seq, timeout_ns, &status); });
    };
    for (int it = 0; it < rounds; ++it) {
        const int64_t lo = bounds[rank], hi = bounds[rank + 1];
        for (int64_t r = lo; r < hi; ++r)
            for (int c = 0; c < k; ++c) own[r * k + c] = 1000.0 * it + 10.0 * r + c + 0.5 * rank;
        emu::launch(dim3(2), dim3(64), 0, [&]() { push_rows_kernel<double>(peers, off_matrix, k, lo, hi - lo, nullptr); });
        for (uint64_t i = 0; i < slot_doubles; ++i) slot[i] = (rank + 1) * (double)(i + 1) + it;
        emu::launch(dim3(2), dim3(64), 0, [&]() { push_slot_kernel(peers, off_slots, slot_doubles, slot.data(), (int)slot_doubles); });
        barrier();
        const double* slots = reinterpret_cast<const double*>(static_cast<unsigned char*>(bases[rank]) + off_slots);
        emu::launch(dim3((unsigned)((slot_doubles + 63) / 64)), dim3(64), 0, [&]() {
            reduce_slots_kernel<double>(slots, slot_doubles, world, (int)slot_doubles, sums_out + (size_t)it * slot_doubles, nullptr);
        });
        barrier();
    }
    return status;
}

// device-generated candidate lists: fill == 0 writes counts[nu]; fill == 1 writes ids given the row pointers
void emu_hashed_candidates(const int64_t* fp, const int32_t* fi, const int64_t* tp, const int32_t* ti, int64_t lo,
                           int64_t hi, int64_t n, int fold, int n_folds, unsigned long long seed, int32_t* counts,
                           const int64_t* out_indptr, int32_t* out_ids, int fill, int grid) {
    HashedCandArgs a;
    a.full_indptr = fp; a.full_indices = fi; a.test_indptr = tp; a.test_indices = ti;
    a.user_lo = lo; a.user_hi = hi; a.n_items = n; a.fold = fold; a.n_folds = n_folds; a.seed = seed;
    a.counts = counts; a.out_indptr = out_indptr; a.out_ids = out_ids;
    if (fill) emu::launch(dim3(grid), dim3(kTopkThreads), 0, [&]() { hashed_candidates_kernel<true>(a); });
    else emu::launch(dim3(grid), dim3(kTopkThreads), 0, [&]() { hashed_candidates_kernel<false>(a); });
}

void emu_predict_f64(const double* U, const double* V, int64_t m, int64_t n, int k, double* out) {
    SweepArgs<double> a;
    a.user_vecs = U; a.item_vecs = V; a.k = k; a.user_lo = 0; a.user_hi = m; a.n_items = n;
    a.dense_out = out; a.dense_out_f32 = nullptr; a.row_list = nullptr; a.thresholds = nullptr; a.counts = nullptr;
    emu::launch(dim3((unsigned)((n + kSweepTile - 1) / kSweepTile), (unsigned)((m + kSweepTile - 1) / kSweepTile)),
                dim3(kSweepThreads), 0, [&]() { score_sweep_kernel<double, kSweepStore>(a); });
}

// sum_nz r (u.v), sum r^2, <U^T U, V^T V>
void emu_rmse_terms_f64(const int64_t* indptr, const int32_t* indices, const double* values, const double* U,
                        const double* V, int64_t m, int64_t n, int k, double* out3) {
    std::vector<double> gu((size_t)k * k), gv((size_t)k * k), s1(m), s2(m);
    gram_impl<double>(U, m, k, 2, gu.data());
    gram_impl<double>(V, n, k, 2, gv.data());
    emu::launch(dim3(1), dim3(256), 0, [&]() { frob_inner_kernel<double>(gu.data(), gv.data(), k * k, out3); });
    emu::launch(dim3((unsigned)((m + 1) / 2)), dim3(64), 0,
                [&]() { sddmm_row_kernel<double>(indptr, indices, values, U, V, k, m, s1.data(), s2.data()); });
    emu::launch(dim3(1), dim3(1024), 0, [&]() { reduce_sum_kernel(s1.data(), m, out3 + 1); });
    emu::launch(dim3(1), dim3(1024), 0, [&]() { reduce_sum_kernel(s2.data(), m, out3 + 2); });
}

void emu_nz_flags_f64(const int64_t* indptr, const int32_t* indices, const double* U, const double* V, int k,
                      const double* thr, int64_t lo, int64_t hi, uint8_t* flags) {
    const int64_t nnz = indptr[hi] - indptr[lo];          // one warp per positive, two warps per block
    if (nnz <= 0) return;
    emu::launch(dim3((unsigned)((nnz + 1) / 2)), dim3(64), 0,
                [&]() { nz_flags_kernel<double>(indptr, indices, U, V, k, thr, lo, hi, flags); });
}

}  // extern "C"
