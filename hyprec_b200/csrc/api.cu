// C ABI of libhyprec_b200.so (declared in include/hyprec_b200.h): owns device memory, launches
// the kernels of gram.cuh / als_solve.cuh / score.cuh / topk.cuh on one stream, reports errors
// by code + message.  No CPU fallback: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <chrono>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/hyprec_b200.h"
#include "als_solve.cuh"
#include "common.cuh"
#include "comm.cuh"
#include "gram.cuh"
#include "host_split.cuh"
#include "score.cuh"
#include "small_f64.cuh"
#include "tc_score.cuh"
#include "tc_chol.cuh"
#include "tc_chol_multi.cuh"
#include "topk.cuh"

namespace {

thread_local std::string g_error;

int fail(int code, const std::string& msg) {
    g_error = msg;
    return code;
}

#define CU_TRY(expr)                                                                              \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(e_ == cudaErrorMemoryAllocation ? HYPREC_E_NOMEM : HYPREC_E_CUDA,         \
                        std::string(#expr) + ": " + cudaGetErrorString(e_));                      \
    } while (0)

#define HY_TRY(expr)            \
    do {                        \
        int rc_ = (expr);       \
        if (rc_ != HYPREC_OK) return rc_; \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of the (device, kernel) pair for the
// whole process, not of a handle: several handles of different n_factors (a hyper-parameter sweep)
// launch the same instantiations with different sizes, possibly from different host threads.  The
// opted-in size therefore lives in one process-wide table and only ever grows.
int ensure_dynamic_smem(const void* func, int bytes) {
    static std::mutex guard;
    static std::map<std::pair<int, const void*>, int> opted;
    int device = 0;
    CU_TRY(cudaGetDevice(&device));
    std::lock_guard<std::mutex> lock(guard);
    int& have = opted[std::make_pair(device, func)];
    if (bytes > have) {
        CU_TRY(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        have = bytes;
    }
    return HYPREC_OK;
}

// Device buffer that only grows.
struct DevBuf {
    void* ptr = nullptr;
    size_t bytes = 0;
    bool external = false;     // a window of the exchange arena (multi-GPU): never freed or regrown here
    int reserve(size_t want) {
        if (want <= bytes) return HYPREC_OK;
        if (external)
            return fail(HYPREC_E_INVALID, "buffer belongs to the multi-GPU exchange arena and cannot grow: hyprec_comm_detach first");
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
        CU_TRY(cudaMalloc(&ptr, want));
        bytes = want;
        return HYPREC_OK;
    }
    void release() {
        if (ptr && !external) cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
        external = false;
    }
    void adopt(void* p, size_t n) {
        release();
        ptr = p;
        bytes = n;
        external = true;
    }
    template <typename U>
    U* as() const { return reinterpret_cast<U*>(ptr); }
};

struct Csr {
    int64_t rows = 0, nnz = 0;
    DevBuf indptr, indices, values;   // int64, int32, T
    std::vector<int64_t> h_indptr;    // host copy of the row pointers (sharding, flag offsets)
    void release() { indptr.release(); indices.release(); values.release(); }
};

constexpr int kNumTimers = 9;

struct EngineBase {
    virtual ~EngineBase() {}
    virtual int set_train(int64_t m, int64_t n, const int64_t* indptr, const int32_t* indices, const double* values) = 0;
    virtual int set_eval(const int64_t* fp, const int32_t* fi, const double* fv, const int64_t* tp,
                         const int32_t* ti, const double* tv) = 0;
    virtual int set_factors(const double* u, const double* v, int k) = 0;
    virtual int get_factors(double* u, double* v) = 0;
    virtual int set_prior(const double* theta) = 0;
    virtual int half_step(int side, double lambda) = 0;
    virtual int rmse(double* out) = 0;
    virtual int rmse_terms(double* out3, bool shard_only) = 0;
    virtual int predict_dense(double* out) = 0;
    virtual int eval_topk(int64_t lo, int64_t hi, const int64_t* cp, const int32_t* ci, int cap, double* thr,
                          int64_t* ones, int32_t* tidx, double* tval, double* thit, uint8_t* thit8, uint8_t* trf,
                          uint8_t* tef) = 0;
    virtual int device_factors(void** u, void** v, int* elem) = 0;
    virtual int set_candidates(int64_t lo, int64_t hi, const int64_t* cp, const int32_t* ci) = 0;
    virtual int set_hashed_candidates(int fold, int n_folds, unsigned long long seed) = 0;
    virtual int score_dense_f32(float* out) = 0;
    virtual int solve_phase_cycles(long long* out16, int reset) = 0;
    virtual int comm_export(int rank, int world, void* blob, size_t blob_bytes) = 0;
    virtual int comm_attach(const void* blobs, size_t blob_bytes, const int64_t* user_bounds, const int64_t* item_bounds) = 0;
    virtual int comm_detach() = 0;
    virtual int comm_attach_ptrs(void* const* bases, const int* devices, const int64_t* user_bounds, const int64_t* item_bounds) = 0;
    virtual int comm_arena(void** base, size_t* bytes) = 0;
    virtual int init_factors_uniform(int k, unsigned long long seed) = 0;
    virtual int get_factor_rows(int side, const int64_t* rows, int64_t n_rows, double* out) = 0;
    virtual int comm_stats(double* out8) = 0;
    virtual bool comm_active() const = 0;
    virtual int eval_stats(double* out8) = 0;
};

// What a rank publishes to its peers (hyprec_comm_export -> hyprec_comm_attach).
struct CommBlob {
    unsigned long long magic;
    cudaIpcMemHandle_t handle;
    long long m, n, k, elem, world, rank, arena_bytes, pid;
};
constexpr unsigned long long kCommMagic = 0x4859505243303031ull;   // "HYPRC001"

}  // namespace

struct hyprec_ctx {
    int device = 0;
    int precision = 0;
    cudaStream_t stream = nullptr;
    int num_sms = 0;
    int max_smem = 0;
    int64_t launches = 0;
    bool lowrank_enabled = true;   // HYPREC_LOWRANK=0 forces the direct k x k path
    int lowrank_cap = 0;           // HYPREC_LOWRANK_CAP=<deg>: rows with more positives go to the direct path (0: largest bucket that fits)
    bool tc_enabled = true;        // HYPREC_TC=0 keeps the score sweep on the SIMT kernel
    bool rmse_fused_enabled = true; // HYPREC_RMSE_FUSED=0: the RMSE cross term always comes from a pass over the positives
    bool tc_multi_enabled = true;  // HYPREC_TC_MULTI=0: systems of <= 128 rows stay on the shared-memory tile kernel
    bool tc_chol_enabled = true;   // HYPREC_TC_CHOL=0: float32 systems factorise in shared-memory tiles (als_solve.cuh)
    bool tc_solve_enabled = true;  // HYPREC_TC_SOLVE=0: keep the row systems' rank-k accumulation on the SIMT cores
    bool split_enabled = true;     // HYPREC_SPLIT=0: rows with tens of thousands of positives stay on one CTA
    bool small_f64_enabled = true; // HYPREC_SMALL_F64=0: float64 rows of <= 32 positives stay on the tile kernel
    // Multi-GPU: how the solved / whitened rows reach the other replicas.  Default: one copy kernel per range over
    // peer memory (push_rows_kernel: 512-byte warp stores).  HYPREC_PUSH=fused stores them from the GEMM epilogues
    // instead -- one kernel fewer, but the tensor-memory epilogue owns one ROW per lane, so its peer stores are
    // 16-byte pieces 1 KB apart: measured 111 ms against 88 ms per epoch on 8 B200s at the full scaled size.
    bool fused_push = false;
    // Low-rank degree buckets: threads per CTA and tile edge (4 or 8) of the shared-memory tile kernel, per bucket
    // (<= 31 / 63 / 127 / 191 / 255 / 319 unknowns).  HYPREC_BUCKET_THREADS / HYPREC_BUCKET_TILE = six comma-separated
    // numbers override them (A/B tests, tools/ab_epoch.py).
    int bucket_threads[6] = {64, 128, 128, 256, 256, 256};
    int bucket_tile[6] = {4, 4, 8, 8, 8, 8};
    int prof_slot = 0;             // HYPREC_SOLVE_PROF_SLOT: which launch class (0 direct, 1..6 buckets, 7 heavy)
    bool prof_enabled = false;     // HYPREC_SOLVE_PROF=1: per-phase clock64() totals of the direct solve
    void* encode_tiled = nullptr;  // cuTensorMapEncodeTiled, resolved through the runtime (no libcuda link)
    int64_t user_lo = 0, user_hi = -1, item_lo = 0, item_hi = -1;   // shard; hi < 0 => all
    // kernel timing: (begin, end) event pairs taken from a pool, harvested into per-class
    // totals after the stream has been synchronised
    std::vector<cudaEvent_t> ev_pool;
    struct Pending { cudaEvent_t b, e; int which; };
    std::vector<Pending> pending;
    double ms_total[kNumTimers] = {0};
    double ms_last[kNumTimers] = {0};
    int64_t ms_calls[kNumTimers] = {0};
    static constexpr int kAux = 7;
    cudaStream_t aux[kAux] = {};             // concurrent row-kernel launches
    cudaEvent_t ev_fork = nullptr, ev_base = nullptr, ev_join[kAux] = {};
    cudaEvent_t marks[8] = {};               // hyprec_mark / hyprec_elapsed_ms
    void* flush_buf = nullptr;
    size_t flush_bytes = 0;
    EngineBase* engine = nullptr;

    cudaEvent_t take_event() {
        if (ev_pool.empty()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            return e;
        }
        cudaEvent_t e = ev_pool.back();
        ev_pool.pop_back();
        return e;
    }
    void harvest() {   // call only when the stream is idle
        bool seen[kNumTimers] = {false};
        for (const Pending& p : pending) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, p.b, p.e) == cudaSuccess) {
                ms_total[p.which] += ms;
                ms_calls[p.which] += 1;
                ms_last[p.which] = seen[p.which] ? ms_last[p.which] + ms : ms;
                seen[p.which] = true;
            }
            ev_pool.push_back(p.b);
            ev_pool.push_back(p.e);
        }
        pending.clear();
    }
};

namespace {

struct Timer {
    hyprec_ctx* h;
    hyprec_ctx::Pending p;
    Timer(hyprec_ctx* ctx, int w) : h(ctx) {
        p.which = w;
        p.b = h->take_event();
        p.e = h->take_event();
        cudaEventRecord(p.b, h->stream);
    }
    ~Timer() {
        cudaEventRecord(p.e, h->stream);
        h->pending.push_back(p);
    }
};

template <typename T>
struct Engine : EngineBase {
    hyprec_ctx* h;
    int64_t m = 0, n = 0;
    int k = 0;
    Csr csr, csc, full, test;
    DevBuf U, V, prior;
    bool has_prior = false;
    bool gram_ok[2] = {false, false};   // [0] = U^T U, [1] = V^T V match the current factors
    // RMSE cross term sum_nz r (u . v) as a by-product of the item half-step (tensor-memory kernels, no
    // prior): cross[i] per item row; valid until anything touches the factors again.
    DevBuf cross;
    bool cross_active = false;          // the launches of the running half-step write cross[]
    bool cross_valid = false;
    int64_t cross_lo = 0, cross_hi = 0;
    DevBuf gram[2], gram_partial, base, colsum, colsum_partial, thresholds;
    DevBuf row_s1, row_s2, scalars, counter;
    DevBuf tiles, linv, linv_t, qmat, pq, ymat, dummy_ptr, dummy_out, prof_buf, bdense, lfac, blk_tmp, blk_base;
    DevBuf fixed_d, gram_partial_d, gram_dbuf, linv_f, linv_t_f, stage_d;
    DevBuf a_scratch_slot[8];
    static constexpr int kStatusSlot = 15;
    static constexpr int kHeavyStream = 6;   // aux streams 0..5: degree buckets, 6: rows above the cap
    bool hits_u8_exact = true;     // every rating an integer in 0..255: hit values fit one byte (hyprec_eval_topk_u8)
    DevBuf out_hit8;
    DevBuf cand_ptr, cand_ids, out_idx, out_val, out_hit, counts, flags, dense_blk, topk_spill, rows_idx, rows_out;

    explicit Engine(hyprec_ctx* ctx) : h(ctx) {}
    ~Engine() override {
        DevBuf* all[] = {&U, &V, &prior, &gram[0], &gram[1], &gram_partial, &base, &colsum, &colsum_partial,
                         &thresholds, &row_s1, &row_s2, &scalars, &counter, &cand_ptr,
                         &cand_ids, &out_idx, &out_val, &out_hit, &out_hit8, &counts, &flags, &dense_blk, &tiles, &linv,
                         &linv_t, &qmat, &pq, &ymat, &dummy_ptr, &dummy_out, &prof_buf, &bdense, &lfac, &blk_tmp, &blk_base, &fixed_d, &gram_partial_d, &gram_dbuf, &linv_f, &linv_t_f, &stage_d, &w_hi, &w_lo, &l_hi, &l_lo, &cross, &plan[0].lists, &plan[1].lists,
                         &a_scratch_slot[0], &a_scratch_slot[1], &a_scratch_slot[2], &a_scratch_slot[3],
                         &a_scratch_slot[4], &a_scratch_slot[5], &a_scratch_slot[6], &a_scratch_slot[7],
                         &u_hi, &u_lo, &v_hi, &v_lo, &unorm, &vnorm, &border_cells, &border_count, &dense_f32, &vmax_buf, &topk_spill, &hashed_counts};
        if (cm.exported) {
            if (cm.on && cm.ipc)
                for (int r = 0; r < cm.world; ++r)
                    if (r != cm.rank && cm.peers.base[r]) cudaIpcCloseMemHandle(cm.peers.base[r]);
            U.release();
            V.release();
            cudaFree(cm.arena);
        }
        for (DevBuf* b : all) b->release();
        DevBuf* more[] = {&gram64[0], &gram64[1], &slot_local, &comm_status, &split_partial, &split_pb, &split_gram, &split_b,
                          &plan[0].seg_lo, &plan[0].seg_len, &plan[0].seg_ptr, &plan[0].split_lo,
                          &plan[1].seg_lo, &plan[1].seg_len, &plan[1].seg_ptr, &plan[1].split_lo, &rows_idx, &rows_out};
        for (DevBuf* b : more) b->release();
        csr.release(); csc.release(); full.release(); test.release();
    }

    // ---------------------------------------------------------------- uploads
    int upload_csr(Csr& dst, int64_t rows, int64_t cols, const int64_t* indptr, const int32_t* indices,
                   const double* values, const char* what) {
        if (!indptr || rows < 0) return fail(HYPREC_E_INVALID, std::string(what) + ": null indptr");
        if (indptr[0] != 0) return fail(HYPREC_E_INVALID, std::string(what) + ": indptr[0] != 0");
        const int64_t nnz = indptr[rows];
        if (nnz > 0 && !indices) return fail(HYPREC_E_INVALID, std::string(what) + ": null indices");
        for (int64_t r = 0; r < rows; ++r) {
            if (indptr[r + 1] < indptr[r]) return fail(HYPREC_E_INVALID, std::string(what) + ": indptr not monotone");
            for (int64_t e = indptr[r]; e < indptr[r + 1]; ++e) {
                if (indices[e] < 0 || indices[e] >= cols)
                    return fail(HYPREC_E_INVALID, std::string(what) + ": column id out of range");
                if (e > indptr[r] && indices[e] <= indices[e - 1])
                    return fail(HYPREC_E_INVALID, std::string(what) + ": column ids must be strictly increasing in a row");
            }
        }
        dst.rows = rows;
        dst.nnz = nnz;
        dst.h_indptr.assign(indptr, indptr + rows + 1);
        HY_TRY(dst.indptr.reserve(sizeof(int64_t) * (rows + 1)));
        HY_TRY(dst.indices.reserve(sizeof(int32_t) * std::max<int64_t>(nnz, 1)));
        HY_TRY(dst.values.reserve(sizeof(T) * std::max<int64_t>(nnz, 1)));
        CU_TRY(cudaMemcpyAsync(dst.indptr.ptr, indptr, sizeof(int64_t) * (rows + 1), cudaMemcpyHostToDevice, h->stream));
        if (nnz > 0) {
            CU_TRY(cudaMemcpyAsync(dst.indices.ptr, indices, sizeof(int32_t) * nnz, cudaMemcpyHostToDevice, h->stream));
            std::vector<T> vals(nnz);
            for (int64_t e = 0; e < nnz; ++e) vals[e] = values ? (T)values[e] : T(1);
            CU_TRY(cudaMemcpyAsync(dst.values.ptr, vals.data(), sizeof(T) * nnz, cudaMemcpyHostToDevice, h->stream));
            CU_TRY(cudaStreamSynchronize(h->stream));   // vals goes out of scope
        }
        return HYPREC_OK;
    }

    int set_train(int64_t m_, int64_t n_, const int64_t* indptr, const int32_t* indices, const double* values) override {
        if (m_ <= 0 || n_ <= 0) return fail(HYPREC_E_INVALID, "set_train_csr: empty matrix");
        if (m_ != m || n_ != n) { k = 0; gram_ok[0] = gram_ok[1] = false; }
        cross_valid = false;
        plan[0].lo = plan[1].lo = -1;
        cand_lo = cand_hi = -1;
        use_cached_cands = false;
        m = m_;
        n = n_;
        HY_TRY(upload_csr(csr, m, n, indptr, indices, values, "train"));
        // item-major copy (counting sort keeps user ids increasing inside an item's list)
        const int64_t nnz = indptr[m];
        std::vector<int64_t> cptr(n + 1, 0);
        for (int64_t e = 0; e < nnz; ++e) ++cptr[indices[e] + 1];
        for (int64_t j = 0; j < n; ++j) cptr[j + 1] += cptr[j];
        std::vector<int32_t> cidx(std::max<int64_t>(nnz, 1));
        std::vector<double> cval(std::max<int64_t>(nnz, 1));
        std::vector<int64_t> fill(cptr.begin(), cptr.end() - 1);
        for (int64_t r = 0; r < m; ++r)
            for (int64_t e = indptr[r]; e < indptr[r + 1]; ++e) {
                const int64_t at = fill[indices[e]]++;
                cidx[at] = (int32_t)r;
                cval[at] = values ? values[e] : 1.0;
            }
        HY_TRY(upload_csr(csc, n, m, cptr.data(), cidx.data(), cval.data(), "train^T"));
        return HYPREC_OK;
    }

    int set_eval(const int64_t* fp, const int32_t* fi, const double* fv, const int64_t* tp, const int32_t* ti,
                 const double* tv) override {
        if (m == 0) return fail(HYPREC_E_INVALID, "set_eval_csr before set_train_csr");
        HY_TRY(upload_csr(full, m, n, fp, fi, fv, "ratings"));
        // hit values are float64 in every mode (evaluator.py:287 multiplies the stored rating)
        if (full.nnz > 0) {
            HY_TRY(full.values.reserve(sizeof(double) * full.nnz));
            std::vector<double> vals(full.nnz);
            hits_u8_exact = true;
            for (int64_t e = 0; e < full.nnz; ++e) {
                vals[e] = fv ? fv[e] : 1.0;
                if (!(vals[e] >= 0.0 && vals[e] <= 255.0 && vals[e] == (double)(uint8_t)vals[e])) hits_u8_exact = false;
            }
            CU_TRY(cudaMemcpyAsync(full.values.ptr, vals.data(), sizeof(double) * full.nnz, cudaMemcpyHostToDevice, h->stream));
            CU_TRY(cudaStreamSynchronize(h->stream));
        }
        HY_TRY(upload_csr(test, m, n, tp, ti, tv, "test"));
        if (cands_are_hashed) { use_cached_cands = false; cands_are_hashed = false; cand_lo = cand_hi = -1; }
        return HYPREC_OK;
    }

    // Host factors are float64 (the reference's dtype); in float32 mode the conversion runs on the
    // device through a float64 staging buffer, so the host side is a plain (DMA-able) copy either way.
    int upload_matrix(DevBuf& dst, const double* src, int64_t rows) {
        const size_t count = (size_t)rows * k;
        HY_TRY(dst.reserve(sizeof(T) * count));
        if (sizeof(T) == sizeof(double)) {
            CU_TRY(cudaMemcpyAsync(dst.ptr, src, sizeof(double) * count, cudaMemcpyHostToDevice, h->stream));
        } else {
            HY_TRY(stage_d.reserve(sizeof(double) * count));
            CU_TRY(cudaMemcpyAsync(stage_d.ptr, src, sizeof(double) * count, cudaMemcpyHostToDevice, h->stream));
            hyprec::cast_kernel<double, T><<<(unsigned)((count + 255) / 256), 256, 0, h->stream>>>(stage_d.as<double>(), (int64_t)count,
                                                                                                 dst.template as<T>());
            h->launches += 1;
        }
        CU_TRY(cudaStreamSynchronize(h->stream));
        return HYPREC_OK;
    }

    int download_matrix(const DevBuf& src, double* dst, int64_t rows) {
        const size_t count = (size_t)rows * k;
        if (sizeof(T) == sizeof(double)) {
            CU_TRY(cudaMemcpyAsync(dst, src.ptr, sizeof(double) * count, cudaMemcpyDeviceToHost, h->stream));
        } else {
            HY_TRY(stage_d.reserve(sizeof(double) * count));
            hyprec::cast_kernel<T, double><<<(unsigned)((count + 255) / 256), 256, 0, h->stream>>>(src.template as<T>(), (int64_t)count,
                                                                                                 stage_d.as<double>());
            h->launches += 1;
            CU_TRY(cudaMemcpyAsync(dst, stage_d.ptr, sizeof(double) * count, cudaMemcpyDeviceToHost, h->stream));
        }
        CU_TRY(cudaStreamSynchronize(h->stream));
        return HYPREC_OK;
    }

    int set_factors(const double* u, const double* v, int k_) override {
        if (m == 0) return fail(HYPREC_E_INVALID, "set_factors before set_train_csr");
        if (k_ <= 0 || k_ > 2000) return fail(HYPREC_E_UNSUPPORTED, "n_factors must be in [1, 2000]");
        if (!u || !v) return fail(HYPREC_E_INVALID, "set_factors: null matrix");
        if (k_ != k) { has_prior = false; }
        k = k_;
        HY_TRY(upload_matrix(U, u, m));
        HY_TRY(upload_matrix(V, v, n));
        factors_changed();
        return HYPREC_OK;
    }

    void factors_changed() {
        gram_ok[0] = gram_ok[1] = false;
        cross_valid = false;
        cm.gram64_ok[0] = cm.gram64_ok[1] = false;
        cm.q_ok[0] = cm.q_ok[1] = false;
        cm.prep_side = -1;
        cm.slots_have_rmse = false;
    }

    // U[e] and V[e] drawn on the device from a counter-based generator (comm.cuh uniform_fill_kernel):
    // numpy.random.random((m, k)) / ((n, k)) of /root/reference/lib/collaborative_filtering.py:173-176 for
    // problem sizes whose factors should not cross PCIe (seeds `seed` and `seed + 1`).
    int init_factors_uniform(int k_, unsigned long long seed) override {
        if (m == 0) return fail(HYPREC_E_INVALID, "init_factors_uniform before set_train_csr");
        if (k_ <= 0 || k_ > 2000) return fail(HYPREC_E_UNSUPPORTED, "n_factors must be in [1, 2000]");
        if (k_ != k) { has_prior = false; }
        k = k_;
        HY_TRY(U.reserve(sizeof(T) * (size_t)m * k));
        HY_TRY(V.reserve(sizeof(T) * (size_t)n * k));
        const int64_t cu = m * k, cv = n * k;
        hyprec::uniform_fill_kernel<T><<<(unsigned)((cu + 255) / 256), 256, 0, h->stream>>>(U.as<T>(), cu, seed);
        hyprec::uniform_fill_kernel<T><<<(unsigned)((cv + 255) / 256), 256, 0, h->stream>>>(V.as<T>(), cv, seed + 1);
        h->launches += 2;
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaStreamSynchronize(h->stream));
        factors_changed();
        return HYPREC_OK;
    }

    int get_factor_rows(int side, const int64_t* rows, int64_t n_rows, double* out) override {
        if (k == 0) return fail(HYPREC_E_INVALID, "get_factor_rows before set_factors");
        if (side != 0 && side != 1) return fail(HYPREC_E_INVALID, "side must be 0 or 1");
        if (n_rows <= 0) return HYPREC_OK;
        if (!rows || !out) return fail(HYPREC_E_INVALID, "get_factor_rows: null array");
        const int64_t limit = side == 0 ? m : n;
        for (int64_t i = 0; i < n_rows; ++i)
            if (rows[i] < 0 || rows[i] >= limit) return fail(HYPREC_E_INVALID, "get_factor_rows: row id out of range");
        HY_TRY(rows_idx.reserve(sizeof(int64_t) * (size_t)n_rows));
        HY_TRY(rows_out.reserve(sizeof(double) * (size_t)n_rows * k));
        CU_TRY(cudaMemcpyAsync(rows_idx.ptr, rows, sizeof(int64_t) * (size_t)n_rows, cudaMemcpyHostToDevice, h->stream));
        hyprec::gather_rows_kernel<T><<<(unsigned)n_rows, 128, 0, h->stream>>>(side == 0 ? U.as<T>() : V.as<T>(), k,
                                                                              rows_idx.as<int64_t>(), n_rows, rows_out.as<double>());
        h->launches += 1;
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaMemcpyAsync(out, rows_out.ptr, sizeof(double) * (size_t)n_rows * k, cudaMemcpyDeviceToHost, h->stream));
        CU_TRY(cudaStreamSynchronize(h->stream));
        return HYPREC_OK;
    }

    int get_factors(double* u, double* v) override {
        if (k == 0) return fail(HYPREC_E_INVALID, "get_factors before set_factors");
        if (u) HY_TRY(download_matrix(U, u, m));
        if (v) HY_TRY(download_matrix(V, v, n));
        return HYPREC_OK;
    }

    int set_prior(const double* theta) override {
        if (!theta) { has_prior = false; return HYPREC_OK; }
        if (k == 0) return fail(HYPREC_E_INVALID, "set_item_prior before set_factors");
        HY_TRY(upload_matrix(prior, theta, n));
        has_prior = true;
        return HYPREC_OK;
    }

    int device_factors(void** u, void** v, int* elem) override {
        if (k == 0) return fail(HYPREC_E_INVALID, "device_factors before set_factors");
        if (u) *u = U.ptr;
        if (v) *v = V.ptr;
        if (elem) *elem = (int)sizeof(T);
        gram_ok[0] = gram_ok[1] = false;   // the caller may write through these pointers
        cm.gram64_ok[0] = cm.gram64_ok[1] = false;
        cm.q_ok[0] = cm.q_ok[1] = false;
        cm.prep_side = -1;
        // With a shard set the pointers serve the exchange of the side just solved (rows of the OTHER
        // ranks change): the cross term of this rank's item rows stays what the item half-step left.
        if (h->user_hi < 0 && h->item_hi < 0) cross_valid = false;
        return HYPREC_OK;
    }

    // ---------------------------------------------------------------- Gram / column sums
    int compute_gram(int which) {   // 0: U, 1: V
        if (gram_ok[which]) return HYPREC_OK;
        const T* y = which == 0 ? U.as<T>() : V.as<T>();
        const int64_t rows = which == 0 ? m : n;
        const int ntile = (k + hyprec::kGramTile - 1) / hyprec::kGramTile;
        const int tiles = hyprec::tri(ntile);
        int nsplit = (int)std::min<int64_t>((rows + 255) / 256, std::max(1, (4 * h->num_sms) / tiles));
        nsplit = std::max(nsplit, 1);
        int64_t per = (rows + nsplit - 1) / nsplit;
        per = (per + hyprec::kGramRows - 1) / hyprec::kGramRows * hyprec::kGramRows;
        nsplit = (int)((rows + per - 1) / per);
        HY_TRY(gram[which].reserve(sizeof(T) * (size_t)k * k));
        if constexpr (sizeof(T) == sizeof(float)) {
            // float32 mode: Y^T Y on the tensor cores (3xTF32, tc_syrk.cuh gram_tc_kernel)
            if (h->tc_solve_enabled && k % 2 == 0 && k >= 16 && k <= 256 && rows >= 256) {
                namespace tc = hyprec::tc;
                const tc::GramTcLayout lay = tc::gram_tc_layout(k);
                if (lay.total <= (size_t)h->max_smem) {
                    const int grid = h->num_sms;
                    int64_t per_cta = (rows + grid - 1) / grid;
                    per_cta = (per_cta + 31) / 32 * 32;
                    HY_TRY(gram_partial.reserve(sizeof(float) * (size_t)grid * k * k));
                    HY_TRY(ensure_dynamic_smem((const void*)tc::gram_tc_kernel, (int)lay.total));
                    Timer t(h, HYPREC_T_GRAM);
                    tc::gram_tc_kernel<<<grid, 256, lay.total, h->stream>>>(reinterpret_cast<const float*>(y), rows, k, per_cta,
                                                                             gram_partial.as<float>());
                    tc::gram_tc_reduce_kernel<<<(k * k + 255) / 256, 256, 0, h->stream>>>(gram_partial.as<float>(), grid, k,
                                                                                           gram[which].template as<float>());
                    h->launches += 2;
                    CU_TRY(cudaGetLastError());
                    gram_ok[which] = true;
                    return HYPREC_OK;
                }
            }
        }
        HY_TRY(gram_partial.reserve(sizeof(T) * (size_t)nsplit * k * k));
        Timer t(h, HYPREC_T_GRAM);
        hyprec::gram_partial_kernel<T><<<dim3(tiles, nsplit), hyprec::kGramThreads, 0, h->stream>>>(
            y, rows, k, per, gram_partial.as<T>());
        hyprec::gram_reduce_kernel<T><<<(k * k + 255) / 256, 256, 0, h->stream>>>(
            gram_partial.as<T>(), nsplit, k, gram[which].as<T>());
        h->launches += 2;
        CU_TRY(cudaGetLastError());
        gram_ok[which] = true;
        return HYPREC_OK;
    }

    int compute_colsum_v() {
        int nsplit = (int)std::min<int64_t>(std::max<int64_t>(n / 64, 1), 4 * h->num_sms);
        const int64_t per = (n + nsplit - 1) / nsplit;
        nsplit = (int)((n + per - 1) / per);
        HY_TRY(colsum_partial.reserve(sizeof(double) * (size_t)nsplit * k));
        HY_TRY(colsum.reserve(sizeof(double) * k));
        hyprec::colsum_partial_kernel<T><<<nsplit, 256, 0, h->stream>>>(V.as<T>(), n, k, per, colsum_partial.as<double>());
        hyprec::colsum_reduce_kernel<<<(k + 255) / 256, 256, 0, h->stream>>>(colsum_partial.as<double>(), nsplit, k,
                                                                               colsum.as<double>());
        h->launches += 2;
        CU_TRY(cudaGetLastError());
        return HYPREC_OK;
    }

    // ---------------------------------------------------------------- ALS half-step
    // Launch the row kernel over a contiguous range or a row list, in direct or low-rank mode.
    int launch_rows(bool user, double lambda, int64_t lo, int64_t hi, const int32_t* list, int64_t n_list,
                    bool lowrank, int max_dim, int threads, const T* fixed_ptr, const T* prior_ptr, T* export_ptr,
                    const int64_t* indptr_override, int64_t work, cudaStream_t st, int slot, int ts = 8, bool tc = false,
                    int64_t ymat_row0 = 0) {
        // export_ptr != null: the single-row launch that only factors B -- its (zero) solution must
        // not land in the factor matrix
        T* latent_ptr = user ? U.as<T>() : V.as<T>();
        if (export_ptr) {
            if (dummy_out.reserve(sizeof(T) * (size_t)k) != HYPREC_OK) return HYPREC_E_NOMEM;
            latent_ptr = dummy_out.as<T>();
        }
        const Csr& rows_csr = user ? csr : csc;
        // shared-memory plan: tiles in shared memory when they fit, else per-CTA global scratch
        // rank-1 vectors staged per round: small systems take long rounds (fewer gather latencies)
        int chunk = lowrank ? (max_dim <= 63 ? 64 : (max_dim <= 127 ? 32 : 16)) : 32;
        if (tc) chunk = 8;   // the staging buffer only serves as the panel buffer
        const size_t stage_bytes = tc ? hyprec::tc::syrk_stage_bytes(max_dim) : 0;
        hyprec::SolveLayout lay = hyprec::solve_layout(ts, max_dim, chunk, true, sizeof(T), stage_bytes);
        while (lay.total > (size_t)h->max_smem && chunk > 8) {
            chunk /= 2;
            lay = hyprec::solve_layout(ts, max_dim, chunk, true, sizeof(T), stage_bytes);
        }
        const bool in_smem = lay.total <= (size_t)h->max_smem;
        if (tc && !in_smem) return fail(HYPREC_E_UNSUPPORTED, "tensor-core row launch does not fit shared memory");
        if (!in_smem) {
            chunk = 16;
            lay = hyprec::solve_layout(ts, max_dim, chunk, false, sizeof(T));
            if (lay.total > (size_t)h->max_smem) return fail(HYPREC_E_UNSUPPORTED, "n_factors too large for the solve kernel");
        }
        auto kern = ts == 4 ? hyprec::als_solve_rows_kernel<T, 4> : hyprec::als_solve_rows_kernel<T, 8>;
        if constexpr (sizeof(T) == sizeof(float)) {
            if (tc) kern = hyprec::als_solve_rows_kernel<T, 8, true>;
        }
        if (!in_smem) {
            if (ts != 8) return fail(HYPREC_E_UNSUPPORTED, "global tile scratch is only built for 8 x 8 tiles");
            kern = hyprec::als_solve_rows_kernel<T, 8, false, true>;
        }
        HY_TRY(ensure_dynamic_smem((const void*)kern, (int)lay.total));
        int per_sm = 0;
        CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, lay.total));
        if (per_sm < 1) return fail(HYPREC_E_UNSUPPORTED, "solve kernel does not fit on an SM");
        const int grid = (int)std::min<int64_t>(work, (int64_t)per_sm * h->num_sms);
        // (work counters of all slots were zeroed at the start of the half-step)
        DevBuf& scratch_buf = a_scratch_slot[slot];
        if (!in_smem) HY_TRY(scratch_buf.reserve(sizeof(T) * (size_t)grid * ts * ts * lay.n_tiles));

        hyprec::SolveArgs<T> a;
        a.indptr = indptr_override ? indptr_override : rows_csr.indptr.template as<int64_t>();
        a.indices = rows_csr.indices.template as<int32_t>();
        a.values = rows_csr.values.template as<T>();
        a.fixed = fixed_ptr;
        a.latent = latent_ptr;
        a.base = base.as<T>();
        a.prior = prior_ptr;
        a.a_scratch = in_smem ? nullptr : scratch_buf.template as<T>();
        a.work_counter = counter.as<int>() + slot;
        a.status = counter.as<int>() + kStatusSlot;
        a.lambda = (T)lambda;
        a.k = k;
        a.row_lo = lo;
        a.row_hi = hi;
        a.chunk_rows = chunk;
        a.row_list = list;
        a.n_list = n_list;
        a.max_dim = max_dim;
        a.lowrank = lowrank ? 1 : 0;
        a.ymat = ymat.as<T>();
        a.ymat_row0 = ymat_row0;
        a.export_tiles = export_ptr;
        a.tmem_cols = tc ? hyprec::tc::syrk_tmem_cols(max_dim) : 0;
        a.cross_out = (cross_active && !export_ptr) ? cross.as<double>() : nullptr;
        a.prof = nullptr;
        if (h->prof_enabled && !export_ptr && slot == h->prof_slot) {
            if (prof_buf.reserve(16 * sizeof(long long)) != HYPREC_OK) return HYPREC_E_NOMEM;
            a.prof = prof_buf.as<long long>();
        }
        kern<<<grid, threads, lay.total, st>>>(a);
        h->launches += 1;
        CU_TRY(cudaGetLastError());
        return HYPREC_OK;
    }


    // float32 systems of 33..256 unknowns with the normal matrix resident in tensor memory (tc_chol.cuh).
    int launch_chol_rows(bool user, double lambda, int64_t lo, int64_t hi, const int32_t* list, int64_t n_list,
                         bool lowrank, int max_dim, const float* fixed_ptr, const float* prior_ptr, int64_t work,
                         cudaStream_t st, int slot, int64_t out_row0 = 0, int n_split = 0, const int64_t* split_lo = nullptr,
                         const float* split_gram_p = nullptr, const float* split_b_p = nullptr) {
        namespace tc = hyprec::tc;
        const tc::CholLayout lay = tc::chol_layout(max_dim);
        if (lay.total > (size_t)h->max_smem) return fail(HYPREC_E_UNSUPPORTED, "tensor-memory row launch does not fit shared memory");
        HY_TRY(ensure_dynamic_smem((const void*)tc::tc_chol_rows_kernel, (int)lay.total));
        const Csr& rows_csr = user ? csr : csc;
        tc::CholRowsArgs a;
        a.indptr = rows_csr.indptr.template as<int64_t>();
        a.indices = rows_csr.indices.template as<int32_t>();
        a.values = rows_csr.values.template as<float>();
        a.fixed = fixed_ptr;
        a.gram = gram[user ? 1 : 0].template as<float>();
        a.prior = prior_ptr;
        a.out = lowrank ? ymat.as<float>() : (user ? U.as<float>() : V.as<float>());
        a.out_row0 = out_row0;
        a.work_counter = counter.as<int>() + slot;
        a.status = counter.as<int>() + kStatusSlot;
        a.lambda = (float)lambda;
        a.k = k;
        a.row_lo = lo;
        a.row_hi = hi;
        a.row_list = list;
        a.n_list = n_list;
        a.max_dim = max_dim;
        a.lowrank = lowrank ? 1 : 0;
        a.tmem_cols = tc::syrk_tmem_cols(max_dim);
        a.cross_out = cross_active ? cross.as<double>() : nullptr;
        a.n_split = n_split;
        a.split_lo = split_lo;
        a.split_gram = split_gram_p;
        a.split_b = split_b_p;
        a.prof = nullptr;
        if (h->prof_enabled && slot == h->prof_slot) {
            if (prof_buf.reserve(16 * sizeof(long long)) != HYPREC_OK) return HYPREC_E_NOMEM;
            a.prof = prof_buf.as<long long>();
        }
        const int grid = (int)std::min<int64_t>(work, (int64_t)h->num_sms);
        tc::tc_chol_rows_kernel<<<grid, tc::kCholThreads, lay.total, st>>>(a);
        h->launches += 1;
        CU_TRY(cudaGetLastError());
        return HYPREC_OK;
    }


    // low-rank systems of <= 32 unknowns, one warp per row, S built with FP64 MMAs (small_f64.cuh), both precisions
    int launch_small(bool user, const int32_t* list, int64_t n_list, const T* q_ptr, cudaStream_t st, int slot, int64_t out_row0) {
        namespace hp = hyprec;
        const Csr& rows_csr = user ? csr : csc;
        const size_t smem = hp::small_smem_bytes();
        auto kern = hp::lowrank_small_kernel<T>;
        HY_TRY(ensure_dynamic_smem((const void*)kern, (int)smem));
        hp::SmallArgsT<T> a;
        a.indptr = rows_csr.indptr.template as<int64_t>();
        a.indices = rows_csr.indices.template as<int32_t>();
        a.values = rows_csr.values.template as<T>();
        a.fixed = q_ptr;
        a.ymat = ymat.as<T>();
        a.ymat_row0 = out_row0;
        a.row_list = list;
        a.n_list = n_list;
        a.k = k;
        a.work_counter = counter.as<int>() + slot;
        a.status = counter.as<int>() + kStatusSlot;
        a.cross_out = cross_active ? cross.as<double>() : nullptr;
        int per_sm = 0;
        CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, hp::kSmallWarps * 32, smem));
        per_sm = std::max(per_sm, 1);
        const int64_t ctas = (n_list + hp::kSmallWarps - 1) / hp::kSmallWarps;
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ctas, (int64_t)per_sm * h->num_sms));
        kern<<<grid, hp::kSmallWarps * 32, smem, st>>>(a);
        h->launches += 1;
        CU_TRY(cudaGetLastError());
        return HYPREC_OK;
    }

    // several float32 low-rank systems of <= slot (32 / 64 / 128) unknowns per CTA pass (tc_chol_multi.cuh)
    int launch_chol_multi(bool user, double lambda, int64_t lo, int64_t hi, const int32_t* list, int64_t n_list, int slot_rows,
                          const float* fixed_ptr, const float* prior_ptr, int64_t work, cudaStream_t st, int slot, int64_t out_row0) {
        namespace tc = hyprec::tc;
        const tc::CholMultiLayout lay = tc::chol_multi_layout();
        if (lay.total > (size_t)h->max_smem) return fail(HYPREC_E_UNSUPPORTED, "tensor-memory row launch does not fit shared memory");
        auto kern = slot_rows == 32 ? tc::tc_chol_multi_kernel<32> : (slot_rows == 64 ? tc::tc_chol_multi_kernel<64> : tc::tc_chol_multi_kernel<128>);
        HY_TRY(ensure_dynamic_smem((const void*)kern, (int)lay.total));
        const Csr& rows_csr = user ? csr : csc;
        tc::CholRowsArgs a;
        a.indptr = rows_csr.indptr.template as<int64_t>();
        a.indices = rows_csr.indices.template as<int32_t>();
        a.values = rows_csr.values.template as<float>();
        a.fixed = fixed_ptr;
        a.gram = nullptr;
        a.prior = prior_ptr;
        a.out = ymat.as<float>();
        a.out_row0 = out_row0;
        a.work_counter = counter.as<int>() + slot;
        a.status = counter.as<int>() + kStatusSlot;
        a.lambda = (float)lambda;
        a.k = k;
        a.row_lo = lo;
        a.row_hi = hi;
        a.row_list = list;
        a.n_list = n_list;
        a.max_dim = slot_rows;
        a.lowrank = 1;
        a.tmem_cols = 256;
        a.cross_out = cross_active ? cross.as<double>() : nullptr;
        a.prof = nullptr;
        if (h->prof_enabled && slot == h->prof_slot) {
            if (prof_buf.reserve(16 * sizeof(long long)) != HYPREC_OK) return HYPREC_E_NOMEM;
            a.prof = prof_buf.as<long long>();
        }
        const int per_pass = 256 / slot_rows;
        const int grid = (int)std::min<int64_t>((work + per_pass - 1) / per_pass, (int64_t)h->num_sms);
        kern<<<grid, tc::kCholThreads, lay.total, st>>>(a);
        h->launches += 1;
        CU_TRY(cudaGetLastError());
        return HYPREC_OK;
    }

    // C[ra x rb] = A[ra x k] B[rb x k]^T through the SIMT sweep kernel (float64 mode; float32 mode
    // goes through gemm_nt_tc below unless the tensor cores are switched off).  With a row
    // list, row i of the product (A is compact) is stored at C[list[i]].
    int gemm_nt(const T* a, int64_t ra, const T* b, int64_t rb, T* c, const int32_t* list = nullptr, void** peers = nullptr,
                int n_peer = 0) {
        hyprec::SweepArgs<T> s;
        s.user_vecs = a;
        s.item_vecs = b;
        s.k = k;
        s.user_lo = 0;
        s.user_hi = ra;
        s.n_items = rb;
        s.dense_out = sizeof(T) == sizeof(double) ? reinterpret_cast<double*>(c) : nullptr;
        s.dense_out_f32 = sizeof(T) == sizeof(double) ? nullptr : reinterpret_cast<float*>(c);
        s.row_list = list;
        s.thresholds = nullptr;
        s.counts = nullptr;
        s.n_peer = n_peer;
        for (int i = 0; i < n_peer; ++i) s.peer_out[i] = peers[i];
        dim3 grid((unsigned)((rb + hyprec::kSweepTile - 1) / hyprec::kSweepTile),
                  (unsigned)((ra + hyprec::kSweepTile - 1) / hyprec::kSweepTile));
        if (n_peer > 0)
            hyprec::score_sweep_kernel<T, hyprec::kSweepStorePeers><<<grid, hyprec::kSweepThreads, 0, h->stream>>>(s);
        else
            hyprec::score_sweep_kernel<T, hyprec::kSweepStore><<<grid, hyprec::kSweepThreads, 0, h->stream>>>(s);
        h->launches += 1;
        CU_TRY(cudaGetLastError());
        return HYPREC_OK;
    }

    // float32 mode: the same product on the tensor cores (tc_score_kernel, 3xTF32) -- the whitening
    // Q = Y L^-T and the back-transform x = L^-T y are real dense contractions.
    DevBuf w_hi, w_lo, l_hi, l_lo;
    int gemm_nt_tc(const float* a, int64_t ra, const float* b, float* c, const int32_t* row_map, void** peers = nullptr,
                   int n_peer = 0) {
        namespace tcn = hyprec::tc;
        if (ra == 0) return HYPREC_OK;
        HY_TRY(w_hi.reserve(sizeof(float) * (size_t)ra * k));
        HY_TRY(w_lo.reserve(sizeof(float) * (size_t)ra * k));
        HY_TRY(l_hi.reserve(sizeof(float) * (size_t)k * k));
        HY_TRY(l_lo.reserve(sizeof(float) * (size_t)k * k));
        const int64_t ca = ra * k, cb = (int64_t)k * k;
        tcn::split_tf32_kernel<float><<<(unsigned)((ca + 255) / 256), 256, 0, h->stream>>>(a, ca, w_hi.as<float>(), w_lo.as<float>());
        tcn::split_tf32_kernel<float><<<(unsigned)((cb + 255) / 256), 256, 0, h->stream>>>(b, cb, l_hi.as<float>(), l_lo.as<float>());
        h->launches += 2;
        CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
        HY_TRY(make_map(&ma_hi, w_hi.as<float>(), ra));
        HY_TRY(make_map(&ma_lo, w_lo.as<float>(), ra));
        HY_TRY(make_map(&mb_hi, l_hi.as<float>(), k));
        HY_TRY(make_map(&mb_lo, l_lo.as<float>(), k));
        tcn::ScoreArgs s;
        s.m = ra;
        s.n = k;
        s.k = k;
        s.thresholds = nullptr;
        s.unorm = nullptr;
        s.vnorm = nullptr;
        s.margin_coef = 0.f;
        s.counts = nullptr;
        s.border_count = nullptr;
        s.border_cells = nullptr;
        s.border_capacity = 0;
        s.dense = c;
        s.dense_ld = k;
        s.row_map = row_map;
        s.n_dense_peer = n_peer;
        for (int i = 0; i < n_peer; ++i) s.dense_peer[i] = static_cast<float*>(peers[i]);
        auto kern = tcn::tc_score_kernel;
        HY_TRY(ensure_dynamic_smem((const void*)kern, (int)tcn::kSmemBytes));
        const int64_t n_tiles = ((ra + tcn::kBM - 1) / tcn::kBM) * ((k + tcn::kBN - 1) / tcn::kBN);
        const int grid = (int)std::min<int64_t>(n_tiles, h->num_sms);
        kern<<<grid, tcn::kThreads, tcn::kSmemBytes, h->stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, s);
        h->launches += 1;
        CU_TRY(cudaGetLastError());
        return HYPREC_OK;
    }
    // whitening products: tensor cores in float32 mode, the float64 sweep kernel otherwise
    // peers / n_peer: replicas of c on the other GPUs that receive the same rows (multi-GPU exchange)
    int whiten_gemm(const T* a, int64_t ra, const T* b, T* c, const int32_t* list, void** peers = nullptr, int n_peer = 0) {
        if (ra <= 0) return HYPREC_OK;
        if constexpr (sizeof(T) == sizeof(float)) {
            if (h->tc_solve_enabled && tc_usable() && h->encode_tiled != nullptr && ra >= 128)
                return gemm_nt_tc(reinterpret_cast<const float*>(a), ra, reinterpret_cast<const float*>(b),
                                  reinterpret_cast<float*>(c), list, peers, n_peer);
        }
        return gemm_nt(a, ra, b, k, c, list, peers, n_peer);
    }

    // Rows of [lo, hi) grouped for the low-rank path: bucket b holds rows with
    // deg <= kBucketDim[b] (sorted by degree, heaviest first), `heavy` the rest.
    struct RowPlan {
        int64_t lo = -1, hi = -1;
        int cap = -2;
        bool allow_split = false;
        int n_buckets = 0;
        std::vector<int64_t> bucket_off;   // offsets into `lists`, size n_buckets + 2 (last = heavy)
        std::vector<int> bucket_dim;
        DevBuf lists;
        // heavy rows split over several CTAs (float32 tensor-memory path): the first n_split rows of the heavy list
        int n_split = 0, n_seg = 0;
        int64_t split_work = 0;            // positives per work item the plan aimed at
        DevBuf seg_lo, seg_len, seg_ptr, split_lo;
    };
    RowPlan plan[2];
    DevBuf split_partial, split_pb, split_gram, split_b;

    // Degree buckets of the low-rank path: upper dimension, threads per CTA, tile edge.
    static constexpr int kMaxBuckets = 6;
    static constexpr int kBucketDim[kMaxBuckets] = {31, 63, 127, 191, 255, 319};
    static constexpr int kSplitCounterSlot = 8;
    static constexpr int64_t kSplitMinWork = 4096;

    // Number of buckets whose systems fit one SM's shared memory in this precision.
    int usable_buckets() const {
        int nbk = 0;
        for (int b = 0; b < kMaxBuckets; ++b)
            if (hyprec::solve_layout(h->bucket_tile[b], kBucketDim[b], 8, true, sizeof(T)).total <= (size_t)h->max_smem) nbk = b + 1;
        return nbk;
    }

    // cap < 0: every row is "heavy" (direct k x k systems only).  allow_split: rows whose positives exceed ~1.5 work
    // items (work item = max(4096, heavy positives of the shard / (4 * SMs))) are cut into segments that
    // split_partial_kernel accumulates on other CTAs; the row's own CTA keeps the tail segment.
    int build_plan(bool user, int64_t lo, int64_t hi, int cap, bool allow_split) {
        RowPlan& pl = plan[user ? 0 : 1];
        if (pl.lo == lo && pl.hi == hi && pl.cap == cap && pl.allow_split == allow_split) return HYPREC_OK;
        const std::vector<int64_t>& ptr = (user ? csr : csc).h_indptr;
        const int nbk = usable_buckets();
        std::vector<std::vector<int32_t>> groups(nbk + 1);
        int64_t heavy_positives = 0;
        for (int64_t r = lo; r < hi; ++r) {
            const int64_t deg = ptr[r + 1] - ptr[r];
            int b = nbk;                        // heavy: direct k x k system
            if (deg <= cap)
                for (int i = 0; i < nbk; ++i)
                    if (deg <= kBucketDim[i]) { b = i; break; }
            if (b == nbk) heavy_positives += deg;
            groups[b].push_back((int32_t)r);
        }
        std::vector<int32_t> flat;
        pl.bucket_off.assign(1, 0);
        pl.bucket_dim.clear();
        for (int b = 0; b <= nbk; ++b) {
            std::stable_sort(groups[b].begin(), groups[b].end(), [&](int32_t x, int32_t y) {
                return ptr[x + 1] - ptr[x] > ptr[y + 1] - ptr[y];
            });
            int maxdeg = 1;
            if (!groups[b].empty()) maxdeg = (int)std::max<int64_t>(1, ptr[groups[b][0] + 1] - ptr[groups[b][0]]);
            pl.bucket_dim.push_back(b < nbk ? std::min(maxdeg, kBucketDim[b]) : k);
            flat.insert(flat.end(), groups[b].begin(), groups[b].end());
            pl.bucket_off.push_back((int64_t)flat.size());
        }
        HY_TRY(pl.lists.reserve(sizeof(int32_t) * std::max<size_t>(flat.size(), 1)));
        if (!flat.empty())
            CU_TRY(cudaMemcpyAsync(pl.lists.ptr, flat.data(), sizeof(int32_t) * flat.size(), cudaMemcpyHostToDevice, h->stream));
        // ---- split plan over the (degree-sorted) heavy list
        pl.n_split = pl.n_seg = 0;
        std::vector<int64_t> seg_lo_h, split_lo_h;
        std::vector<int32_t> seg_len_h, seg_ptr_h(1, 0);
        const int64_t work = std::max<int64_t>(kSplitMinWork, heavy_positives / (4 * (int64_t)std::max(1, h->num_sms)));
        pl.split_work = work;
        if (allow_split) {
            for (int32_t r : groups[nbk]) {
                const int64_t deg = ptr[r + 1] - ptr[r];
                if (deg <= work + work / 2) break;          // sorted: nothing heavier follows
                const int64_t parts = std::min<int64_t>(64, (deg + work - 1) / work);
                const int64_t part_len = ((deg + parts - 1) / parts + 31) / 32 * 32;
                int64_t at = ptr[r];
                for (int64_t q = 0; q + 1 < parts; ++q) {
                    seg_lo_h.push_back(at);
                    seg_len_h.push_back((int32_t)part_len);
                    at += part_len;
                }
                split_lo_h.push_back(at);                     // the tail stays with the row's own CTA
                seg_ptr_h.push_back((int32_t)seg_lo_h.size());
            }
            pl.n_split = (int)split_lo_h.size();
            pl.n_seg = (int)seg_lo_h.size();
            if (pl.n_split > 0) {
                HY_TRY(pl.seg_lo.reserve(sizeof(int64_t) * seg_lo_h.size()));
                HY_TRY(pl.seg_len.reserve(sizeof(int32_t) * seg_len_h.size()));
                HY_TRY(pl.seg_ptr.reserve(sizeof(int32_t) * seg_ptr_h.size()));
                HY_TRY(pl.split_lo.reserve(sizeof(int64_t) * split_lo_h.size()));
                CU_TRY(cudaMemcpyAsync(pl.seg_lo.ptr, seg_lo_h.data(), sizeof(int64_t) * seg_lo_h.size(), cudaMemcpyHostToDevice, h->stream));
                CU_TRY(cudaMemcpyAsync(pl.seg_len.ptr, seg_len_h.data(), sizeof(int32_t) * seg_len_h.size(), cudaMemcpyHostToDevice, h->stream));
                CU_TRY(cudaMemcpyAsync(pl.seg_ptr.ptr, seg_ptr_h.data(), sizeof(int32_t) * seg_ptr_h.size(), cudaMemcpyHostToDevice, h->stream));
                CU_TRY(cudaMemcpyAsync(pl.split_lo.ptr, split_lo_h.data(), sizeof(int64_t) * split_lo_h.size(), cudaMemcpyHostToDevice, h->stream));
            }
        }
        CU_TRY(cudaStreamSynchronize(h->stream));     // the host vectors go out of scope
        pl.lo = lo;
        pl.hi = hi;
        pl.cap = cap;
        pl.allow_split = allow_split;
        pl.n_buckets = nbk;
        return HYPREC_OK;
    }

    // The heavy list of a plan as k x k systems in tensor memory (float32): segments of split rows first
    // (split_partial_kernel + split_reduce_kernel), then one tc_chol_rows_kernel launch over the whole list.
    int launch_heavy(bool user, double lambda, int64_t lo, int64_t hi, const RowPlan& pl, const T* fixed_ptr, const T* prior_ptr,
                     cudaStream_t st, int slot) {
        if constexpr (sizeof(T) == sizeof(float)) {
            namespace tc = hyprec::tc;
            const int nbk = pl.n_buckets;
            const int64_t n_heavy = pl.bucket_off[nbk + 1] - pl.bucket_off[nbk];
            if (n_heavy == 0) return HYPREC_OK;
            const int32_t* list = pl.lists.template as<int32_t>() + pl.bucket_off[nbk];
            if (pl.n_split > 0) {
                const Csr& rows_csr = user ? csr : csc;
                const tc::GramTcLayout lay = tc::gram_tc_layout(k);
                if (lay.total > (size_t)h->max_smem) return fail(HYPREC_E_UNSUPPORTED, "split rows do not fit shared memory");
                HY_TRY(ensure_dynamic_smem((const void*)tc::split_partial_kernel, (int)lay.total));
                HY_TRY(split_partial.reserve(sizeof(float) * (size_t)pl.n_seg * k * k));
                HY_TRY(split_pb.reserve(sizeof(float) * (size_t)pl.n_seg * k));
                HY_TRY(split_gram.reserve(sizeof(float) * (size_t)pl.n_split * k * k));
                HY_TRY(split_b.reserve(sizeof(float) * (size_t)pl.n_split * k));
                tc::SplitArgs a;
                a.fixed = reinterpret_cast<const float*>(fixed_ptr);
                a.indices = rows_csr.indices.template as<int32_t>();
                a.values = rows_csr.values.template as<float>();
                a.seg_lo = pl.seg_lo.template as<int64_t>();
                a.seg_len = pl.seg_len.template as<int32_t>();
                a.n_seg = pl.n_seg;
                a.k = k;
                a.partial = split_partial.as<float>();
                a.partial_b = split_pb.as<float>();
                a.work_counter = counter.as<int>() + kSplitCounterSlot;
                const int grid = std::min(pl.n_seg, h->num_sms);
                tc::split_partial_kernel<<<grid, 256, lay.total, st>>>(a);
                const float ratio = (float)((hyprec::kConfPos - hyprec::kConfNeg) / hyprec::kConfNeg);
                tc::split_reduce_kernel<<<dim3((unsigned)((k * k + 255) / 256), (unsigned)pl.n_split), 256, 0, st>>>(
                    split_partial.as<float>(), split_pb.as<float>(), pl.seg_ptr.template as<int32_t>(),
                    gram[user ? 1 : 0].template as<float>(), k, ratio, split_gram.as<float>(), split_b.as<float>());
                h->launches += 2;
                CU_TRY(cudaGetLastError());
            }
            return launch_chol_rows(user, lambda, lo, hi, list, n_heavy, false, k, reinterpret_cast<const float*>(fixed_ptr),
                                    reinterpret_cast<const float*>(prior_ptr), n_heavy, st, slot, 0, pl.n_split,
                                    pl.n_split ? pl.split_lo.template as<int64_t>() : nullptr,
                                    pl.n_split ? split_gram.as<float>() : nullptr, pl.n_split ? split_b.as<float>() : nullptr);
        } else {
            (void)user; (void)lambda; (void)lo; (void)hi; (void)pl; (void)fixed_ptr; (void)prior_ptr; (void)st; (void)slot;
            return fail(HYPREC_E_UNSUPPORTED, "tensor-memory rows are float32 only");
        }
    }

    // L^-1 and L^-T (dense k x k, ALWAYS float64) of B = 0.1 G + lambda I, G = Gram of the fixed side.
    // B is cut into r x r blocks that fit one SM's shared memory as float64 tiles (r = 1 up to
    // k ~ 230, else blocks of <= 192): every diagonal block is factored by a direct-mode launch of the
    // row kernel on a system without positives (tiles exported) and inverted by tri_inverse_kernel;
    // panels and Schur updates are small float64 GEMMs (right-looking block Cholesky), and the inverse
    // is assembled by block forward substitution.  In float32 mode the Gram is recomputed in float64
    // from a float64 copy of the fixed side (the whitening must not see float32 rounding of B), and
    // float32 copies of L^-1 / L^-T are made for the whitening GEMMs.
    // float64 Gram of rows [r_lo, r_hi) of a rows x k factor matrix (any factor dtype) into out (k x k).
    int gram64_rows(const T* y, int64_t r_lo, int64_t r_hi, double* out) {
        namespace hp = hyprec;
        const int64_t rows = r_hi - r_lo;
        if (rows <= 0) {
            CU_TRY(cudaMemsetAsync(out, 0, sizeof(double) * (size_t)k * k, h->stream));
            return HYPREC_OK;
        }
        const double* src = nullptr;
        if constexpr (sizeof(T) == sizeof(double)) {
            src = reinterpret_cast<const double*>(y) + r_lo * k;
        } else {
            HY_TRY(fixed_d.reserve(sizeof(double) * (size_t)rows * k));
            const int64_t cnt = rows * k;
            hp::cast_kernel<T, double><<<(unsigned)((cnt + 255) / 256), 256, 0, h->stream>>>(y + r_lo * k, cnt, fixed_d.as<double>());
            h->launches += 1;
            src = fixed_d.as<double>();
        }
        const int ntile = (k + hp::kGramTile - 1) / hp::kGramTile, tiles_g = hp::tri(ntile);
        int nsplit = (int)std::min<int64_t>((rows + 255) / 256, std::max(1, (4 * h->num_sms) / tiles_g));
        nsplit = std::max(nsplit, 1);
        int64_t per = (rows + nsplit - 1) / nsplit;
        per = (per + hp::kGramRows - 1) / hp::kGramRows * hp::kGramRows;
        nsplit = (int)((rows + per - 1) / per);
        HY_TRY(gram_partial_d.reserve(sizeof(double) * (size_t)nsplit * k * k));
        hp::gram_partial_kernel<double><<<dim3(tiles_g, nsplit), hp::kGramThreads, 0, h->stream>>>(src, rows, k, per,
                                                                                                  gram_partial_d.as<double>());
        hp::gram_reduce_kernel<double><<<(k * k + 255) / 256, 256, 0, h->stream>>>(gram_partial_d.as<double>(), nsplit, k, out);
        h->launches += 2;
        CU_TRY(cudaGetLastError());
        return HYPREC_OK;
    }

    int compute_linv(bool user, double lambda, const T* fixed_ptr, int64_t n_fixed) {
        const double* gram_d = nullptr;
        if (sizeof(T) == sizeof(double)) {
            gram_d = gram[user ? 1 : 0].template as<double>();
        } else {
            HY_TRY(gram_dbuf.reserve(sizeof(double) * (size_t)k * k));
            HY_TRY(gram64_rows(fixed_ptr, 0, n_fixed, gram_dbuf.as<double>()));
            gram_d = gram_dbuf.as<double>();
        }
        return linv_from_gram(gram_d, lambda);
    }

    // L^-1 / L^-T of B = 0.1 G + lambda I from the float64 Gram G (see the comment above compute_linv).
    int linv_from_gram(const double* gram_d, double lambda) {
        namespace hp = hyprec;
        HY_TRY(linv.reserve(sizeof(double) * (size_t)k * k));
        HY_TRY(linv_t.reserve(sizeof(double) * (size_t)k * k));
        HY_TRY(dummy_ptr.reserve(2 * sizeof(int64_t)));
        CU_TRY(cudaMemsetAsync(dummy_ptr.ptr, 0, 2 * sizeof(int64_t), h->stream));
        const bool whole = hp::solve_layout(8, k, 8, true, sizeof(double)).total <= (size_t)h->max_smem;
        const int r = whole ? 1 : (k + 191) / 192;
        const int bs0 = whole ? k : (((k + r - 1) / r) + 7) / 8 * 8;
        auto off = [&](int j) { return j * bs0; };
        auto size = [&](int j) { return std::min(bs0, k - j * bs0); };
        HY_TRY(bdense.reserve(sizeof(double) * (size_t)k * k));
        HY_TRY(lfac.reserve(sizeof(double) * (size_t)k * k));
        HY_TRY(blk_tmp.reserve(sizeof(double) * (size_t)bs0 * bs0));
        const int blk_tiles_n = hp::kTileElems * hp::tri(hp::aug_blocks(bs0));
        HY_TRY(blk_base.reserve(sizeof(double) * (size_t)blk_tiles_n));
        HY_TRY(tiles.reserve(sizeof(double) * (size_t)blk_tiles_n));
        double* B = bdense.as<double>();
        double* L = lfac.as<double>();
        double* X = linv.as<double>();
        hp::build_base_dense_kernel<double><<<(k * k + 255) / 256, 256, 0, h->stream>>>(gram_d, k, lambda, B);
        CU_TRY(cudaMemsetAsync(X, 0, sizeof(double) * (size_t)k * k, h->stream));
        h->launches += 1;
        auto gemm = [&](int M, int N, int K, const double* A, const double* Bm, int trans_b, double* C, int ldc,
                        int ldb, double alpha, double beta) {
            dim3 grid((N + 15) / 16, (M + 15) / 16);
            hp::dgemm_small_kernel<<<grid, 256, 0, h->stream>>>(M, N, K, A, k, Bm, ldb, trans_b, C, ldc, alpha, beta);
            h->launches += 1;
        };
        for (int j = 0; j < r; ++j) {
            const int sj = size(j), oj = off(j);
            const int nt = hp::kTileElems * hp::tri(hp::aug_blocks(sj));
            hp::tiles_from_dense_kernel<<<(nt + 255) / 256, 256, 0, h->stream>>>(B, k, oj, sj, blk_base.as<double>());
            h->launches += 1;
            HY_TRY(factor_block(blk_base.as<double>(), sj, tiles.as<double>()));
            HY_TRY(launch_tri_inverse(tiles.as<double>(), sj, X + (size_t)oj * k + oj, linv_t.as<double>() + (size_t)oj * k + oj, k));
            for (int i = j + 1; i < r; ++i)     // L_ij = B_ij L_jj^-T
                gemm(size(i), sj, sj, B + (size_t)off(i) * k + oj, X + (size_t)oj * k + oj, 1, L + (size_t)off(i) * k + oj, k, k, 1.0, 0.0);
            for (int i = j + 1; i < r; ++i)     // Schur update of the remaining lower blocks
                for (int l = j + 1; l <= i; ++l)
                    gemm(size(i), size(l), sj, L + (size_t)off(i) * k + oj, L + (size_t)off(l) * k + oj, 1,
                         B + (size_t)off(i) * k + off(l), k, k, -1.0, 1.0);
        }
        for (int j = 0; j < r; ++j)             // X_ij = -L_ii^-1 sum_{q=j}^{i-1} L_iq X_qj
            for (int i = j + 1; i < r; ++i) {
                for (int q = j; q < i; ++q)
                    gemm(size(i), size(j), size(q), L + (size_t)off(i) * k + off(q), X + (size_t)off(q) * k + off(j), 0,
                         blk_tmp.as<double>(), bs0, k, 1.0, q == j ? 0.0 : 1.0);
                dim3 grid((size(j) + 15) / 16, (size(i) + 15) / 16);
                hp::dgemm_small_kernel<<<grid, 256, 0, h->stream>>>(size(i), size(j), size(i), X + (size_t)off(i) * k + off(i), k,
                                                                     blk_tmp.as<double>(), bs0, 0, X + (size_t)off(i) * k + off(j),
                                                                     k, -1.0, 0.0);
                h->launches += 1;
            }
        if (r > 1) {
            hp::transpose_dense_kernel<<<(k * k + 255) / 256, 256, 0, h->stream>>>(X, k, linv_t.as<double>());
            h->launches += 1;
        }
        if (sizeof(T) != sizeof(double)) {
            HY_TRY(linv_f.reserve(sizeof(float) * (size_t)k * k));
            HY_TRY(linv_t_f.reserve(sizeof(float) * (size_t)k * k));
            const int64_t cnt = (int64_t)k * k;
            hp::cast_kernel<double, float><<<(unsigned)((cnt + 255) / 256), 256, 0, h->stream>>>(X, cnt, linv_f.as<float>());
            hp::cast_kernel<double, float><<<(unsigned)((cnt + 255) / 256), 256, 0, h->stream>>>(linv_t.as<double>(), cnt, linv_t_f.as<float>());
            h->launches += 2;
        }
        CU_TRY(cudaGetLastError());
        return HYPREC_OK;
    }

    // L^-1 / L^-T in the factor dtype (the float64 originals, or their float32 copies)
    const T* linv_T() const { return sizeof(T) == sizeof(double) ? linv.template as<T>() : linv_f.template as<T>(); }
    const T* linv_t_T() const { return sizeof(T) == sizeof(double) ? linv_t.template as<T>() : linv_t_f.template as<T>(); }

    int launch_tri_inverse(const double* tile_ptr, int dim, double* out, double* out_t, int ld) {
        namespace hp = hyprec;
        const int nbk = (dim + hp::kTile - 1) / hp::kTile;
        const int n_tiles = hp::tri(hp::aug_blocks(dim));
        size_t inv_smem = sizeof(double) * ((size_t)nbk * hp::kTileElems + hp::kTileElems);
        const size_t with_tiles = inv_smem + sizeof(double) * (size_t)hp::kTileElems * n_tiles;
        const int tis = with_tiles <= (size_t)h->max_smem ? 1 : 0;
        if (tis) inv_smem = with_tiles;
        auto inv_kern = hp::tri_inverse_kernel<double>;
        HY_TRY(ensure_dynamic_smem((const void*)inv_kern, (int)inv_smem));
        inv_kern<<<nbk, 64, inv_smem, h->stream>>>(tile_ptr, dim, tis, out, out_t, ld);
        h->launches += 1;
        CU_TRY(cudaGetLastError());
        return HYPREC_OK;
    }

    // Direct-mode float64 factorisation of one dim x dim SPD matrix given as base tiles; tiles exported.
    int factor_block(const double* base_tiles, int dim, double* export_ptr) {
        namespace hp = hyprec;
        const hp::SolveLayout lay = hp::solve_layout(8, dim, 8, true, sizeof(double));
        if (lay.total > (size_t)h->max_smem) return fail(HYPREC_E_UNSUPPORTED, "factor_block: block too large");
        auto kern = hp::als_solve_rows_kernel<double, 8>;
        HY_TRY(ensure_dynamic_smem((const void*)kern, (int)lay.total));
        HY_TRY(dummy_out.reserve(sizeof(double) * (size_t)k));
        HY_TRY(reserve_counters());
        CU_TRY(cudaMemsetAsync(counter.as<int>(), 0, sizeof(int), h->stream));   // slot 0
        hp::SolveArgs<double> a;
        a.indptr = dummy_ptr.as<int64_t>();
        a.indices = csr.indices.template as<int32_t>();
        a.values = nullptr;
        a.fixed = nullptr;
        a.latent = dummy_out.as<double>();
        a.base = base_tiles;
        a.prior = nullptr;
        a.a_scratch = nullptr;
        a.work_counter = counter.as<int>();
        a.status = counter.as<int>() + kStatusSlot;
        a.lambda = 0.0;
        a.k = dim;
        a.row_lo = 0;
        a.row_hi = 1;
        a.chunk_rows = 8;
        a.row_list = nullptr;
        a.n_list = 0;
        a.max_dim = dim;
        a.lowrank = 0;
        a.ymat = nullptr;
        a.ymat_row0 = 0;
        a.export_tiles = export_ptr;
        a.tmem_cols = 0;
        a.prof = nullptr;
        kern<<<1, hp::kSolveThreads, lay.total, h->stream>>>(a);
        h->launches += 1;
        CU_TRY(cudaGetLastError());
        return HYPREC_OK;
    }

    // ---------------------------------------------------------------- multi-GPU exchange (comm.cuh)
    struct Comm {
        bool exported = false, on = false, ipc = true;
        int rank = 0, world = 1;
        void* arena = nullptr;
        size_t bytes = 0, off_u = 0, off_v = 0, off_q[2] = {0, 0}, off_slots[2] = {0, 0}, off_scal = 0, off_flags = 0;
        size_t slot_doubles = 0;
        hyprec::PeerList peers;
        std::vector<int64_t> ub, ib;
        unsigned long long seq = 0;
        int prep_side = -1;             // side whose L^-1 / whitened replica are current (-1: none)
        double prep_lambda = -1.0;
        bool q_ok[2] = {false, false};  // whitened replica of U / V complete on this rank
        bool gram64_ok[2] = {false, false};
        bool slots_have_rmse = false;   // the item side's slots carry (s1, s2) of the current factors
        double comm_ms = 0.0;           // (statistics) bytes this rank stored into peers since attach
        double bytes_pushed = 0.0;
    };
    Comm cm;
    DevBuf gram64[2], slot_local, comm_status;
    static constexpr unsigned long long kCommTimeoutNs = 20ull * 1000000000ull;

    T* q_of(int side) {   // whitened replica of U (0) / V (1)
        if (cm.exported) return reinterpret_cast<T*>(static_cast<unsigned char*>(cm.arena) + cm.off_q[side]);
        return qmat.as<T>();
    }

    int comm_export(int rank, int world, void* blob, size_t blob_bytes) override {
        if (k == 0) return fail(HYPREC_E_INVALID, "hyprec_comm_export before set_factors");
        if (world < 1 || world > hyprec::kMaxRanks || rank < 0 || rank >= world)
            return fail(HYPREC_E_INVALID, "hyprec_comm_export: world must be in [1, 8] and rank inside it");
        if (blob_bytes < sizeof(CommBlob)) return fail(HYPREC_E_INVALID, "hyprec_comm_export: blob too small");
        if (cm.exported) return fail(HYPREC_E_INVALID, "hyprec_comm_export: already exported (hyprec_comm_detach first)");
        auto up = [](size_t x) { return (x + 1023) & ~(size_t)1023; };
        size_t o = 0;
        cm.off_u = o;        o += up(sizeof(T) * (size_t)m * k);
        cm.off_v = o;        o += up(sizeof(T) * (size_t)n * k);
        cm.off_q[0] = o;     o += up(sizeof(T) * (size_t)m * k);
        cm.off_q[1] = o;     o += up(sizeof(T) * (size_t)n * k);
        cm.slot_doubles = (size_t)k * k + 8;
        cm.off_slots[0] = o; o += up(sizeof(double) * cm.slot_doubles * world);
        cm.off_slots[1] = o; o += up(sizeof(double) * cm.slot_doubles * world);
        cm.off_scal = o;     o += up(sizeof(double) * 8 * world);
        cm.off_flags = o;    o += up(sizeof(unsigned long long) * hyprec::kMaxRanks);
        cm.bytes = o;
        CU_TRY(cudaStreamSynchronize(h->stream));
        CU_TRY(cudaMalloc(&cm.arena, cm.bytes));
        unsigned char* base = static_cast<unsigned char*>(cm.arena);
        CU_TRY(cudaMemset(base + cm.off_q[0], 0, cm.bytes - cm.off_q[0]));
        CU_TRY(cudaMemcpy(base + cm.off_u, U.ptr, sizeof(T) * (size_t)m * k, cudaMemcpyDeviceToDevice));
        CU_TRY(cudaMemcpy(base + cm.off_v, V.ptr, sizeof(T) * (size_t)n * k, cudaMemcpyDeviceToDevice));
        U.adopt(base + cm.off_u, sizeof(T) * (size_t)m * k);
        V.adopt(base + cm.off_v, sizeof(T) * (size_t)n * k);
        CommBlob b;
        memset(&b, 0, sizeof(b));
        b.magic = kCommMagic;
        CU_TRY(cudaIpcGetMemHandle(&b.handle, cm.arena));
        b.m = m; b.n = n; b.k = k; b.elem = sizeof(T); b.world = world; b.rank = rank; b.arena_bytes = (long long)cm.bytes;
        memcpy(blob, &b, sizeof(b));
        cm.rank = rank;
        cm.world = world;
        cm.exported = true;
        cm.on = false;
        return HYPREC_OK;
    }

    int comm_attach(const void* blobs, size_t blob_bytes, const int64_t* user_bounds, const int64_t* item_bounds) override {
        if (!cm.exported) return fail(HYPREC_E_INVALID, "hyprec_comm_attach before hyprec_comm_export");
        if (cm.on) return fail(HYPREC_E_INVALID, "hyprec_comm_attach: already attached");
        if (blob_bytes < sizeof(CommBlob) || !blobs || !user_bounds || !item_bounds)
            return fail(HYPREC_E_INVALID, "hyprec_comm_attach: bad arguments");
        if (user_bounds[0] != 0 || user_bounds[cm.world] != m || item_bounds[0] != 0 || item_bounds[cm.world] != n)
            return fail(HYPREC_E_INVALID, "hyprec_comm_attach: bounds must run from 0 to the number of rows");
        for (int r = 0; r < cm.world; ++r)
            if (user_bounds[r + 1] < user_bounds[r] || item_bounds[r + 1] < item_bounds[r])
                return fail(HYPREC_E_INVALID, "hyprec_comm_attach: bounds must be monotone");
        memset(&cm.peers, 0, sizeof(cm.peers));
        cm.peers.world = cm.world;
        cm.peers.rank = cm.rank;
        for (int r = 0; r < cm.world; ++r) {
            CommBlob b;
            memcpy(&b, static_cast<const unsigned char*>(blobs) + (size_t)r * blob_bytes, sizeof(b));
            if (b.magic != kCommMagic || b.rank != r || b.world != cm.world || b.m != m || b.n != n || b.k != k ||
                b.elem != (long long)sizeof(T) || b.arena_bytes != (long long)cm.bytes)
                return fail(HYPREC_E_INVALID, "hyprec_comm_attach: rank " + std::to_string(r) + " published a different problem");
            if (r == cm.rank) {
                cm.peers.base[r] = cm.arena;
            } else {
                void* ptr = nullptr;
                cudaError_t e = cudaIpcOpenMemHandle(&ptr, b.handle, cudaIpcMemLazyEnablePeerAccess);
                if (e != cudaSuccess) {
                    for (int q = 0; q < r; ++q)
                        if (q != cm.rank && cm.peers.base[q]) cudaIpcCloseMemHandle(cm.peers.base[q]);
                    cudaGetLastError();
                    return fail(HYPREC_E_CUDA, std::string("cudaIpcOpenMemHandle (peer ") + std::to_string(r) + "): " + cudaGetErrorString(e));
                }
                cm.peers.base[r] = ptr;
            }
        }
        cm.ipc = true;
        return finish_attach(user_bounds, item_bounds);
    }

    // Peers that live in this process (one handle per GPU driven by one host process, or several handles on one
    // GPU in the tests): their arenas are plain device pointers, no IPC mapping.
    int comm_attach_ptrs(void* const* bases, const int* devices, const int64_t* user_bounds, const int64_t* item_bounds) override {
        if (!cm.exported) return fail(HYPREC_E_INVALID, "hyprec_comm_attach_local before hyprec_comm_export");
        if (cm.on) return fail(HYPREC_E_INVALID, "hyprec_comm_attach_local: already attached");
        if (!bases || !devices || !user_bounds || !item_bounds) return fail(HYPREC_E_INVALID, "hyprec_comm_attach_local: bad arguments");
        if (user_bounds[0] != 0 || user_bounds[cm.world] != m || item_bounds[0] != 0 || item_bounds[cm.world] != n)
            return fail(HYPREC_E_INVALID, "hyprec_comm_attach_local: bounds must run from 0 to the number of rows");
        for (int r = 0; r < cm.world; ++r)
            if (user_bounds[r + 1] < user_bounds[r] || item_bounds[r + 1] < item_bounds[r])
                return fail(HYPREC_E_INVALID, "hyprec_comm_attach_local: bounds must be monotone");
        memset(&cm.peers, 0, sizeof(cm.peers));
        cm.peers.world = cm.world;
        cm.peers.rank = cm.rank;
        for (int r = 0; r < cm.world; ++r) {
            if (!bases[r]) return fail(HYPREC_E_INVALID, "hyprec_comm_attach_local: null arena");
            if (r == cm.rank && bases[r] != cm.arena) return fail(HYPREC_E_INVALID, "hyprec_comm_attach_local: own arena mismatch");
            if (devices[r] != h->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(devices[r], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return fail(HYPREC_E_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
                cudaGetLastError();
            }
            cm.peers.base[r] = bases[r];
        }
        cm.ipc = false;
        return finish_attach(user_bounds, item_bounds);
    }

    int comm_arena(void** base, size_t* bytes) override {
        if (!cm.exported) return fail(HYPREC_E_INVALID, "hyprec_comm_arena before hyprec_comm_export");
        if (base) *base = cm.arena;
        if (bytes) *bytes = cm.bytes;
        return HYPREC_OK;
    }

    int finish_attach(const int64_t* user_bounds, const int64_t* item_bounds) {
        cm.ub.assign(user_bounds, user_bounds + cm.world + 1);
        cm.ib.assign(item_bounds, item_bounds + cm.world + 1);
        h->user_lo = cm.ub[cm.rank]; h->user_hi = cm.ub[cm.rank + 1];
        h->item_lo = cm.ib[cm.rank]; h->item_hi = cm.ib[cm.rank + 1];
        HY_TRY(gram64[0].reserve(sizeof(double) * (size_t)k * k));
        HY_TRY(gram64[1].reserve(sizeof(double) * (size_t)k * k));
        HY_TRY(slot_local.reserve(sizeof(double) * cm.slot_doubles));
        HY_TRY(comm_status.reserve(sizeof(int) * 4));
        CU_TRY(cudaMemsetAsync(comm_status.ptr, 0, sizeof(int) * 4, h->stream));
        CU_TRY(cudaStreamSynchronize(h->stream));
        cm.seq = 0;
        cm.prep_side = -1;
        cm.q_ok[0] = cm.q_ok[1] = false;
        cm.gram64_ok[0] = cm.gram64_ok[1] = false;
        cm.slots_have_rmse = false;
        cm.bytes_pushed = 0.0;
        plan[0].lo = plan[1].lo = -1;
        cm.on = true;
        return HYPREC_OK;
    }

    int comm_detach() override {
        if (!cm.exported) return HYPREC_OK;
        cudaStreamSynchronize(h->stream);
        if (cm.on && cm.ipc)
            for (int r = 0; r < cm.world; ++r)
                if (r != cm.rank && cm.peers.base[r]) cudaIpcCloseMemHandle(cm.peers.base[r]);
        cm.on = false;
        // the factor matrices move back into buffers of their own (the arena is freed; peers must have detached
        // -- the caller puts a process barrier between the ranks' detach calls and anything that reuses memory)
        DevBuf nu, nv;
        HY_TRY(nu.reserve(sizeof(T) * (size_t)m * k));
        HY_TRY(nv.reserve(sizeof(T) * (size_t)n * k));
        CU_TRY(cudaMemcpy(nu.ptr, U.ptr, sizeof(T) * (size_t)m * k, cudaMemcpyDeviceToDevice));
        CU_TRY(cudaMemcpy(nv.ptr, V.ptr, sizeof(T) * (size_t)n * k, cudaMemcpyDeviceToDevice));
        U.release(); V.release();
        U = nu; V = nv;
        cudaFree(cm.arena);
        cm.arena = nullptr;
        cm.exported = false;
        h->user_lo = 0; h->user_hi = -1; h->item_lo = 0; h->item_hi = -1;
        gram_ok[0] = gram_ok[1] = false;
        cross_valid = false;
        plan[0].lo = plan[1].lo = -1;
        return HYPREC_OK;
    }

    bool comm_active() const override { return cm.on; }

    int comm_stats(double* out8) override {
        for (int i = 0; i < 8; ++i) out8[i] = 0.0;
        out8[0] = cm.on ? 1.0 : 0.0;
        out8[1] = cm.bytes_pushed;
        out8[2] = (double)cm.seq;
        out8[3] = (double)cm.bytes;
        return HYPREC_OK;
    }

    // every store this rank issued so far is visible to its peers; every peer has done the same
    int comm_barrier() {
        ++cm.seq;
        hyprec::comm_signal_kernel<<<1, 32, 0, h->stream>>>(cm.peers, cm.off_flags, cm.seq);
        hyprec::comm_wait_kernel<<<1, 32, 0, h->stream>>>(cm.peers, cm.off_flags, cm.seq, kCommTimeoutNs, comm_status.as<int>());
        h->launches += 2;
        CU_TRY(cudaGetLastError());
        return HYPREC_OK;
    }

    hyprec::PeerList other_peers_as_list() const { return cm.peers; }

    // rows [lo, hi) or a row list of U (side 0) / V (side 1): local replica -> every other replica
    int comm_push_rows(int side, int64_t lo, int64_t n_rows, const int32_t* list) {
        return comm_push_at(side == 0 ? cm.off_u : cm.off_v, lo, n_rows, list);
    }

    int comm_push_at(size_t off_matrix, int64_t lo, int64_t n_rows, const int32_t* list) {
        if (n_rows <= 0 || cm.world == 1) return HYPREC_OK;
        const int grid = (int)std::min<int64_t>((n_rows + 7) / 8, (int64_t)h->num_sms * 8);
        hyprec::push_rows_kernel<T><<<grid, 256, 0, h->stream>>>(cm.peers, off_matrix, k, lo, n_rows, list);
        h->launches += 1;
        cm.bytes_pushed += (double)n_rows * k * sizeof(T) * (cm.world - 1);
        CU_TRY(cudaGetLastError());
        return HYPREC_OK;
    }

    // peer destinations (other ranks' replicas at arena offset `off`) for the GEMM epilogues
    int peer_outputs(size_t off, void** out7) const {
        int c = 0;
        for (int r = 0; r < cm.world; ++r)
            if (r != cm.rank) out7[c++] = static_cast<unsigned char*>(cm.peers.base[r]) + off;
        return c;
    }

    // Whitening state of side `f` without help from the peers (first half-step after set_factors, or a new
    // lambda): float64 Gram of the whole replica, L^-1, whitened replica -- every rank computes the same.
    int comm_prep_local(int f, double lambda, bool lowrank) {
        const T* y = f == 0 ? U.as<T>() : V.as<T>();
        const int64_t rows = f == 0 ? m : n;
        if (!cm.gram64_ok[f]) {
            Timer t(h, HYPREC_T_GRAM);
            HY_TRY(gram64_rows(y, 0, rows, gram64[f].as<double>()));
            HY_TRY(gram[f].reserve(sizeof(T) * (size_t)k * k));
            if constexpr (sizeof(T) == sizeof(double)) {
                CU_TRY(cudaMemcpyAsync(gram[f].ptr, gram64[f].ptr, sizeof(double) * (size_t)k * k, cudaMemcpyDeviceToDevice, h->stream));
            } else {
                const int64_t cnt = (int64_t)k * k;
                hyprec::cast_kernel<double, T><<<(unsigned)((cnt + 255) / 256), 256, 0, h->stream>>>(gram64[f].as<double>(), cnt, gram[f].template as<T>());
                h->launches += 1;
            }
            cm.gram64_ok[f] = true;
            gram_ok[f] = true;
        }
        if (lowrank && !(cm.prep_side == f && cm.prep_lambda == lambda && cm.q_ok[f])) {
            Timer t(h, HYPREC_T_PREP);
            HY_TRY(linv_from_gram(gram64[f].as<double>(), lambda));
            HY_TRY(whiten_gemm(y, rows, linv_T(), q_of(f), nullptr));
            cm.prep_side = f;
            cm.prep_lambda = lambda;
            cm.q_ok[f] = true;
        }
        return HYPREC_OK;
    }

    // After this rank solved rows [lo, hi) of side s: partial float64 Gram of those rows (+ the RMSE sums when the
    // item half-step produced them) into every rank's mailbox, barrier, fixed-order reduction; then L^-1 of the new
    // Gram, and the whitened rows of this shard stored into every replica from the GEMM epilogue; barrier.
    int comm_post(int s, int64_t lo, int64_t hi, double lambda, bool lowrank, bool have_cross) {
        T* y = s == 0 ? U.as<T>() : V.as<T>();
        const int kk = k * k;
        double* slot = slot_local.as<double>();
        {
            Timer t(h, HYPREC_T_GRAM);
            HY_TRY(gram64_rows(y, lo, hi, slot));
        }
        CU_TRY(cudaMemsetAsync(slot + kk, 0, sizeof(double) * 8, h->stream));
        if (s == 1 && have_cross) {
            // sum over this rank's item rows of sum_j r_j (u_j . v_i) and of r^2 (fixed order)
            Timer t(h, HYPREC_T_RMSE);
            HY_TRY(row_s2.reserve(sizeof(double) * std::max<int64_t>(m, 1024)));
            if (hi > lo) {
                hyprec::reduce_sum_kernel<<<1, 1024, 0, h->stream>>>(cross.as<double>() + lo, hi - lo, slot + kk);
                const int64_t e0 = csc.h_indptr[lo], cnt = csc.h_indptr[hi] - e0;
                const int nblk = (int)std::min<int64_t>(1024, (cnt + 4095) / 4096);
                if (nblk > 0) {
                    hyprec::sumsq_partial_kernel<T><<<nblk, 256, 0, h->stream>>>(csc.values.template as<T>() + e0, cnt, row_s2.as<double>());
                    hyprec::reduce_sum_kernel<<<1, 1024, 0, h->stream>>>(row_s2.as<double>(), nblk, slot + kk + 1);
                }
                h->launches += 3;
            }
            hyprec::set_double_kernel<<<1, 1, 0, h->stream>>>(slot + kk + 2, 1.0);
            h->launches += 1;
        }
        {
            Timer t(h, HYPREC_T_COMM);
            hyprec::push_slot_kernel<<<8, 256, 0, h->stream>>>(cm.peers, cm.off_slots[s], cm.slot_doubles, slot, kk + 8);
            h->launches += 1;
            cm.bytes_pushed += (double)(kk + 8) * 8.0 * (cm.world - 1);
            HY_TRY(comm_barrier());
            const double* slots = reinterpret_cast<const double*>(static_cast<unsigned char*>(cm.arena) + cm.off_slots[s]);
            HY_TRY(gram[s].reserve(sizeof(T) * (size_t)k * k));
            hyprec::reduce_slots_kernel<T><<<(kk + 255) / 256, 256, 0, h->stream>>>(slots, cm.slot_doubles, cm.world, kk,
                                                                                  gram64[s].as<double>(), gram[s].template as<T>());
            h->launches += 1;
            CU_TRY(cudaGetLastError());
        }
        cm.gram64_ok[s] = true;
        gram_ok[s] = true;
        if (s == 1) cm.slots_have_rmse = true;   // whether every rank's slot is valid is read from the slots themselves
        if (lowrank) {
            Timer t(h, HYPREC_T_PREP);
            HY_TRY(linv_from_gram(gram64[s].as<double>(), lambda));
            // whitened rows of this shard -> every replica (the all-gather is the GEMM's epilogue)
            if (hi > lo) {
                void* peers7[7];
                const int np = peer_outputs(cm.off_q[s] + sizeof(T) * (size_t)lo * k, peers7);
                if (h->fused_push) {
                    HY_TRY(whiten_gemm(y + lo * k, hi - lo, linv_T(), q_of(s) + lo * k, nullptr, peers7, np));
                    cm.bytes_pushed += (double)(hi - lo) * k * sizeof(T) * np;
                } else {
                    HY_TRY(whiten_gemm(y + lo * k, hi - lo, linv_T(), q_of(s) + lo * k, nullptr));
                    HY_TRY(comm_push_at(cm.off_q[s], lo, hi - lo, nullptr));
                }
            }
            cm.prep_side = s;
            cm.prep_lambda = lambda;
            cm.q_ok[s] = true;
        } else {
            cm.q_ok[s] = false;
        }
        {
            Timer t(h, HYPREC_T_COMM);
            HY_TRY(comm_barrier());
        }
        return HYPREC_OK;
    }

    int half_step(int side, double lambda) override {
        if (k == 0) return fail(HYPREC_E_INVALID, "als_half_step before set_factors");
        if (side != HYPREC_SIDE_USER && side != HYPREC_SIDE_ITEM) return fail(HYPREC_E_INVALID, "side must be 0 or 1");
        const bool user = side == HYPREC_SIDE_USER;
        const bool comm = cm.on;
        const int fside = user ? 1 : 0;          // the fixed side's index (0: U, 1: V)
        const int sside = user ? 0 : 1;          // the side being solved
        const int64_t total = user ? m : n;
        const int64_t n_fixed = user ? n : m;
        int64_t lo = user ? h->user_lo : h->item_lo, hi = user ? h->user_hi : h->item_hi;
        if (hi < 0) { lo = 0; hi = total; }
        if (lo < 0 || hi > total || lo > hi) return fail(HYPREC_E_INVALID, "shard outside the matrix");

        // The low-rank path pays off when rows have fewer positives than k.  Its whitening
        // (L^-1 of 0.1 G + lambda I) is always computed in float64; the deg x deg systems themselves
        // are well conditioned and run in the factor dtype.
        const bool lowrank = h->lowrank_enabled && lambda > 0.0 && k >= 16;
        HY_TRY(reserve_counters());
        if (comm) {
            // every replica this half-step stores into is idle on its owner (evaluation, uploads)
            {
                Timer t(h, HYPREC_T_COMM);
                HY_TRY(comm_barrier());
            }
            cm.slots_have_rmse = false;
            HY_TRY(comm_prep_local(fside, lambda, lowrank));   // no-op when the previous half-step left it ready
        } else {
            HY_TRY(compute_gram(fside));   // Gram of the FIXED side
        }
        const int nb = hyprec::aug_blocks(k), n_tiles = hyprec::tri(nb);
        HY_TRY(base.reserve(sizeof(T) * (size_t)hyprec::kTileElems * n_tiles));
        {
            const int total_e = hyprec::kTileElems * n_tiles;
            hyprec::build_base_tiles_kernel<T><<<(total_e + 255) / 256, 256, 0, h->stream>>>(
                gram[fside].as<T>(), k, (T)lambda, base.as<T>());
            h->launches += 1;
        }
        if (hi == lo && !comm) return HYPREC_OK;
        CU_TRY(cudaMemsetAsync(counter.ptr, 0, 15 * sizeof(int), h->stream));   // (the status slot keeps what earlier steps raised)
        const T* fixed_ptr = user ? V.as<T>() : U.as<T>();
        const T* prior_ptr = (!user && has_prior) ? prior.as<T>() : nullptr;
        HY_TRY(ymat.reserve(sizeof(T)));

        // float32 k x k systems (64 <= k <= 256) accumulate sum y y^T on the tensor cores (tc_syrk.cuh)
        const bool tc_direct = h->tc_solve_enabled && sizeof(T) == sizeof(float) && k >= 64 && k <= 256;
        const bool chol_direct = tc_direct && h->tc_chol_enabled && k % 4 == 0;
        // the item half-step's tensor-memory kernels hand out the RMSE cross term per row (see `cross`)
        cross_valid = false;
        cross_active = false;
        const bool cross_wanted = !user && !has_prior && h->rmse_fused_enabled;
        auto begin_cross = [&]() -> int {
            HY_TRY(cross.reserve(sizeof(double) * (size_t)n));
            if (hi > lo) CU_TRY(cudaMemsetAsync(cross.as<double>() + lo, 0, sizeof(double) * (size_t)(hi - lo), h->stream));
            cross_active = true;
            return HYPREC_OK;
        };
        void* peers7[7];
        const int n_peer_latent = comm ? peer_outputs(user ? cm.off_u : cm.off_v, peers7) : 0;
        // (evaluated alike on every rank: a sharded run sums the ranks' parts, so all take the same branch)
        // (every row kernel -- the SIMT tile kernel in both precisions and the tensor-memory kernels -- hands the term out)
        const bool cross_cfg = cross_wanted;
        if (hi == lo) {
            // (comm, empty shard) nothing to solve; still takes part in the exchange below
            if (cross_cfg) HY_TRY(begin_cross());
        } else if (!lowrank) {
            if (cross_cfg) HY_TRY(begin_cross());
            {
                Timer t(h, HYPREC_T_SOLVE);
                if constexpr (sizeof(T) == sizeof(float)) {
                    if (chol_direct) {
                        HY_TRY(build_plan(user, lo, hi, -1, h->split_enabled));
                        const RowPlan& pl = plan[user ? 0 : 1];
                        HY_TRY(launch_heavy(user, lambda, lo, hi, pl, fixed_ptr, prior_ptr, h->stream, 0));
                    }
                }
                if (!chol_direct)
                    HY_TRY(launch_rows(user, lambda, lo, hi, nullptr, 0, false, k, hyprec::kSolveThreads, fixed_ptr, prior_ptr,
                                       nullptr, nullptr, hi - lo, h->stream, 0, 8, tc_direct));
            }
            if (comm) {
                Timer t(h, HYPREC_T_COMM);
                HY_TRY(comm_push_rows(sside, lo, hi - lo, nullptr));
            }
        } else {
            const int cap_fits = std::min(k, kBucketDim[usable_buckets() - 1]);
            int cap = cap_fits;
            // float64: rows of 128..191 positives cost ~350 us in low-rank form against ~165 us as direct k x k rows
            // when the k x k tiles fit shared memory (measured on B200, profiles/r02_lowrank_cap.md)
            if (sizeof(T) == sizeof(double) && hyprec::solve_layout(8, k, 8, true, sizeof(double)).total <= (size_t)h->max_smem)
                cap = std::min(cap, 127);
            // float32 with the tensor-memory k x k kernel: rows of 192..255 positives are 7 % cheaper as direct rows
            // (no second pass over the gathered rows, K = deg instead of k in the accumulation; measured 141.9 ->
            // 139.2 ms per epoch at a quarter of the scaled shape)
            if (chol_direct && k > 191) cap = std::min(cap, 191);
            if (h->lowrank_cap > 0) cap = std::min(cap_fits, h->lowrank_cap);   // tuning knob (overrides the two rules above)
            HY_TRY(build_plan(user, lo, hi, cap, chol_direct && h->split_enabled));
            const RowPlan& pl = plan[user ? 0 : 1];
            const int nbk = pl.n_buckets;
            const int32_t* lists = pl.lists.template as<int32_t>();
            const int64_t n_light = pl.bucket_off[nbk];
            const int64_t n_heavy = pl.bucket_off[nbk + 1] - pl.bucket_off[nbk];
            HY_TRY(ymat.reserve(sizeof(T) * (size_t)std::max<int64_t>(n_light, 1) * k));
            // The fused cross term is used only under conditions every rank evaluates alike (a sharded run
            // sums the ranks' parts): tensor-memory kernels for every bucket that can occur at this k.
            if (cross_cfg) HY_TRY(begin_cross());
            // rows above the cap do not depend on the factor of B: start them right away on their
            // own stream (base tiles and counters are ready on the main stream)
            CU_TRY(cudaEventRecord(h->ev_base, h->stream));
            if (n_heavy > 0) {
                CU_TRY(cudaStreamWaitEvent(h->aux[kHeavyStream], h->ev_base, 0));
                if constexpr (sizeof(T) == sizeof(float)) {
                    if (chol_direct)
                        HY_TRY(launch_heavy(user, lambda, lo, hi, pl, fixed_ptr, prior_ptr, h->aux[kHeavyStream], 1 + kMaxBuckets));
                }
                if (!chol_direct)
                    HY_TRY(launch_rows(user, lambda, lo, hi, lists + pl.bucket_off[nbk], n_heavy, false, k,
                                       hyprec::kSolveThreads, fixed_ptr, prior_ptr, nullptr, nullptr, n_heavy,
                                       h->aux[kHeavyStream], 1 + kMaxBuckets, 8, tc_direct));
                CU_TRY(cudaEventRecord(h->ev_join[kHeavyStream], h->aux[kHeavyStream]));
            }
            if (n_light > 0) {
                {
                    Timer t(h, HYPREC_T_PREP);
                    if (!comm) {
                        // 1-2. factor B = 0.1 G + lambda I = L L^T once and form L^-1, L^-T
                        HY_TRY(compute_linv(user, lambda, fixed_ptr, n_fixed));
                        // 3. whitened factors Q = Y L^-T
                        HY_TRY(qmat.reserve(sizeof(T) * (size_t)n_fixed * k));
                        HY_TRY(whiten_gemm(fixed_ptr, n_fixed, linv_T(), qmat.as<T>(), nullptr));
                    }
                    // (theta L^-T for the item prior)
                    if (prior_ptr) {
                        HY_TRY(pq.reserve(sizeof(T) * (size_t)total * k));
                        HY_TRY(whiten_gemm(prior_ptr, total, linv_T(), pq.as<T>(), nullptr));
                    }
                }
                const T* q_ptr = q_of(fside);
                {
                    // 4. deg x deg systems, one launch per degree bucket, all buckets concurrently
                    Timer t(h, HYPREC_T_SOLVE);
                    // (small systems on 4 x 4 tiles: more threads with a tile, shorter FMA chains)
                    CU_TRY(cudaEventRecord(h->ev_fork, h->stream));
                    for (int b = 0; b < nbk; ++b) {
                        const int64_t cnt = pl.bucket_off[b + 1] - pl.bucket_off[b];
                        if (cnt == 0) continue;
                        CU_TRY(cudaStreamWaitEvent(h->aux[b], h->ev_fork, 0));
                        // float32: systems of 64+ rows build S on the tensor cores (tc_syrk.cuh), and with
                        // HYPREC_TC_CHOL (default) every bucket up to 256 rows also factorises in tensor
                        // memory -- the small ones several per CTA pass (tc_chol_multi.cuh)
                        const bool f32 = sizeof(T) == sizeof(float);
                        const bool tc = h->tc_solve_enabled && f32 && h->bucket_tile[b] == 8 && k % 4 == 0 && pl.bucket_dim[b] <= 256;
                        const bool chol = tc && h->tc_chol_enabled && pl.bucket_dim[b] > 128;
                        // rows of <= 32 positives: one warp per row, float64 arithmetic in both precisions (small_f64.cuh)
                        const bool small = h->small_f64_enabled && pl.bucket_dim[b] <= hyprec::kSmallDim && prior_ptr == nullptr;
                        const bool multi = !small && h->tc_chol_enabled && f32 && k % 4 == 0 && pl.bucket_dim[b] <= 128 && h->tc_multi_enabled;
                        if (small) {
                            HY_TRY(launch_small(user, lists + pl.bucket_off[b], cnt, q_ptr, h->aux[b], 1 + b, pl.bucket_off[b]));
                        } else if constexpr (sizeof(T) == sizeof(float)) {
                            if (multi) {
                                const int slot_rows = pl.bucket_dim[b] <= 32 ? 32 : (pl.bucket_dim[b] <= 64 ? 64 : 128);
                                HY_TRY(launch_chol_multi(user, lambda, lo, hi, lists + pl.bucket_off[b], cnt, slot_rows,
                                                         reinterpret_cast<const float*>(q_ptr),
                                                         prior_ptr ? pq.as<float>() : nullptr, cnt, h->aux[b], 1 + b, pl.bucket_off[b]));
                            } else if (chol) {
                                HY_TRY(launch_chol_rows(user, lambda, lo, hi, lists + pl.bucket_off[b], cnt, true, pl.bucket_dim[b],
                                                        reinterpret_cast<const float*>(q_ptr), prior_ptr ? pq.as<float>() : nullptr, cnt,
                                                        h->aux[b], 1 + b, pl.bucket_off[b]));
                            }
                        }
                        if (!chol && !multi && !small)
                        HY_TRY(launch_rows(user, lambda, lo, hi, lists + pl.bucket_off[b], cnt, true, pl.bucket_dim[b],
                                           tc ? hyprec::kSolveThreads : h->bucket_threads[b], q_ptr,
                                           prior_ptr ? pq.as<T>() : nullptr, nullptr, nullptr, cnt, h->aux[b], 1 + b,
                                           h->bucket_tile[b], tc, pl.bucket_off[b]));
                        CU_TRY(cudaEventRecord(h->ev_join[b], h->aux[b]));
                        CU_TRY(cudaStreamWaitEvent(h->stream, h->ev_join[b], 0));
                    }
                }
                {
                    // 5. x = L^-T y for the low-rank rows: one GEMM straight into the factor matrix -- and, in a
                    // multi-GPU run, into every other rank's replica (the exchange is this GEMM's epilogue)
                    Timer t(h, HYPREC_T_PREP);
                    T* latent = user ? U.as<T>() : V.as<T>();
                    if (h->fused_push || !comm) {
                        HY_TRY(whiten_gemm(ymat.as<T>(), n_light, linv_t_T(), latent, lists, peers7, n_peer_latent));
                        if (comm) cm.bytes_pushed += (double)n_light * k * sizeof(T) * n_peer_latent;
                    } else {
                        HY_TRY(whiten_gemm(ymat.as<T>(), n_light, linv_t_T(), latent, lists));
                        HY_TRY(comm_push_rows(sside, 0, n_light, lists));
                    }
                }
            }
            if (n_heavy > 0) {
                Timer t(h, HYPREC_T_SOLVE);   // whatever of the heavy rows' launch is still running
                CU_TRY(cudaStreamWaitEvent(h->stream, h->ev_join[kHeavyStream], 0));
            }
            if (comm && n_heavy > 0) {
                Timer t(h, HYPREC_T_COMM);
                HY_TRY(comm_push_rows(sside, 0, n_heavy, lists + pl.bucket_off[nbk]));
            }
        }
        gram_ok[sside] = false;
        const bool have_cross = cross_active;
        if (cross_active) {
            cross_active = false;
            cross_valid = true;
            cross_lo = lo;
            cross_hi = hi;
        }
        if (comm) {
            cm.gram64_ok[sside] = false;
            cm.q_ok[sside] = false;
            HY_TRY(comm_post(sside, lo, hi, lambda, lowrank, have_cross));
            return HYPREC_OK;    // status words are read with the RMSE (no host synchronisation inside an epoch)
        }
        return read_status("als_half_step");
    }

    int reserve_counters() {
        if (counter.ptr != nullptr) return HYPREC_OK;
        HY_TRY(counter.reserve(16 * sizeof(int)));
        CU_TRY(cudaMemsetAsync(counter.ptr, 0, 16 * sizeof(int), h->stream));
        return HYPREC_OK;
    }

    // status slot of the row kernels (1: non-finite solution) and of the exchange (2: a peer never arrived)
    int read_status(const char* where) {
        int status[16] = {0};
        int cstat[4] = {0};
        CU_TRY(cudaMemcpyAsync(status, counter.ptr, sizeof(status), cudaMemcpyDeviceToHost, h->stream));
        if (cm.on) CU_TRY(cudaMemcpyAsync(cstat, comm_status.ptr, sizeof(cstat), cudaMemcpyDeviceToHost, h->stream));
        CU_TRY(cudaStreamSynchronize(h->stream));
        if (cstat[0] != 0)
            return fail(HYPREC_E_CUDA, std::string(where) + ": a peer GPU did not reach the exchange barrier within 20 s");
        if (status[kStatusSlot] != 0) {
            CU_TRY(cudaMemsetAsync(counter.as<int>() + kStatusSlot, 0, sizeof(int), h->stream));
            return fail(HYPREC_E_SINGULAR, std::string(where) + ": non-finite solution (singular normal matrix; lambda == 0?)");
        }
        return HYPREC_OK;
    }

    // ---------------------------------------------------------------- RMSE
    // out3 = { <U^T U, V^T V>_F, sum_nz r (u.v), sum_nz r^2 }; the two sparse sums cover the
    // handle's user shard only when shard_only (the caller all-reduces them across ranks).
    // Sharded run with the exchange attached: the Gram matrices are the reduced mailbox sums (identical on
    // every rank), the sparse sums are reduced over the ranks in rank order -- from the item half-step's
    // mailbox when every rank's slot carries them, else from a pass over this rank's users plus one more
    // mailbox round.  out3 holds GLOBAL sums: every rank computes the same RMSE bit for bit.
    int rmse_terms_comm(double* out3) {
        HY_TRY(reserve_counters());
        for (int f = 0; f < 2; ++f)
            if (!cm.gram64_ok[f]) HY_TRY(comm_prep_local(f, cm.prep_lambda, false));
        HY_TRY(scalars.reserve(sizeof(double) * 8));
        CU_TRY(cudaMemsetAsync(scalars.ptr, 0, sizeof(double) * 8, h->stream));
        const int kk = k * k;
        const double* slots_v = reinterpret_cast<const double*>(static_cast<unsigned char*>(cm.arena) + cm.off_slots[1]);
        {
            Timer t(h, HYPREC_T_RMSE);
            hyprec::frob_inner_kernel<double><<<1, 1024, 0, h->stream>>>(gram64[0].as<double>(), gram64[1].as<double>(), kk, scalars.as<double>());
            // scalars[1..3] = sum over ranks of (s1, s2, valid)
            hyprec::reduce_slots_kernel<double><<<1, 32, 0, h->stream>>>(slots_v + kk, cm.slot_doubles, cm.world, 3,
                                                                        scalars.as<double>() + 1, nullptr);
            h->launches += 2;
        }
        double host[4] = {0, 0, 0, 0};
        bool fused = false;
        if (cm.slots_have_rmse) {
            CU_TRY(cudaMemcpyAsync(host, scalars.ptr, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            HY_TRY(read_status("rmse"));
            fused = host[3] == (double)cm.world;      // the same slots on every rank: the same decision everywhere
        }
        if (!fused) {
            const int64_t lo = h->user_lo, hi = h->user_hi;
            HY_TRY(row_s1.reserve(sizeof(double) * m));
            HY_TRY(row_s2.reserve(sizeof(double) * std::max<int64_t>(m, 1024)));
            double* slot = slot_local.as<double>();
            CU_TRY(cudaMemsetAsync(slot, 0, sizeof(double) * 8, h->stream));
            {
                Timer t(h, HYPREC_T_RMSE);
                if (hi > lo) {
                    const int wpb = 8;
                    const int64_t rows = hi - lo;
                    hyprec::sddmm_row_kernel<T><<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, h->stream>>>(
                        csr.indptr.template as<int64_t>() + lo, csr.indices.template as<int32_t>(),
                        csr.values.template as<T>(), U.as<T>() + lo * k, V.as<T>(), k, rows, row_s1.as<double>(),
                        row_s2.as<double>());
                    hyprec::reduce_sum_kernel<<<1, 1024, 0, h->stream>>>(row_s1.as<double>(), rows, slot);
                    hyprec::reduce_sum_kernel<<<1, 1024, 0, h->stream>>>(row_s2.as<double>(), rows, slot + 1);
                    h->launches += 3;
                }
            }
            {
                Timer t(h, HYPREC_T_COMM);
                hyprec::push_slot_kernel<<<1, 32, 0, h->stream>>>(cm.peers, cm.off_scal, 8, slot, 8);
                h->launches += 1;
                HY_TRY(comm_barrier());
                const double* sc = reinterpret_cast<const double*>(static_cast<unsigned char*>(cm.arena) + cm.off_scal);
                hyprec::reduce_slots_kernel<double><<<1, 32, 0, h->stream>>>(sc, 8, cm.world, 2, scalars.as<double>() + 1, nullptr);
                h->launches += 1;
                // nobody may overwrite the scalar mailbox before every rank has read it
                HY_TRY(comm_barrier());
            }
            CU_TRY(cudaMemcpyAsync(host, scalars.ptr, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            HY_TRY(read_status("rmse"));
        }
        out3[0] = host[0];
        out3[1] = host[1];
        out3[2] = host[2];
        return HYPREC_OK;
    }

    int rmse_terms(double* out3, bool shard_only) override {
        if (k == 0) return fail(HYPREC_E_INVALID, "rmse before set_factors");
        if (cm.on) return rmse_terms_comm(out3);
        int64_t lo = 0, hi = m;
        if (shard_only && h->user_hi >= 0) { lo = h->user_lo; hi = h->user_hi; }
        HY_TRY(compute_gram(0));
        HY_TRY(compute_gram(1));
        HY_TRY(row_s1.reserve(sizeof(double) * m));
        HY_TRY(row_s2.reserve(sizeof(double) * std::max<int64_t>(m, 1024)));
        HY_TRY(scalars.reserve(sizeof(double) * 4));
        CU_TRY(cudaMemsetAsync(scalars.ptr, 0, sizeof(double) * 4, h->stream));
        {
            Timer t(h, HYPREC_T_RMSE);
            hyprec::frob_inner_kernel<T><<<1, 1024, 0, h->stream>>>(gram[0].as<T>(), gram[1].as<T>(), k * k,
                                                                     scalars.as<double>());
            h->launches += 1;
            int64_t ilo = 0, ihi = n;
            if (shard_only && h->item_hi >= 0) { ilo = h->item_lo; ihi = h->item_hi; }
            if (cross_valid && cross_lo == ilo && cross_hi == ihi) {
                // the item half-step left sum_j r_j (u_j . v_i) per item row: no pass over the positives
                if (ihi > ilo)
                    hyprec::reduce_sum_kernel<<<1, 1024, 0, h->stream>>>(cross.as<double>() + ilo, ihi - ilo, scalars.as<double>() + 1);
                if (hi > lo) {
                    const int64_t e0 = csr.h_indptr[lo], cnt = csr.h_indptr[hi] - e0;
                    const int nblk = (int)std::min<int64_t>(1024, (cnt + 4095) / 4096);
                    if (nblk > 0) {
                        hyprec::sumsq_partial_kernel<T><<<nblk, 256, 0, h->stream>>>(csr.values.template as<T>() + e0, cnt, row_s2.as<double>());
                        hyprec::reduce_sum_kernel<<<1, 1024, 0, h->stream>>>(row_s2.as<double>(), nblk, scalars.as<double>() + 2);
                    }
                }
                h->launches += 3;
            } else if (hi > lo) {
                const int wpb = 8;
                const int64_t rows = hi - lo;
                hyprec::sddmm_row_kernel<T><<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, h->stream>>>(
                    csr.indptr.template as<int64_t>() + lo, csr.indices.template as<int32_t>(),
                    csr.values.template as<T>(), U.as<T>() + lo * k, V.as<T>(), k, rows, row_s1.as<double>(),
                    row_s2.as<double>());
                hyprec::reduce_sum_kernel<<<1, 1024, 0, h->stream>>>(row_s1.as<double>(), rows, scalars.as<double>() + 1);
                hyprec::reduce_sum_kernel<<<1, 1024, 0, h->stream>>>(row_s2.as<double>(), rows, scalars.as<double>() + 2);
                h->launches += 3;
            }
        }
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaMemcpyAsync(out3, scalars.ptr, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CU_TRY(cudaStreamSynchronize(h->stream));
        return HYPREC_OK;
    }

    int rmse(double* out) override {
        double s[3];
        HY_TRY(rmse_terms(s, false));
        double rss = s[0] - 2.0 * s[1] + s[2];
        if (rss < 0.0) rss = 0.0;
        *out = sqrt(rss / ((double)m * (double)n));
        return HYPREC_OK;
    }

    // ---------------------------------------------------------------- dense prediction
    int predict_dense(double* out) override {
        if (k == 0) return fail(HYPREC_E_INVALID, "predict_dense before set_factors");
        if (!out) return fail(HYPREC_E_INVALID, "predict_dense: null output");
        int64_t blk = std::max<int64_t>(hyprec::kSweepTile, ((int64_t)256 << 20) / (8 * n) / hyprec::kSweepTile * hyprec::kSweepTile);
        blk = std::min(blk, (m + hyprec::kSweepTile - 1) / hyprec::kSweepTile * hyprec::kSweepTile);
        HY_TRY(dense_blk.reserve(sizeof(double) * (size_t)blk * n));
        for (int64_t lo = 0; lo < m; lo += blk) {
            const int64_t hi = std::min(m, lo + blk);
            hyprec::SweepArgs<T> a;
            a.user_vecs = U.as<T>();
            a.item_vecs = V.as<T>();
            a.k = k;
            a.user_lo = lo;
            a.user_hi = hi;
            a.n_items = n;
            a.dense_out = dense_blk.as<double>();
            a.dense_out_f32 = nullptr;
            a.row_list = nullptr;
            a.thresholds = nullptr;
            a.counts = nullptr;
            dim3 grid((unsigned)((n + hyprec::kSweepTile - 1) / hyprec::kSweepTile),
                      (unsigned)((hi - lo + hyprec::kSweepTile - 1) / hyprec::kSweepTile));
            {
                Timer t(h, HYPREC_T_SWEEP);
                hyprec::score_sweep_kernel<T, hyprec::kSweepStore><<<grid, hyprec::kSweepThreads, 0, h->stream>>>(a);
                h->launches += 1;
            }
            CU_TRY(cudaGetLastError());
            CU_TRY(cudaMemcpyAsync(out + lo * n, dense_blk.ptr, sizeof(double) * (size_t)(hi - lo) * n,
                                   cudaMemcpyDeviceToHost, h->stream));
            CU_TRY(cudaStreamSynchronize(h->stream));
        }
        return HYPREC_OK;
    }

    // ---------------------------------------------------------------- tensor-core score sweep
    DevBuf u_hi, u_lo, v_hi, v_lo, unorm, vnorm, border_cells, border_count, dense_f32, vmax_buf;
    static constexpr unsigned int kBorderCapacity = 8u << 20;   // cells the exact re-check can take
    unsigned int last_border_cells = 0;

    typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

    int make_map(CUtensorMap* map, float* base, int64_t rows) {
        const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
        const cuuint64_t strides[1] = {(cuuint64_t)k * sizeof(float)};
        const cuuint32_t box[2] = {(cuuint32_t)hyprec::tc::kBK, (cuuint32_t)hyprec::tc::kBM};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = ((EncodeTiledFn)h->encode_tiled)(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr,
                                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(HYPREC_E_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
        return HYPREC_OK;
    }

    bool tc_usable() const { return h->tc_enabled && k % 4 == 0 && k >= 8; }
    // |s~ - s| <= coef * ||u|| * ||v||: split residual + float32 accumulation over 3 * ceil(k / 8)
    // tensor-core adds, doubled (observed maximum on B200: 1.2e-6 at k = 200)
    float tc_margin_coef() const { return (float)((3.0 * ((k + 7) / 8) + 32.0) * 2.0 * 1.1920929e-7); }

    // Operands of the tensor-core sweeps over users [lo, hi): hi / lo TF32 splits of U[lo:hi] and V, row norms,
    // TMA descriptors.
    CUtensorMap tc_mu_hi, tc_mu_lo, tc_mv_hi, tc_mv_lo;
    int tc_prepare(int64_t lo, int64_t hi) {
        namespace tcn = hyprec::tc;
        const int64_t nu = hi - lo;
        HY_TRY(u_hi.reserve(sizeof(float) * (size_t)m * k));
        HY_TRY(u_lo.reserve(sizeof(float) * (size_t)m * k));
        HY_TRY(v_hi.reserve(sizeof(float) * (size_t)n * k));
        HY_TRY(v_lo.reserve(sizeof(float) * (size_t)n * k));
        HY_TRY(unorm.reserve(sizeof(float) * (size_t)m));
        HY_TRY(vnorm.reserve(sizeof(float) * (size_t)n));
        HY_TRY(border_cells.reserve(sizeof(uint2) * (size_t)kBorderCapacity));
        HY_TRY(border_count.reserve(sizeof(unsigned int)));
        const int64_t cu = nu * k, cv = n * k;
        tcn::split_tf32_kernel<T><<<(unsigned)((cu + 255) / 256), 256, 0, h->stream>>>(U.as<T>() + lo * k, cu, u_hi.as<float>(), u_lo.as<float>());
        tcn::split_tf32_kernel<T><<<(unsigned)((cv + 255) / 256), 256, 0, h->stream>>>(V.as<T>(), cv, v_hi.as<float>(), v_lo.as<float>());
        tcn::row_norm_kernel<T><<<(unsigned)((nu + 7) / 8), 256, 0, h->stream>>>(U.as<T>() + lo * k, nu, k, unorm.as<float>());
        tcn::row_norm_kernel<T><<<(unsigned)((n + 7) / 8), 256, 0, h->stream>>>(V.as<T>(), n, k, vnorm.as<float>());
        h->launches += 4;
        CU_TRY(cudaGetLastError());
        HY_TRY(make_map(&tc_mu_hi, u_hi.as<float>(), nu));
        HY_TRY(make_map(&tc_mu_lo, u_lo.as<float>(), nu));
        HY_TRY(make_map(&tc_mv_hi, v_hi.as<float>(), n));
        HY_TRY(make_map(&tc_mv_lo, v_lo.as<float>(), n));
        return HYPREC_OK;
    }

    // The sweep itself over the prepared operands: per-row counts of score >= threshold (exact, through the margin +
    // re-check scheme), the dense float32 score matrix (diagnostic API only), and / or the survivors of the fused
    // top-k (cells of candidate items whose score reaches the user's cut).
    struct EmitArgs {
        const float* cut = nullptr;
        const unsigned int* bits = nullptr;
        int64_t bits_ld = 0;
    };
    int tc_run_sweep(int64_t lo, int64_t hi, bool want_counts, bool want_dense, int64_t* dense_ld_out, const EmitArgs* emit) {
        namespace tcn = hyprec::tc;
        const int64_t nu = hi - lo;
        const int64_t tiles_n = (n + tcn::kBN - 1) / tcn::kBN;
        const int64_t dense_ld = tiles_n * tcn::kBN;
        if (want_dense) HY_TRY(dense_f32.reserve(sizeof(float) * (size_t)nu * dense_ld));
        if (dense_ld_out) *dense_ld_out = dense_ld;
        CU_TRY(cudaMemsetAsync(border_count.ptr, 0, sizeof(unsigned int), h->stream));
        tcn::ScoreArgs a;
        a.m = nu;
        a.n = n;
        a.k = k;
        a.thresholds = thresholds.as<double>() + lo;
        a.unorm = unorm.as<float>();
        a.vnorm = vnorm.as<float>();
        a.margin_coef = tc_margin_coef();
        a.counts = want_counts ? counts.as<unsigned long long>() : nullptr;
        a.border_count = border_count.as<unsigned int>();
        a.border_cells = border_cells.as<uint2>();
        a.border_capacity = kBorderCapacity;
        a.dense = want_dense ? dense_f32.as<float>() : nullptr;
        a.dense_ld = dense_ld;
        a.row_map = nullptr;
        if (emit) {
            a.surv_cut = emit->cut;
            a.cand_bits = emit->bits;
            a.bits_ld = emit->bits_ld;
            a.surv_cnt = surv_cnt.as<unsigned int>();
            a.surv = surv.as<uint2>();
            a.surv_cap = hyprec::kSurvCap;
        }
        auto kern = tcn::tc_score_kernel;
        HY_TRY(ensure_dynamic_smem((const void*)kern, (int)tcn::kSmemBytes));
        const int64_t n_tiles = ((nu + tcn::kBM - 1) / tcn::kBM) * tiles_n;
        const int grid = (int)std::min<int64_t>(n_tiles, h->num_sms);
        kern<<<grid, tcn::kThreads, tcn::kSmemBytes, h->stream>>>(tc_mu_hi, tc_mu_lo, tc_mv_hi, tc_mv_lo, a);
        h->launches += 1;
        CU_TRY(cudaGetLastError());
        if (want_counts) {
            tcn::recheck_cells_kernel<T><<<h->num_sms * 4, 256, 0, h->stream>>>(
                border_cells.as<uint2>(), border_count.as<unsigned int>(), kBorderCapacity, U.as<T>() + lo * k, V.as<T>(), k,
                thresholds.as<double>() + lo, counts.as<unsigned long long>());
            h->launches += 1;
            CU_TRY(cudaGetLastError());
        }
        return HYPREC_OK;
    }

    int tc_sweep(int64_t lo, int64_t hi, bool want_counts, bool want_dense, int64_t* dense_ld_out) {
        HY_TRY(tc_prepare(lo, hi));
        return tc_run_sweep(lo, hi, want_counts, want_dense, dense_ld_out, nullptr);
    }

    // ---- fused top-k: sampled columns -> per-user cut (topk.cuh) -------------------------------------------------
    DevBuf surv, surv_cnt, cut_buf, cut_emit_buf, sample_f, sample_hi, sample_lo, sample_scores, cand_bits, cand_distinct, topk_redo;
    bool cand_bits_ok = false;      // the bitmap matches the candidate lists on the device
    bool fused_topk_enabled = getenv("HYPREC_FUSED_TOPK") == nullptr || getenv("HYPREC_FUSED_TOPK")[0] != '0';   // =0: exact scoring of every candidate
    bool last_topk_fused = false;

    int build_cand_bitmap(int64_t nu) {
        const int64_t words = (n + 31) / 32;
        if (cand_bits_ok) return HYPREC_OK;
        HY_TRY(cand_bits.reserve(sizeof(unsigned int) * (size_t)nu * words));
        CU_TRY(cudaMemsetAsync(cand_bits.ptr, 0, sizeof(unsigned int) * (size_t)nu * words, h->stream));
        const int grid = (int)std::min<int64_t>(nu, (int64_t)16 * h->num_sms);
        HY_TRY(cand_distinct.reserve(sizeof(int32_t) * (size_t)nu));
        hyprec::cand_bitmap_kernel<<<grid, 256, 0, h->stream>>>(cand_ptr.as<int64_t>(), cand_ids.as<int32_t>(), nu, words,
                                                                 cand_bits.as<unsigned int>(), cand_distinct.as<int32_t>());
        h->launches += 1;
        CU_TRY(cudaGetLastError());
        cand_bits_ok = true;
        return HYPREC_OK;
    }

    // cut / cut - 2E per user of [lo, hi) from a tensor-core pass over ~32 n / target evenly spread item columns
    int sample_cuts(int64_t lo, int64_t hi, bool have_list, int cap, int target) {
        namespace tcn = hyprec::tc;
        const int64_t nu = hi - lo;
        const int64_t n_sample = hyprec::eval_sample_columns(n, target, 256);
        const int64_t tiles_j = (n_sample + tcn::kBN - 1) / tcn::kBN;
        const int64_t ld = tiles_j * tcn::kBN;
        const int64_t cnt = n_sample * k;
        HY_TRY(sample_f.reserve(sizeof(float) * (size_t)cnt));
        HY_TRY(sample_hi.reserve(sizeof(float) * (size_t)cnt));
        HY_TRY(sample_lo.reserve(sizeof(float) * (size_t)cnt));
        HY_TRY(sample_scores.reserve(sizeof(float) * (size_t)nu * ld));
        HY_TRY(cut_buf.reserve(sizeof(float) * (size_t)nu));
        HY_TRY(cut_emit_buf.reserve(sizeof(float) * (size_t)nu));
        HY_TRY(vmax_buf.reserve(2 * sizeof(float)));
        hyprec::gather_rows_cast_kernel<T><<<(unsigned)((cnt + 255) / 256), 256, 0, h->stream>>>(V.as<T>(), n_sample, n, k, sample_f.as<float>());
        tcn::split_tf32_kernel<float><<<(unsigned)((cnt + 255) / 256), 256, 0, h->stream>>>(sample_f.as<float>(), cnt, sample_hi.as<float>(), sample_lo.as<float>());
        h->launches += 2;
        CUtensorMap ms_hi, ms_lo;
        HY_TRY(make_map(&ms_hi, sample_hi.as<float>(), n_sample));
        HY_TRY(make_map(&ms_lo, sample_lo.as<float>(), n_sample));
        tcn::ScoreArgs a;
        a.m = nu;
        a.n = n_sample;
        a.k = k;
        a.thresholds = nullptr;
        a.unorm = nullptr;
        a.vnorm = nullptr;
        a.margin_coef = 0.f;
        a.counts = nullptr;
        a.border_count = nullptr;
        a.border_cells = nullptr;
        a.border_capacity = 0;
        a.dense = sample_scores.as<float>();
        a.dense_ld = ld;
        a.row_map = nullptr;
        auto kern = tcn::tc_score_kernel;
        HY_TRY(ensure_dynamic_smem((const void*)kern, (int)tcn::kSmemBytes));
        const int64_t n_tiles = ((nu + tcn::kBM - 1) / tcn::kBM) * tiles_j;
        const int grid = (int)std::min<int64_t>(n_tiles, h->num_sms);
        kern<<<grid, tcn::kThreads, tcn::kSmemBytes, h->stream>>>(tc_mu_hi, tc_mu_lo, ms_hi, ms_lo, a);
        h->launches += 1;
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaMemsetAsync(vmax_buf.ptr, 0, 2 * sizeof(float), h->stream));
        hyprec::max_float_kernel<<<(unsigned)std::min<int64_t>((n + 2047) / 2048, h->num_sms), 256, 0, h->stream>>>(
            vnorm.as<float>(), n, vmax_buf.as<float>());
        hyprec::SampleCutArgs c;
        c.scores = sample_scores.as<float>();
        c.scores_ld = ld;
        c.n_sample = n_sample;
        c.n_items = n;
        c.cand_bits = have_list ? cand_bits.as<unsigned int>() : nullptr;
        c.bits_ld = (n + 31) / 32;
        c.cand_indptr = have_list ? cand_ptr.as<int64_t>() : nullptr;
        c.n_users = nu;
        c.capacity = cap;
        c.target = target;
        c.unorm = unorm.as<float>();
        c.vmax = vmax_buf.as<float>();
        c.err_coef = tc_margin_coef();
        c.cut = cut_buf.as<float>();
        c.cut_emit = cut_emit_buf.as<float>();
        // room for the sampled candidates of a user (all sampled columns when there is no list; with one, the list's
        // share of the catalogue plus a quarter), at least 2 CTAs per SM when that is little
        {
            const double share = have_list ? std::min(1.0, 1.25 * (double)cand_max / (double)std::max<int64_t>(n, 1) + 0.01) : 1.0;
            const int64_t want = std::max<int64_t>(2048, (int64_t)(share * (double)n_sample) + 64);
            const int64_t most = ((int64_t)h->max_smem - 2048) / 4;
            c.stage_cap = (int)std::min<int64_t>(std::min<int64_t>(want, n_sample), most);
        }
        const size_t cut_smem = sizeof(unsigned) * (size_t)c.stage_cap;
        HY_TRY(ensure_dynamic_smem((const void*)hyprec::sample_cut_kernel, (int)cut_smem));
        int cut_per_sm = 0;
        CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cut_per_sm, hyprec::sample_cut_kernel, hyprec::kTopkThreads, cut_smem));
        const int cgrid = (int)std::min<int64_t>(nu, (int64_t)std::max(cut_per_sm, 1) * h->num_sms);
        hyprec::sample_cut_kernel<<<cgrid, hyprec::kTopkThreads, cut_smem, h->stream>>>(c);
        h->launches += 2;
        CU_TRY(cudaGetLastError());
        HY_TRY(surv.reserve(sizeof(uint2) * (size_t)nu * hyprec::kSurvCap));
        HY_TRY(surv_cnt.reserve(sizeof(unsigned int) * (size_t)nu));
        CU_TRY(cudaMemsetAsync(surv_cnt.ptr, 0, sizeof(unsigned int) * (size_t)nu, h->stream));
        return HYPREC_OK;
    }

    int eval_stats(double* out8) override {
        for (int i = 0; i < 8; ++i) out8[i] = 0.0;
        out8[0] = last_topk_fused ? 1.0 : 0.0;
        out8[2] = (double)last_border_cells;
        if (last_topk_fused && vmax_buf.ptr) {
            unsigned int fb = 0;
            CU_TRY(cudaMemcpyAsync(&fb, vmax_buf.as<float>() + 1, sizeof(fb), cudaMemcpyDeviceToHost, h->stream));
            CU_TRY(cudaStreamSynchronize(h->stream));
            out8[1] = (double)fb;
        }
        out8[3] = (double)(surv.bytes + surv_cnt.bytes + sample_scores.bytes + cand_bits.bytes);
        out8[4] = (double)dense_f32.bytes;
        return HYPREC_OK;
    }

    int solve_phase_cycles(long long* out16, int reset) override {
        HY_TRY(prof_buf.reserve(16 * sizeof(long long)));
        CU_TRY(cudaStreamSynchronize(h->stream));
        if (out16) CU_TRY(cudaMemcpy(out16, prof_buf.ptr, 16 * sizeof(long long), cudaMemcpyDeviceToHost));
        if (reset) CU_TRY(cudaMemset(prof_buf.ptr, 0, 16 * sizeof(long long)));
        return HYPREC_OK;
    }

    int score_dense_f32(float* out) override {
        if (k == 0) return fail(HYPREC_E_INVALID, "score_dense_f32 before set_factors");
        if (!tc_usable()) return fail(HYPREC_E_UNSUPPORTED, "tensor-core score sweep unavailable (HYPREC_TC=0, not sm_100, or n_factors % 4 != 0)");
        int64_t ld = 0;
        {
            Timer t(h, HYPREC_T_SWEEP);
            HY_TRY(tc_sweep(0, m, false, true, &ld));
        }
        CU_TRY(cudaMemcpy2DAsync(out, sizeof(float) * n, dense_f32.ptr, sizeof(float) * ld, sizeof(float) * n, m,
                                 cudaMemcpyDeviceToHost, h->stream));
        CU_TRY(cudaStreamSynchronize(h->stream));
        return HYPREC_OK;
    }

    // ---------------------------------------------------------------- evaluation
    int64_t cand_lo = -1, cand_hi = -1, cand_max = 0;
    bool use_cached_cands = false;
    bool cands_are_hashed = false;      // the cached lists were generated on the device (below)
    int hashed_fold = 0, hashed_folds = 0;
    unsigned long long hashed_seed = 0;
    DevBuf hashed_counts;

    // Device-generated candidate lists (topk.cuh, hashed_candidates_kernel): from now on a NULL-list
    // eval_topk over [lo, hi) ranks, for every user, its test positives plus the 1 / n_folds share of its
    // unrated items that the hash assigns to `fold`.  n_folds <= 0 switches back to "all items".
    int set_hashed_candidates(int fold, int n_folds, unsigned long long seed) override {
        if (n_folds > 0 && (fold < 0 || fold >= n_folds)) return fail(HYPREC_E_INVALID, "hashed candidates: fold outside [0, n_folds)");
        if (cands_are_hashed) { use_cached_cands = false; cand_lo = cand_hi = -1; cands_are_hashed = false; }
        hashed_fold = fold;
        hashed_folds = n_folds > 0 ? n_folds : 0;
        hashed_seed = seed;
        return HYPREC_OK;
    }

    int generate_hashed_candidates(int64_t lo, int64_t hi) {
        if (full.rows != m || test.rows != m)
            return fail(HYPREC_E_INVALID, "hashed candidates need the ratings and test matrices (set_eval_csr)");
        const int64_t nu = hi - lo;
        if (nu <= 0) return HYPREC_OK;
        hyprec::HashedCandArgs a;
        a.full_indptr = full.indptr.template as<int64_t>();
        a.full_indices = full.indices.template as<int32_t>();
        a.test_indptr = test.indptr.template as<int64_t>();
        a.test_indices = test.indices.template as<int32_t>();
        a.user_lo = lo;
        a.user_hi = hi;
        a.n_items = n;
        a.fold = hashed_fold;
        a.n_folds = hashed_folds;
        a.seed = hashed_seed;
        HY_TRY(hashed_counts.reserve(sizeof(int32_t) * (size_t)nu));
        a.counts = hashed_counts.as<int32_t>();
        a.out_indptr = nullptr;
        a.out_ids = nullptr;
        const int grid = (int)std::min<int64_t>(nu, (int64_t)8 * h->num_sms);
        hyprec::hashed_candidates_kernel<false><<<grid, hyprec::kTopkThreads, 0, h->stream>>>(a);
        h->launches += 1;
        CU_TRY(cudaGetLastError());
        std::vector<int32_t> counts_h((size_t)nu);
        CU_TRY(cudaMemcpyAsync(counts_h.data(), hashed_counts.ptr, sizeof(int32_t) * (size_t)nu, cudaMemcpyDeviceToHost, h->stream));
        CU_TRY(cudaStreamSynchronize(h->stream));
        std::vector<int64_t> ptr_h((size_t)nu + 1, 0);
        int64_t max_cand = 0;
        for (int64_t u = 0; u < nu; ++u) {
            ptr_h[u + 1] = ptr_h[u] + counts_h[u];
            max_cand = std::max<int64_t>(max_cand, counts_h[u]);
        }
        const int64_t total = ptr_h[nu];
        HY_TRY(cand_ptr.reserve(sizeof(int64_t) * (size_t)(nu + 1)));
        HY_TRY(cand_ids.reserve(sizeof(int32_t) * (size_t)std::max<int64_t>(total, 1)));
        CU_TRY(cudaMemcpyAsync(cand_ptr.ptr, ptr_h.data(), sizeof(int64_t) * (size_t)(nu + 1), cudaMemcpyHostToDevice, h->stream));
        a.out_indptr = cand_ptr.as<int64_t>();
        a.out_ids = cand_ids.as<int32_t>();
        hyprec::hashed_candidates_kernel<true><<<grid, hyprec::kTopkThreads, 0, h->stream>>>(a);
        h->launches += 1;
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaStreamSynchronize(h->stream));             // ptr_h is read by the copy above
        cand_lo = lo;
        cand_hi = hi;
        cand_max = std::max<int64_t>(max_cand, 1);
        cand_bits_ok = false;
        use_cached_cands = true;
        cands_are_hashed = true;
        return HYPREC_OK;
    }

    // Candidate streams of users [lo, hi): validated (one min/max pass), copied to the device and
    // kept until replaced, so that a fold's lists cross PCIe once.
    int upload_candidates(int64_t lo, int64_t hi, const int64_t* cp, const int32_t* ci) {
        const int64_t nu = hi - lo;
        if (!cp || !ci) return fail(HYPREC_E_INVALID, "candidates: null array");
        if (cp[0] != 0) return fail(HYPREC_E_INVALID, "candidates: cand_indptr[0] != 0");
        int64_t max_cand = 0;
        for (int64_t u = 0; u < nu; ++u) {
            if (cp[u + 1] < cp[u]) return fail(HYPREC_E_INVALID, "candidates: cand_indptr not monotone");
            max_cand = std::max(max_cand, cp[u + 1] - cp[u]);
        }
        const int64_t total = cp[nu];
        int32_t lo_id = 0, hi_id = 0;
        for (int64_t e = 0; e < total; ++e) {   // branch-free: vectorises
            lo_id = std::min(lo_id, ci[e]);
            hi_id = std::max(hi_id, ci[e]);
        }
        if (lo_id < 0 || hi_id >= n) return fail(HYPREC_E_INVALID, "candidates: item id out of range");
        HY_TRY(cand_ptr.reserve(sizeof(int64_t) * (nu + 1)));
        HY_TRY(cand_ids.reserve(sizeof(int32_t) * std::max<int64_t>(total, 1)));
        CU_TRY(cudaMemcpyAsync(cand_ptr.ptr, cp, sizeof(int64_t) * (nu + 1), cudaMemcpyHostToDevice, h->stream));
        if (total > 0)
            CU_TRY(cudaMemcpyAsync(cand_ids.ptr, ci, sizeof(int32_t) * total, cudaMemcpyHostToDevice, h->stream));
        cand_lo = lo;
        cand_hi = hi;
        cand_max = std::max<int64_t>(max_cand, 1);
        cand_bits_ok = false;
        cands_are_hashed = false;
        return HYPREC_OK;
    }

    int set_candidates(int64_t lo, int64_t hi, const int64_t* cp, const int32_t* ci) override {
        if (m == 0) return fail(HYPREC_E_INVALID, "set_candidates before set_train_csr");
        if (!cp && !ci) { use_cached_cands = false; cands_are_hashed = false; cand_lo = cand_hi = -1; return HYPREC_OK; }
        if (lo < 0 || hi > m || lo > hi) return fail(HYPREC_E_INVALID, "set_candidates: user range outside the matrix");
        HY_TRY(upload_candidates(lo, hi, cp, ci));
        CU_TRY(cudaStreamSynchronize(h->stream));
        use_cached_cands = true;
        return HYPREC_OK;
    }

    int eval_topk(int64_t lo, int64_t hi, const int64_t* cp, const int32_t* ci, int cap, double* thr_out,
                  int64_t* ones_out, int32_t* tidx, double* tval, double* thit, uint8_t* thit8, uint8_t* trf,
                  uint8_t* tef) override {
        if (k == 0) return fail(HYPREC_E_INVALID, "eval_topk before set_factors");
        if (lo < 0 || hi > m || lo > hi) return fail(HYPREC_E_INVALID, "eval_topk: user range outside the matrix");
        const int64_t nu = hi - lo;
        const bool want_topk = tidx || tval || thit || thit8;
        if (thit8 && !hits_u8_exact)
            return fail(HYPREC_E_UNSUPPORTED, "eval_topk: one-byte hit values need ratings that are integers in 0..255 "
                                              "(hit = rating x rounded prediction, evaluator.py:287-288); ask for the float64 hits");
        if (want_topk && (cap <= 0 || cap > 4096)) return fail(HYPREC_E_UNSUPPORTED, "capacity must be in [1, 4096]");
        if (ci && !cp) return fail(HYPREC_E_INVALID, "eval_topk: candidate ids without indptr");
        if (cp && !ci) return fail(HYPREC_E_INVALID, "eval_topk: candidate indptr without ids");
        if (tef && test.rows != m) return fail(HYPREC_E_INVALID, "eval_topk: test flags need set_eval_csr first");

        // thresholds for every user (cheap), float64
        HY_TRY(compute_colsum_v());
        HY_TRY(thresholds.reserve(sizeof(double) * m));
        {
            const int wpb = 8;
            hyprec::row_threshold_kernel<T><<<(unsigned)((m + wpb - 1) / wpb), wpb * 32, 0, h->stream>>>(
                U.as<T>(), colsum.as<double>(), k, m, n, thresholds.as<double>());
            h->launches += 1;
            CU_TRY(cudaGetLastError());
        }
        if (nu == 0) return HYPREC_OK;
        if (thr_out)
            CU_TRY(cudaMemcpyAsync(thr_out, thresholds.as<double>() + lo, sizeof(double) * nu, cudaMemcpyDeviceToHost, h->stream));

        // ---- candidate streams first: the sweep's epilogue needs to know whose candidate an item is
        if (want_topk && !ci && !cp && hashed_folds > 0 &&
            !(use_cached_cands && cands_are_hashed && cand_lo == lo && cand_hi == hi))
            HY_TRY(generate_hashed_candidates(lo, hi));
        int64_t max_cand = n;
        bool have_list = false;
        if (want_topk) {
            const bool cached = !ci && cp == nullptr && cand_lo == lo && cand_hi == hi && cand_max > 0 && use_cached_cands;
            if (ci) {
                HY_TRY(upload_candidates(lo, hi, cp, ci));
                max_cand = cand_max;
            } else if (cached) {
                max_cand = cand_max;
            }
            have_list = ci != nullptr || cached;
            max_cand = std::max<int64_t>(max_cand, 1);
        }

        // ---- fused score -> mask -> top-k: the tensor-core sweep (needed for the rounded counts anyway) emits only
        // the candidate cells that can reach a user's top; no users x items score matrix exists (topk.cuh).
        const int target = hyprec::eval_cut_target(cap);
        const int64_t bitmap_bytes = have_list ? (int64_t)sizeof(unsigned int) * nu * ((n + 31) / 32) : 0;
        const bool fused = fused_topk_enabled && tc_usable() && want_topk && n >= 4096 && cap >= 64 &&
                           target <= (hyprec::kSurvCap * 7) / 10 && max_cand > cap && bitmap_bytes <= ((int64_t)4 << 30);
        bool counts_done = false;
        if (ones_out) {
            HY_TRY(counts.reserve(sizeof(unsigned long long) * nu));
            CU_TRY(cudaMemsetAsync(counts.ptr, 0, sizeof(unsigned long long) * nu, h->stream));
        }
        if (tc_usable() && (ones_out || fused)) {
            EmitArgs emit;
            {
                Timer t(h, HYPREC_T_COUNT);
                HY_TRY(tc_prepare(lo, hi));
            }
            if (fused) {
                Timer t(h, HYPREC_T_TOPK);
                if (have_list) HY_TRY(build_cand_bitmap(nu));
                HY_TRY(sample_cuts(lo, hi, have_list, cap, target));
                emit.cut = cut_emit_buf.as<float>();
                emit.bits = have_list ? cand_bits.as<unsigned int>() : nullptr;
                emit.bits_ld = (n + 31) / 32;
            }
            {
                Timer t(h, HYPREC_T_COUNT);
                HY_TRY(tc_run_sweep(lo, hi, ones_out != nullptr, false, nullptr, fused ? &emit : nullptr));
            }
            if (ones_out) {
                unsigned int border = 0;
                CU_TRY(cudaMemcpyAsync(&border, border_count.ptr, sizeof(border), cudaMemcpyDeviceToHost, h->stream));
                CU_TRY(cudaStreamSynchronize(h->stream));
                last_border_cells = border;
                counts_done = border <= kBorderCapacity;   // else: list overflowed, redo on the SIMT kernel
                if (!counts_done) CU_TRY(cudaMemsetAsync(counts.ptr, 0, sizeof(unsigned long long) * nu, h->stream));
            }
        }
        if (ones_out) {
            if (!counts_done) {
                hyprec::SweepArgs<T> a;
                a.user_vecs = U.as<T>();
                a.item_vecs = V.as<T>();
                a.k = k;
                a.user_lo = lo;
                a.user_hi = hi;
                a.n_items = n;
                a.dense_out = nullptr;
                a.dense_out_f32 = nullptr;
                a.row_list = nullptr;
                a.thresholds = thresholds.as<double>();
                a.counts = counts.as<unsigned long long>();
                dim3 grid((unsigned)((n + hyprec::kSweepTile - 1) / hyprec::kSweepTile),
                          (unsigned)((nu + hyprec::kSweepTile - 1) / hyprec::kSweepTile));
                Timer t(h, HYPREC_T_COUNT);
                hyprec::score_sweep_kernel<T, hyprec::kSweepCount><<<grid, hyprec::kSweepThreads, 0, h->stream>>>(a);
                h->launches += 1;
            }
            CU_TRY(cudaGetLastError());
            static_assert(sizeof(unsigned long long) == sizeof(int64_t), "count width");
            CU_TRY(cudaMemcpyAsync(ones_out, counts.ptr, sizeof(int64_t) * nu, cudaMemcpyDeviceToHost, h->stream));
        }

        if (want_topk) {
            int sort_size = 2;
            while (sort_size < cap) sort_size <<= 1;
            hyprec::TopkLayout lay = hyprec::topk_layout((int)max_cand, k, sort_size, fused, fused);
            // Lists beyond one SM's shared memory (full catalogue / naive-split lists of a citeulike-sized item
            // set): the per-candidate arrays go to a per-CTA global scratch, everything else stays in shared memory.
            const bool spill = lay.total > (size_t)h->max_smem;
            size_t spill_stride = 0;
            if (spill) {
                spill_stride = (lay.off_u + 255) & ~(size_t)255;         // keys, positions, float keys
                lay = hyprec::topk_layout(0, k, sort_size, fused, fused);
                if (lay.total > (size_t)h->max_smem)
                    return fail(HYPREC_E_UNSUPPORTED, "eval_topk: n_factors / capacity too large for one CTA's shared memory");
            }
            auto kern = spill ? hyprec::eval_topk_kernel<T, true> : hyprec::eval_topk_kernel<T, false>;
            HY_TRY(ensure_dynamic_smem((const void*)kern, (int)lay.total));
            int per_sm = 0;
            CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, hyprec::kTopkThreads, lay.total));
            per_sm = std::max(per_sm, 1);
            int grid = (int)std::min<int64_t>(nu, (int64_t)per_sm * h->num_sms);
            if (spill) {
                const size_t budget = (size_t)4 << 30;                   // scratch of all resident CTAs
                grid = (int)std::max<size_t>(1, std::min<size_t>((size_t)grid, budget / spill_stride));
                HY_TRY(topk_spill.reserve(spill_stride * (size_t)grid));
            }
            HY_TRY(out_idx.reserve(sizeof(int32_t) * (size_t)nu * cap));
            HY_TRY(out_val.reserve(sizeof(double) * (size_t)nu * cap));
            HY_TRY(out_hit.reserve(sizeof(double) * (size_t)nu * cap));
            hyprec::TopkArgs<T> a;
            a.user_vecs = U.as<T>();
            a.item_vecs = V.as<T>();
            a.k = k;
            a.n_items = n;
            a.user_lo = lo;
            a.user_hi = hi;
            a.cand_indptr = have_list ? cand_ptr.as<int64_t>() : nullptr;
            a.cand_ids = have_list ? cand_ids.as<int32_t>() : nullptr;
            a.thresholds = thresholds.as<double>();
            const bool have_full = full.rows == m;
            a.full_indptr = have_full ? full.indptr.template as<int64_t>() : nullptr;
            a.full_indices = have_full ? full.indices.template as<int32_t>() : nullptr;
            a.full_values = have_full ? full.values.template as<double>() : nullptr;
            a.capacity = cap;
            a.sort_size = sort_size;
            a.max_cand = (int)max_cand;
            a.topk_idx = out_idx.as<int32_t>();
            a.topk_val = out_val.as<double>();
            a.topk_hit = out_hit.as<double>();
            a.approx = nullptr;
            a.approx_ld = 0;
            a.unorm = nullptr;
            a.err_coef = 0.f;
            a.vmax = nullptr;
            a.fallbacks = nullptr;
            a.spill = spill ? topk_spill.as<unsigned char>() : nullptr;
            a.spill_stride = spill_stride;
            if (fused) {
                a.surv = surv.as<uint2>();
                a.surv_cnt = surv_cnt.as<unsigned int>();
                a.surv_cap = hyprec::kSurvCap;
                a.cut = cut_buf.as<float>();
                a.cand_distinct = have_list ? cand_distinct.as<int32_t>() : nullptr;
                a.unorm = unorm.as<float>();
                a.err_coef = tc_margin_coef();
                a.vmax = vmax_buf.as<float>();
                a.fallbacks = reinterpret_cast<unsigned int*>(vmax_buf.as<float>() + 1);
            }
            if (fused) {
                // pass 1: the emitted cells alone, small per-candidate arrays, several CTAs per SM; users that need
                // the candidate stream (repeats in the stream, ties in the order, a misjudged cut) are flagged ...
                HY_TRY(topk_redo.reserve((size_t)nu));
                CU_TRY(cudaMemsetAsync(topk_redo.ptr, 0, (size_t)nu, h->stream));
                hyprec::TopkArgs<T> f = a;
                f.pass = 1;
                f.redo = topk_redo.as<unsigned char>();
                f.max_cand = hyprec::kSurvCap;
                f.spill = nullptr;
                f.spill_stride = 0;
                const hyprec::TopkLayout flay = hyprec::topk_layout(hyprec::kSurvCap, k, sort_size, true, true);
                auto fkern = hyprec::eval_topk_kernel<T, false>;
                HY_TRY(ensure_dynamic_smem((const void*)fkern, (int)std::max(flay.total, spill ? (size_t)0 : lay.total)));
                int fper_sm = 0;
                CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fper_sm, fkern, hyprec::kTopkThreads, flay.total));
                fper_sm = std::max(fper_sm, 1);
                const int fgrid = (int)std::min<int64_t>(nu, (int64_t)fper_sm * h->num_sms);
                Timer t(h, HYPREC_T_TOPK);
                fkern<<<fgrid, hyprec::kTopkThreads, flay.total, h->stream>>>(f);
                // ... and pass 2 does those through the stream, with arrays for the whole stream
                a.pass = 2;
                a.redo = topk_redo.as<unsigned char>();
                kern<<<grid, hyprec::kTopkThreads, lay.total, h->stream>>>(a);
                h->launches += 2;
            } else {
                Timer t(h, HYPREC_T_TOPK);
                kern<<<grid, hyprec::kTopkThreads, lay.total, h->stream>>>(a);
                h->launches += 1;
            }
            CU_TRY(cudaGetLastError());
            last_topk_fused = fused;
            if (tidx) CU_TRY(cudaMemcpyAsync(tidx, out_idx.ptr, sizeof(int32_t) * (size_t)nu * cap, cudaMemcpyDeviceToHost, h->stream));
            if (tval) CU_TRY(cudaMemcpyAsync(tval, out_val.ptr, sizeof(double) * (size_t)nu * cap, cudaMemcpyDeviceToHost, h->stream));
            if (thit) CU_TRY(cudaMemcpyAsync(thit, out_hit.ptr, sizeof(double) * (size_t)nu * cap, cudaMemcpyDeviceToHost, h->stream));
            if (thit8) {
                // the same hit values, one byte each: 1/8 of the device -> host bytes of the float64 form
                const size_t count = (size_t)nu * cap;
                HY_TRY(out_hit8.reserve(count));
                hyprec::cast_kernel<double, uint8_t><<<(unsigned)((count + 255) / 256), 256, 0, h->stream>>>(
                    out_hit.as<double>(), (int64_t)count, out_hit8.as<uint8_t>());
                h->launches += 1;
                CU_TRY(cudaGetLastError());
                CU_TRY(cudaMemcpyAsync(thit8, out_hit8.ptr, count, cudaMemcpyDeviceToHost, h->stream));
            }
        }

        // rounded predictions at train / test non-zeros
        const Csr* which[2] = {&csr, &test};
        uint8_t* outs[2] = {trf, tef};
        for (int s = 0; s < 2; ++s) {
            if (!outs[s]) continue;
            const Csr& c = *which[s];
            if (c.rows != m) return fail(HYPREC_E_INVALID, "eval_topk: test matrix missing (set_eval_csr)");
            const int64_t e0 = c.h_indptr[lo], e1 = c.h_indptr[hi];
            if (e1 == e0) continue;
            HY_TRY(flags.reserve((size_t)(e1 - e0)));
            const int wpb = 8;
            {
                Timer t(h, HYPREC_T_FLAGS);
                hyprec::nz_flags_kernel<T><<<(unsigned)((e1 - e0 + wpb - 1) / wpb), wpb * 32, 0, h->stream>>>(
                    c.indptr.template as<int64_t>(), c.indices.template as<int32_t>(), U.as<T>(), V.as<T>(), k,
                    thresholds.as<double>(), lo, hi, flags.as<uint8_t>());
                h->launches += 1;
            }
            CU_TRY(cudaGetLastError());
            CU_TRY(cudaMemcpyAsync(outs[s], flags.ptr, (size_t)(e1 - e0), cudaMemcpyDeviceToHost, h->stream));
            CU_TRY(cudaStreamSynchronize(h->stream));   // flags buffer is reused for the next matrix
        }
        CU_TRY(cudaStreamSynchronize(h->stream));
        return HYPREC_OK;
    }
};

// Peak-rate probe: 16 independent FMA chains per thread, enough CTAs to fill every SM.
template <typename T>
__global__ void fma_probe_kernel(T* out, int iters, T seed) {
    T acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = seed + T(i) + T(threadIdx.x);
    const T a = T(1.0000001), b = T(1e-9);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = acc[i] * a + b;
    }
    T s = T(0);
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == T(-1)) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true; keeps the loop alive
}

template <typename T>
int run_fma_probe(hyprec_ctx* h, double* tflops) {
    const int iters = 4096, threads = 256, blocks = h->num_sms * 16;
    T* out = nullptr;
    CU_TRY(cudaMalloc(&out, sizeof(T) * (size_t)threads * blocks));
    cudaEvent_t b, e;
    CU_TRY(cudaEventCreate(&b));
    CU_TRY(cudaEventCreate(&e));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CU_TRY(cudaEventRecord(b, h->stream));
        fma_probe_kernel<T><<<blocks, threads, 0, h->stream>>>(out, iters, T(rep));
        CU_TRY(cudaEventRecord(e, h->stream));
        CU_TRY(cudaEventSynchronize(e));
        float ms = 0.f;
        CU_TRY(cudaEventElapsedTime(&ms, b, e));
        const double flops = 2.0 * 16.0 * iters * (double)threads * blocks;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
        h->launches += 1;
    }
    cudaEventDestroy(b);
    cudaEventDestroy(e);
    cudaFree(out);
    *tflops = best;
    return HYPREC_OK;
}

int check(hyprec_t* h) {
    if (!h || !h->engine) return fail(HYPREC_E_INVALID, "null handle");
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) return fail(HYPREC_E_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    return HYPREC_OK;
}

}  // namespace

extern "C" {

int hyprec_abi_version(void) { return 2; }

const char* hyprec_last_error(void) { return g_error.c_str(); }

int hyprec_create(hyprec_t** out, int device_id, int precision) {
    if (!out) return fail(HYPREC_E_INVALID, "hyprec_create: null output");
    *out = nullptr;
    if (precision != HYPREC_F64 && precision != HYPREC_F32) return fail(HYPREC_E_INVALID, "precision must be 0 (f64) or 1 (f32)");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(HYPREC_E_CUDA, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                                       " (hyprec_b200 has no CPU fallback)");
    if (device_id < 0 || device_id >= count) return fail(HYPREC_E_INVALID, "device id out of range");
    CU_TRY(cudaSetDevice(device_id));
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device_id));
    hyprec_ctx* h = new hyprec_ctx();
    h->device = device_id;
    h->precision = precision;
    h->num_sms = prop.multiProcessorCount;
    h->max_smem = (int)prop.sharedMemPerBlockOptin;
    CU_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (int i = 0; i < hyprec_ctx::kAux; ++i) {
        CU_TRY(cudaStreamCreateWithFlags(&h->aux[i], cudaStreamNonBlocking));
        CU_TRY(cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming));
    }
    CU_TRY(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&h->ev_base, cudaEventDisableTiming));
    const char* tcv = getenv("HYPREC_TC");
    h->tc_enabled = !(tcv && tcv[0] == '0') && prop.major == 10;
    if (h->tc_enabled) {
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &h->encode_tiled, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess || !h->encode_tiled) {
            h->tc_enabled = false;
            cudaGetLastError();
        }
    }
    const char* pf = getenv("HYPREC_SOLVE_PROF");
    h->prof_enabled = pf && pf[0] == '1';
    const char* ts_env = getenv("HYPREC_TC_SOLVE");
    h->tc_solve_enabled = h->tc_enabled && !(ts_env && ts_env[0] == '0');
    const char* tch = getenv("HYPREC_TC_CHOL");
    h->tc_chol_enabled = h->tc_solve_enabled && !(tch && tch[0] == '0');
    const char* rf = getenv("HYPREC_RMSE_FUSED");
    h->rmse_fused_enabled = !(rf && rf[0] == '0');
    const char* tmu = getenv("HYPREC_TC_MULTI");
    h->tc_multi_enabled = !(tmu && tmu[0] == '0');
    const char* ps = getenv("HYPREC_SOLVE_PROF_SLOT");
    if (ps) h->prof_slot = atoi(ps);
    const char* pu = getenv("HYPREC_PUSH");
    h->fused_push = pu && pu[0] == 'f';
    const char* sm64 = getenv("HYPREC_SMALL_F64");
    h->small_f64_enabled = !(sm64 && sm64[0] == '0');
    const char* spl = getenv("HYPREC_SPLIT");
    h->split_enabled = !(spl && spl[0] == '0');
    const char* lr = getenv("HYPREC_LOWRANK");
    h->lowrank_enabled = !(lr && lr[0] == '0');
    auto six = [](const char* name, int* out, int lo, int hi, int multiple) {
        const char* v = getenv(name);
        if (!v) return;
        int got[6], n = 0;
        while (n < 6 && *v) {
            char* end = nullptr;
            const long x = strtol(v, &end, 10);
            if (end == v || x < lo || x > hi || x % multiple != 0) return;      // malformed: keep the defaults
            got[n++] = (int)x;
            v = (*end == ',') ? end + 1 : end;
            if (*end != ',' && *end != 0) return;
        }
        if (n == 6) for (int i = 0; i < 6; ++i) out[i] = got[i];
    };
    if (precision == HYPREC_F64) {
        // float64: systems of up to 127 unknowns on 4 x 4 tiles (no register spills, twice the threads per system):
        // 1.90 -> 1.74 ms per citeulike-a epoch, 3.45 -> 3.15 ms at the citeulike-t shape (tools/ab_epoch.py,
        // profiles/r02_ab_*.jsonl).  float32 keeps 8 x 8 tiles there: its <= 127 bucket runs in tensor memory.
        const int threads64[6] = {64, 96, 256, 256, 256, 256}, tile64[6] = {4, 4, 4, 8, 8, 8};
        for (int i = 0; i < 6; ++i) { h->bucket_threads[i] = threads64[i]; h->bucket_tile[i] = tile64[i]; }
    }
    six("HYPREC_BUCKET_THREADS", h->bucket_threads, 32, 1024, 32);
    six("HYPREC_BUCKET_TILE", h->bucket_tile, 4, 8, 4);
    const char* lc = getenv("HYPREC_LOWRANK_CAP");
    if (lc) h->lowrank_cap = std::max(0, atoi(lc));
    if (precision == HYPREC_F64) h->engine = new Engine<double>(h);
    else h->engine = new Engine<float>(h);
    *out = h;
    return HYPREC_OK;
}

int hyprec_destroy(hyprec_t* h) {
    if (!h) return HYPREC_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    delete h->engine;
    h->harvest();
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    for (int i = 0; i < hyprec_ctx::kAux; ++i) {
        if (h->aux[i]) cudaStreamDestroy(h->aux[i]);
        if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_base) cudaEventDestroy(h->ev_base);
    if (h->flush_buf) cudaFree(h->flush_buf);
    for (int i = 0; i < 8; ++i)
        if (h->marks[i]) cudaEventDestroy(h->marks[i]);
    cudaStreamDestroy(h->stream);
    delete h;
    return HYPREC_OK;
}

int64_t hyprec_launch_count(const hyprec_t* h) { return h ? h->launches : 0; }

int hyprec_set_train_csr(hyprec_t* h, int64_t m, int64_t n, const int64_t* indptr, const int32_t* indices,
                         const double* values) {
    HY_TRY(check(h));
    return h->engine->set_train(m, n, indptr, indices, values);
}

int hyprec_set_shard(hyprec_t* h, int64_t user_lo, int64_t user_hi, int64_t item_lo, int64_t item_hi) {
    HY_TRY(check(h));
    h->user_lo = user_lo; h->user_hi = user_hi; h->item_lo = item_lo; h->item_hi = item_hi;
    return HYPREC_OK;
}

int hyprec_set_factors(hyprec_t* h, const double* u, const double* v, int k) {
    HY_TRY(check(h));
    return h->engine->set_factors(u, v, k);
}

int hyprec_get_factors(hyprec_t* h, double* u, double* v) {
    HY_TRY(check(h));
    return h->engine->get_factors(u, v);
}

int hyprec_device_factors(hyprec_t* h, void** u, void** v, int* elem) {
    HY_TRY(check(h));
    return h->engine->device_factors(u, v, elem);
}

int hyprec_set_item_prior(hyprec_t* h, const double* theta) {
    HY_TRY(check(h));
    return h->engine->set_prior(theta);
}

int hyprec_als_half_step(hyprec_t* h, int side, double lambda) {
    HY_TRY(check(h));
    return h->engine->half_step(side, lambda);
}

int hyprec_rmse(hyprec_t* h, double* out) {
    HY_TRY(check(h));
    if (!out) return fail(HYPREC_E_INVALID, "hyprec_rmse: null output");
    return h->engine->rmse(out);
}

int hyprec_rmse_terms(hyprec_t* h, double* out3) {
    HY_TRY(check(h));
    if (!out3) return fail(HYPREC_E_INVALID, "hyprec_rmse_terms: null output");
    return h->engine->rmse_terms(out3, true);
}

int hyprec_train(hyprec_t* h, int n_iter, double lambda, double start_error, hyprec_epoch_cb cb, void* user,
                 int* epochs_run, double* last_rmse) {
    HY_TRY(check(h));
    if ((h->user_hi >= 0 || h->item_hi >= 0) && !h->engine->comm_active())
        return fail(HYPREC_E_INVALID, "hyprec_train with a shard set but no exchange attached (hyprec_comm_attach): loop "
                                      "hyprec_als_half_step and exchange the shards yourself");
    double error = start_error;
    int done = 0;
    for (int epoch = 1; epoch <= n_iter; ++epoch) {
        const double old_error = error;
        const auto t0 = std::chrono::steady_clock::now();
        HY_TRY(h->engine->half_step(HYPREC_SIDE_USER, lambda));
        HY_TRY(h->engine->half_step(HYPREC_SIDE_ITEM, lambda));
        const auto t1 = std::chrono::steady_clock::now();
        HY_TRY(h->engine->rmse(&error));
        done = epoch;
        if (cb) cb(epoch, error, std::chrono::duration<double>(t1 - t0).count(), user);
        if (error >= old_error) break;   // collaborative_filtering.py:256
    }
    if (epochs_run) *epochs_run = done;
    if (last_rmse) *last_rmse = error;
    return HYPREC_OK;
}

int hyprec_predict_dense(hyprec_t* h, double* out) {
    HY_TRY(check(h));
    return h->engine->predict_dense(out);
}

int hyprec_set_eval_csr(hyprec_t* h, const int64_t* fp, const int32_t* fi, const double* fv, const int64_t* tp,
                        const int32_t* ti, const double* tv) {
    HY_TRY(check(h));
    return h->engine->set_eval(fp, fi, fv, tp, ti, tv);
}

int hyprec_solve_phase_cycles(hyprec_t* h, long long* out16, int reset) {
    HY_TRY(check(h));
    return h->engine->solve_phase_cycles(out16, reset);
}

int hyprec_score_dense_f32(hyprec_t* h, float* out) {
    HY_TRY(check(h));
    if (!out) return fail(HYPREC_E_INVALID, "hyprec_score_dense_f32: null output");
    return h->engine->score_dense_f32(out);
}

int hyprec_set_candidates(hyprec_t* h, int64_t user_lo, int64_t user_hi, const int64_t* cand_indptr,
                          const int32_t* cand_ids) {
    HY_TRY(check(h));
    return h->engine->set_candidates(user_lo, user_hi, cand_indptr, cand_ids);
}

int hyprec_set_hashed_candidates(hyprec_t* h, int fold, int n_folds, uint64_t seed) {
    HY_TRY(check(h));
    return h->engine->set_hashed_candidates(fold, n_folds, (unsigned long long)seed);
}

int hyprec_eval_topk(hyprec_t* h, int64_t user_lo, int64_t user_hi, const int64_t* cand_indptr,
                     const int32_t* cand_ids, int capacity, double* thresholds, int64_t* ones_per_row,
                     int32_t* topk_idx, double* topk_val, double* topk_hit, uint8_t* train_flags,
                     uint8_t* test_flags) {
    HY_TRY(check(h));
    return h->engine->eval_topk(user_lo, user_hi, cand_indptr, cand_ids, capacity, thresholds, ones_per_row,
                                topk_idx, topk_val, topk_hit, nullptr, train_flags, test_flags);
}

int hyprec_eval_topk_u8(hyprec_t* h, int64_t user_lo, int64_t user_hi, const int64_t* cand_indptr,
                        const int32_t* cand_ids, int capacity, double* thresholds, int64_t* ones_per_row,
                        int32_t* topk_idx, double* topk_val, uint8_t* topk_hit, uint8_t* train_flags,
                        uint8_t* test_flags) {
    HY_TRY(check(h));
    return h->engine->eval_topk(user_lo, user_hi, cand_indptr, cand_ids, capacity, thresholds, ones_per_row,
                                topk_idx, topk_val, nullptr, topk_hit, train_flags, test_flags);
}

int hyprec_kernel_ms(hyprec_t* h, int which, double* total_ms, int64_t* calls, double* last_ms) {
    HY_TRY(check(h));
    if (which < 0 || which >= kNumTimers) return fail(HYPREC_E_INVALID, "hyprec_kernel_ms: bad timer id");
    CU_TRY(cudaStreamSynchronize(h->stream));
    h->harvest();
    if (total_ms) *total_ms = h->ms_total[which];
    if (calls) *calls = h->ms_calls[which];
    if (last_ms) *last_ms = h->ms_last[which];
    return HYPREC_OK;
}

int hyprec_reset_kernel_ms(hyprec_t* h) {
    HY_TRY(check(h));
    CU_TRY(cudaStreamSynchronize(h->stream));
    h->harvest();
    for (int i = 0; i < kNumTimers; ++i) {
        h->ms_total[i] = 0;
        h->ms_last[i] = 0;
        h->ms_calls[i] = 0;
    }
    return HYPREC_OK;
}

int hyprec_probe_fma_tflops(hyprec_t* h, int precision, double* tflops) {
    HY_TRY(check(h));
    if (!tflops) return fail(HYPREC_E_INVALID, "hyprec_probe_fma_tflops: null output");
    if (precision == HYPREC_F64) return run_fma_probe<double>(h, tflops);
    if (precision == HYPREC_F32) return run_fma_probe<float>(h, tflops);
    return fail(HYPREC_E_INVALID, "precision must be 0 (f64) or 1 (f32)");
}

int hyprec_host_alloc(hyprec_t* h, size_t bytes, void** out) {
    HY_TRY(check(h));
    if (!out) return fail(HYPREC_E_INVALID, "hyprec_host_alloc: null output");
    CU_TRY(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return HYPREC_OK;
}

int hyprec_host_free(hyprec_t* h, void* ptr) {
    HY_TRY(check(h));
    if (ptr) CU_TRY(cudaFreeHost(ptr));
    return HYPREC_OK;
}

int hyprec_mark(hyprec_t* h, int slot) {
    HY_TRY(check(h));
    if (slot < 0 || slot >= 8) return fail(HYPREC_E_INVALID, "hyprec_mark: slot must be in [0, 8)");
    if (!h->marks[slot]) CU_TRY(cudaEventCreate(&h->marks[slot]));
    CU_TRY(cudaEventRecord(h->marks[slot], h->stream));
    return HYPREC_OK;
}

int hyprec_elapsed_ms(hyprec_t* h, int from, int to, double* ms) {
    HY_TRY(check(h));
    if (from < 0 || from >= 8 || to < 0 || to >= 8 || !h->marks[from] || !h->marks[to] || !ms)
        return fail(HYPREC_E_INVALID, "hyprec_elapsed_ms: both slots must have been marked");
    CU_TRY(cudaEventSynchronize(h->marks[to]));
    float f = 0.f;
    CU_TRY(cudaEventElapsedTime(&f, h->marks[from], h->marks[to]));
    *ms = (double)f;
    return HYPREC_OK;
}

int hyprec_eval_stats(hyprec_t* h, double* out8) {
    HY_TRY(check(h));
    if (!out8) return fail(HYPREC_E_INVALID, "hyprec_eval_stats: null output");
    return h->engine->eval_stats(out8);
}

int hyprec_comm_blob_bytes(void) { return (int)sizeof(CommBlob); }

int hyprec_comm_export(hyprec_t* h, int rank, int world, void* blob, size_t blob_bytes) {
    HY_TRY(check(h));
    if (!blob) return fail(HYPREC_E_INVALID, "hyprec_comm_export: null blob");
    return h->engine->comm_export(rank, world, blob, blob_bytes);
}

int hyprec_comm_attach(hyprec_t* h, const void* blobs, size_t blob_bytes, const int64_t* user_bounds, const int64_t* item_bounds) {
    HY_TRY(check(h));
    return h->engine->comm_attach(blobs, blob_bytes, user_bounds, item_bounds);
}

int hyprec_comm_arena(hyprec_t* h, void** base, size_t* bytes) {
    HY_TRY(check(h));
    return h->engine->comm_arena(base, bytes);
}

int hyprec_comm_attach_local(hyprec_t* h, void* const* arenas, const int* devices, const int64_t* user_bounds,
                             const int64_t* item_bounds) {
    HY_TRY(check(h));
    return h->engine->comm_attach_ptrs(arenas, devices, user_bounds, item_bounds);
}

int hyprec_comm_detach(hyprec_t* h) {
    HY_TRY(check(h));
    return h->engine->comm_detach();
}

int hyprec_comm_stats(hyprec_t* h, double* out8) {
    HY_TRY(check(h));
    if (!out8) return fail(HYPREC_E_INVALID, "hyprec_comm_stats: null output");
    return h->engine->comm_stats(out8);
}

int hyprec_init_factors_uniform(hyprec_t* h, int k, uint64_t seed) {
    HY_TRY(check(h));
    return h->engine->init_factors_uniform(k, (unsigned long long)seed);
}

int hyprec_get_factor_rows(hyprec_t* h, int side, const int64_t* rows, int64_t n_rows, double* out) {
    HY_TRY(check(h));
    return h->engine->get_factor_rows(side, rows, n_rows, out);
}

int hyprec_flush_l2(hyprec_t* h) {
    HY_TRY(check(h));
    const size_t want = (size_t)384 << 20;   // 3x the 126 MB L2
    if (!h->flush_buf) {
        CU_TRY(cudaMalloc(&h->flush_buf, want));
        h->flush_bytes = want;
    }
    CU_TRY(cudaMemsetAsync(h->flush_buf, 0x5a, h->flush_bytes, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    return HYPREC_OK;
}

#include "host_api.cuh"   // hyprec_host_*: host-only entry points (no device involved)

}  // extern "C"
