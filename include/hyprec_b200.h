/*
 * hyprec_b200.h -- C ABI of the B200-native HyPRec hot path.
 *
 * The reference (mostafa-mahmoud/HyPRec) is pure Python and has no FFI of its own; its only
 * replaceable seam is the Python class lib/collaborative_filtering.py:13 CollaborativeFiltering
 * (SURVEY.md section 8b).  hyprec_b200/collaborative_filtering.py keeps that class surface and
 * binds the functions below with ctypes; each entry point names the reference arithmetic it
 * replaces.  Plain pointers and sizes only: no torch / numpy types cross this boundary.
 *
 * Conventions
 *   - every function returns 0 on success or a negative HYPREC_E_* code; the message is
 *     available from hyprec_last_error() (thread-local, valid until the next failing call);
 *   - one handle owns one GPU (one process per GPU; multi-GPU runs create one handle per rank);
 *     a handle is not thread-safe;
 *   - host buffers are caller-owned and may be pageable or pinned; device memory is owned by
 *     the library until hyprec_destroy();
 *   - users x items matrices are row-major; factor matrices are rows x k row-major float64 on
 *     the host regardless of the device precision;
 *   - confidence weights are fixed at 1.0 on train non-zeros and 0.1 elsewhere, as in
 *     lib/collaborative_filtering.py:120-144 (build_confidence_matrix).
 */
#ifndef HYPREC_B200_H
#define HYPREC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HYPREC_OK 0
#define HYPREC_E_INVALID (-1)   /* bad argument / call order                                   */
#define HYPREC_E_CUDA (-2)      /* CUDA runtime error (message has the CUDA string)             */
#define HYPREC_E_NOMEM (-3)     /* host or device allocation failed                             */
#define HYPREC_E_UNSUPPORTED (-4) /* shape outside what the kernels handle (message says which) */
#define HYPREC_E_SINGULAR (-5)  /* non-finite solution: the reference raises LinAlgError here    */

#define HYPREC_F64 0            /* float64 storage and arithmetic (parity with numpy float64)   */
#define HYPREC_F32 1            /* float32 factor storage / solve, float64 reductions           */

#define HYPREC_SIDE_USER 0
#define HYPREC_SIDE_ITEM 1

typedef struct hyprec_ctx hyprec_t;

/* ---- lifetime -------------------------------------------------------------------------- */
int hyprec_create(hyprec_t** out, int device_id, int precision);
int hyprec_destroy(hyprec_t* h);
const char* hyprec_last_error(void);
int hyprec_abi_version(void);
/* number of kernels this handle has launched since creation (bench.py "gpu_launches") */
int64_t hyprec_launch_count(const hyprec_t* h);

/* ---- data ------------------------------------------------------------------------------ */
/* Train matrix as CSR (users x items).  Replaces the dense self.train_data the reference
 * scans in build_confidence_matrix (collaborative_filtering.py:120-144) and the ratings
 * argument of als_step (:69).  indices must be sorted within a row; values NULL => all 1.0.
 * The item-major (CSC) copy is built here. */
int hyprec_set_train_csr(hyprec_t* h, int64_t n_users, int64_t n_items, const int64_t* indptr,
                         const int32_t* indices, const double* values);

/* Rows this handle solves in a half-step (multi-GPU sharding).  Default:
 * all rows.  Rows outside the range are left untouched in the local replica; the caller
 * exchanges shards (all-gather over the pointers hyprec_device_factors returns). */
int hyprec_set_shard(hyprec_t* h, int64_t user_lo, int64_t user_hi, int64_t item_lo, int64_t item_hi);

/* self.user_vecs / self.item_vecs (collaborative_filtering.py:173-178); k = n_factors. */
int hyprec_set_factors(hyprec_t* h, const double* user_vecs, const double* item_vecs, int k);
int hyprec_get_factors(hyprec_t* h, double* user_vecs, double* item_vecs);
/* Device addresses and element size (8 or 4) of the replicated factor matrices, for the
 * caller's collective (torch.distributed / NCCL all-gather of the solved shard).  Every call tells
 * the library that the matrices may have changed (cached Gram matrices are dropped).  With a shard
 * set, the contract is the exchange of the side just solved: only rows OUTSIDE this handle's shard
 * may be written through the pointers. */
int hyprec_device_factors(hyprec_t* h, void** user_vecs_dev, void** item_vecs_dev, int* elem_size);

/* The same initial state without a host copy, for problem sizes whose factors should stay on the device
 * (BASELINE configs[3]: 3 GB of float32 factors): U and V drawn uniformly in [0, 1) from a counter-based
 * generator -- element e = (float)(mix64(seed' * 0xD1342543DE82EF95 + e) >> 40) / 2^24 with seed' = seed for U
 * and seed + 1 for V; oracle.als.uniform_factors restates it.  Stands for numpy.random.random((m, k)),
 * numpy.random.random((n, k)) at collaborative_filtering.py:173-176. */
int hyprec_init_factors_uniform(hyprec_t* h, int k, uint64_t seed);
/* Rows `rows[0..n_rows)` of U (side 0) or V (side 1) as float64, n_rows x k: parity checks on a subsample of a
 * problem whose full factors are never copied to the host. */
int hyprec_get_factor_rows(hyprec_t* h, int side, const int64_t* rows, int64_t n_rows, double* out);

/* self.document_distribution for the update_with_items rule (:93-96); NULL clears it. */
int hyprec_set_item_prior(hyprec_t* h, const double* theta);

/* ---- ALS ------------------------------------------------------------------------------- */
/* One half-step == als_step(latent, fixed, train, _lambda, type) (collaborative_filtering.py
 * :69-99) for the handle's shard: Gram of the fixed side, per-row CSR gather of
 * sum y y^T and sum r y, k x k Cholesky solve, row written back. */
int hyprec_als_half_step(hyprec_t* h, int side, double lambda);

/* evaluator.get_rmse(U.dot(V.T), train) (evaluator.py:236-252, called at
 * collaborative_filtering.py:247) without forming U.V^T; float64 accumulation in every mode. */
int hyprec_rmse(hyprec_t* h, double* out);

/* The three sums the RMSE is made of, for sharded runs: out3 = { <U^T U, V^T V>_F,
 * sum_nz r (u.v), sum_nz r^2 } where the two sparse sums cover only the handle's user shard.
 * rmse = sqrt((out3[0] - 2 * SUM_ranks out3[1] + SUM_ranks out3[2]) / (n_users * n_items)). */
int hyprec_rmse_terms(hyprec_t* h, double* out3);

/* partial_train() (collaborative_filtering.py:224-259): epochs of user half-step, item
 * half-step, RMSE; stops after the epoch whose error >= the previous one (:256).
 * start_error is the initial `error` (inf, or the epoch-0 RMSE when verbose, :228-234).
 * The callback (may be NULL) sees every epoch: (epoch starting at 1, rmse, seconds of the two
 * half-steps).  Single-GPU only: with a shard set the caller loops half-steps itself. */
typedef void (*hyprec_epoch_cb)(int epoch, double rmse, double seconds, void* user);
int hyprec_train(hyprec_t* h, int n_iter, double lambda, double start_error, hyprec_epoch_cb cb,
                 void* user, int* epochs_run, double* last_rmse);

/* ---- multi-GPU exchange over NVLink peer memory ----------------------------------------- */
/* One process per GPU, one handle per process (the reference has no counterpart: what
 * shards is the independence of rows inside als_step, collaborative_filtering.py:82-86).  After
 * hyprec_set_train_csr + hyprec_set_factors (same data on every rank):
 *   1. every rank calls hyprec_comm_export: its replicas of U, V and of the whitened factors plus a mailbox move
 *      into one arena; `blob` (hyprec_comm_blob_bytes() bytes) receives the arena's CUDA IPC handle;
 *   2. the caller all-gathers the blobs (any transport: torch.distributed, MPI, files);
 *   3. every rank calls hyprec_comm_attach with all blobs in rank order and the row partition
 *      (user_bounds / item_bounds: world + 1 ascending offsets from 0 to n_users / n_items; rank r owns
 *      [bounds[r], bounds[r+1]) of each side).
 * From then on hyprec_als_half_step solves the rank's rows and stores them into EVERY rank's replica from the
 * epilogue of the kernel that produced them, reduces the k x k Gram through the mailbox, whitens its own rows into
 * every replica and returns without synchronising the host; hyprec_rmse / hyprec_rmse_terms return the GLOBAL value
 * (bit-identical on every rank) and hyprec_train runs the sharded epoch loop.  Every rank must issue the same
 * sequence of half-step / rmse / train calls.  A peer that does not arrive within 20 s fails the next
 * hyprec_rmse / hyprec_train with HYPREC_E_CUDA instead of hanging the GPU.
 * hyprec_comm_detach (collective: put a process barrier before anything reuses the memory) undoes it. */
int hyprec_comm_blob_bytes(void);
int hyprec_comm_export(hyprec_t* h, int rank, int world, void* blob, size_t blob_bytes);
int hyprec_comm_attach(hyprec_t* h, const void* blobs, size_t blob_bytes, const int64_t* user_bounds,
                       const int64_t* item_bounds);
int hyprec_comm_detach(hyprec_t* h);
/* The same exchange between handles of ONE process (one host process driving several GPUs, or several handles on
 * one GPU): after hyprec_comm_export on every handle, hyprec_comm_arena returns a handle's arena and
 * hyprec_comm_attach_local takes all arenas (rank order) with the device ordinal each lives on -- plain device
 * pointers, peer access enabled where the devices differ, no IPC.  The handles must then be driven from one host
 * thread each (a half-step waits on the device for its peers' half-steps). */
int hyprec_comm_arena(hyprec_t* h, void** base, size_t* bytes);
int hyprec_comm_attach_local(hyprec_t* h, void* const* arenas, const int* devices, const int64_t* user_bounds,
                             const int64_t* item_bounds);
/* out8 = { attached (0/1), bytes this rank stored into peer replicas since attach, barriers passed, arena bytes, 0... } */
int hyprec_comm_stats(hyprec_t* h, double* out8);

/* ---- prediction / evaluation ------------------------------------------------------------ */
/* get_predictions(): U.dot(V.T) (collaborative_filtering.py:270), n_users x n_items float64
 * written to host memory.  Meant for the API's dense return value at small shapes. */
int hyprec_predict_dense(hyprec_t* h, double* out);

/* Approximate predictions from the tensor-core sweep (tcgen05, 3xTF32 split, float32 accumulate),
 * n_users x n_items float32 on the host: |out - u.v| <= ~3e-5 * ||u|| * ||v||.
 * The evaluation runs the same sweep without storing it (cells far from a threshold are decided in the epilogue,
 * top-k candidates above a per-user cut are emitted as a short list); this dense copy exists for tests and
 * diagnostics only.  HYPREC_E_UNSUPPORTED when unavailable. */
int hyprec_score_dense_f32(hyprec_t* h, float* out);

/* The matrices the report needs besides train: full ratings (evaluator.ratings) and test.
 * Same CSR conventions as hyprec_set_train_csr. */
int hyprec_set_eval_csr(hyprec_t* h, const int64_t* full_indptr, const int32_t* full_indices,
                        const double* full_values, const int64_t* test_indptr,
                        const int32_t* test_indices, const double* test_values);

/* Candidate streams of users [user_lo, user_hi) -- evaluator.test_indices[user * (1 + fold)] in
 * stream order (evaluator.py:225) -- copied to the device once per fold and used by every later
 * hyprec_eval_topk call over the same user range that passes cand_indptr = cand_ids = NULL.
 * Passing NULL, NULL here drops them (eval_topk with NULL lists then means "all items"). */
int hyprec_set_candidates(hyprec_t* h, int64_t user_lo, int64_t user_hi, const int64_t* cand_indptr,
                          const int32_t* cand_ids);

/* Candidate lists generated on the device, for problem sizes whose lists cannot be enumerated by the host
 * (BASELINE configs[3]: 4e5 candidates x 1e6 users).  The reference's fold list of a user is its fold positives
 * plus a 1/k_folds share of the items it never rated (evaluator.py:176-192, drawn from per-user shuffles);
 * here the share is decided by a 64-bit mix of (user, item, seed):
 *     candidate(u, i)  <=>  i in test_u  or  (i not in ratings_u  and  mix(u, i, seed) % n_folds == fold)
 * in ascending item order.  After this call every hyprec_eval_topk with cand_indptr = cand_ids = NULL ranks these
 * lists (generated per user range, kept until the range, the parameters or the eval matrices change).  Needs
 * hyprec_set_eval_csr.  n_folds <= 0 switches the behaviour off again.  No counterpart call in the reference. */
int hyprec_set_hashed_candidates(hyprec_t* h, int fold, int n_folds, uint64_t seed);

/* Everything get_evaluation_report() (abstract_recommender.py:100-131) derives from scores,
 * for users [user_lo, user_hi) (0, n_users for a single GPU); output arrays are indexed
 * relative to user_lo:
 *   thresholds[u]        row average of the prediction row (abstract_recommender.py:183)
 *   ones_per_row[u]      #items with score >= threshold  (rounded_predictions, :184-186)
 *   topk_idx/val/hit     top-`capacity` of the user's candidate stream in the order
 *                        load_top_recommendations leaves it (evaluator.py:215-234 +
 *                        util/top_recommendations.py:23-40), -1 padded; hit = full rating of
 *                        that cell times its rounded prediction (evaluator.py:287-288)
 *   train_flags/test_flags  rounded prediction at every train / test non-zero of these users,
 *                        in CSR order (calculate_recall, evaluator.py:254-266)
 * cand_indptr is relative to user_lo (cand_indptr[0] == 0); cand_ids NULL => the lists installed by
 * hyprec_set_candidates for this user range, or, without those, every item in index order
 * (recommend_items, abstract_recommender.py:200-203).  Any output may be NULL. */
int hyprec_eval_topk(hyprec_t* h, int64_t user_lo, int64_t user_hi, const int64_t* cand_indptr,
                     const int32_t* cand_ids, int capacity, double* thresholds,
                     int64_t* ones_per_row, int32_t* topk_idx, double* topk_val, double* topk_hit,
                     uint8_t* train_flags, uint8_t* test_flags);
/* The same call with the hit values as ONE byte per list entry instead of a float64: for a top-200 report at the
 * citeulike-a shape the results shrink from 58.5 MB (ids + float64 scores + float64 hits) to 5.6 MB when topk_val is
 * NULL as well, and the device -> host copy is most of the end-to-end evaluation time.  Exact when every rating is an
 * integer in 0..255 (the reference's data is 0 / 1, util/data_parser.py:111-115); HYPREC_E_UNSUPPORTED otherwise. */
int hyprec_eval_topk_u8(hyprec_t* h, int64_t user_lo, int64_t user_hi, const int64_t* cand_indptr,
                        const int32_t* cand_ids, int capacity, double* thresholds,
                        int64_t* ones_per_row, int32_t* topk_idx, double* topk_val, uint8_t* topk_hit,
                        uint8_t* train_flags, uint8_t* test_flags);

/* Diagnostics of the most recent hyprec_eval_topk: out8 = { top-k fed by sweep survivors (1) or by exact scoring of
 * every candidate (0), users of that call that fell back to exact scoring, cells the rounded-count sweep re-checked
 * in float64, bytes of the survivor / sample / bitmap buffers, bytes of the dense float32 score buffer (only
 * hyprec_score_dense_f32 allocates it), 0... }. */
int hyprec_eval_stats(hyprec_t* h, double* out8);

/* Device time of the library's kernels, measured with CUDA events on the launching stream and
 * accumulated per class since creation / the last reset: total ms, number of timed regions, and
 * the ms of the regions of the most recent API call.  For bench.py's roofline. */
#define HYPREC_T_GRAM 0
#define HYPREC_T_SOLVE 1
#define HYPREC_T_RMSE 2
#define HYPREC_T_SWEEP 3
#define HYPREC_T_TOPK 4
#define HYPREC_T_FLAGS 5
#define HYPREC_T_COUNT 6
#define HYPREC_T_PREP 7   /* low-rank path: factor of 0.1 G + lambda I, L^-1, whitening / un-whitening GEMMs */
#define HYPREC_T_COMM 8   /* multi-GPU: mailbox pushes, row pushes, exchange barriers (hyprec_comm_attach) */
int hyprec_kernel_ms(hyprec_t* h, int which, double* total_ms, int64_t* calls, double* last_ms);
/* Whole-region device timing for bench.py: hyprec_mark records CUDA event `slot` (0..7) on the library's stream
 * behind everything issued so far; hyprec_elapsed_ms waits for event `to` and returns the device time between the
 * two marks.  One clock for single-GPU and sharded epochs (the max over ranks is the caller's). */
int hyprec_mark(hyprec_t* h, int slot);
int hyprec_elapsed_ms(hyprec_t* h, int from, int to, double* ms);
int hyprec_reset_kernel_ms(hyprec_t* h);
/* Diagnostic (HYPREC_SOLVE_PROF=1 at hyprec_create): clock64() totals per phase of the direct row
 * solve, accumulated by block 0 -- out16[0..7] = gather+accumulate, base add, first diagonal tile,
 * panel solves, next block column, next diagonal tile, wait for the trailing update, back
 * substitution.  reset != 0 clears the counters after reading. */
int hyprec_solve_phase_cycles(hyprec_t* h, long long* out16, int reset);

/* Measured FMA rate of this GPU (TFLOP/s, best of 4 after a warm-up) in the given precision:
 * the roofline denominator for the SIMT solve kernel, which MEASURED_PEAKS.json does not hold. */
int hyprec_probe_fma_tflops(hyprec_t* h, int precision, double* tflops);
/* Page-locked host memory for callers that want full PCIe rate on the factor / top-k transfers
 * (every host pointer of this API may be pageable or pinned). */
int hyprec_host_alloc(hyprec_t* h, size_t bytes, void** out);
int hyprec_host_free(hyprec_t* h, void* ptr);
/* Overwrite a buffer 3x the size of L2 so the next kernel starts cold (benchmark hygiene). */
int hyprec_flush_l2(hyprec_t* h);

/* ---- host-side split (no device involved) ------------------------------------------------------------------
 * The k-fold candidate lists of Evaluator.get_kfold_indices (lib/evaluator.py:136-195): per user, in user order,
 * one legacy numpy.random.shuffle of its rated items and one of its non-rated items, dealt into k_folds lists of
 * int(n_items / k_folds) ids (the fold's share of the positives first).  mt_key[624] / mt_pos are numpy's global
 * MT19937 state (numpy.random.get_state()[1:3]), advanced exactly as the reference's calls advance it, so the lists
 * and everything drawn afterwards (the initial factors, collaborative_filtering.py:173-176) are bit-identical.
 * Slot u * k_folds + f occupies ids[slot_ptr[slot] .. slot_ptr[slot + 1]); slot_positives[slot] of its leading
 * entries are rated cells.  ids64 and / or ids32 receive the lists (either may be NULL); capacity = entries they
 * hold (n_users * k_folds * (n_items / k_folds) always suffices).  HYPREC_E_UNSUPPORTED, with the generator state
 * untouched, when a list would need the reference's negative slice bounds (a fold's positives outnumber
 * n_items / k_folds): the caller then runs the reference's loop.  indices must be strictly increasing within a row.
 * The stream is walked sequentially (which draws a shuffle consumes does not depend on the data shuffled); the
 * shuffles themselves run on up to 16 host threads (HYPREC_HOST_THREADS). */
/* The draws of Evaluator.naive_split_users (lib/evaluator.py:92-93) from numpy's global MT19937 state (as above):
 * per user, in user order, sizes[u] = int(test_percentage * len_u) samples WITH replacement among the user's len_u
 * ratings (numpy.random.choice -> legacy randint: masked rejection draws; none when len_u == 1).  picks receives
 * sum(sizes) CSR entry numbers (indptr[u] + position), user after user. */
int hyprec_host_naive_picks(uint32_t* mt_key, int* mt_pos, int64_t n_users, const int64_t* indptr, const int64_t* sizes,
                            int64_t* picks);
/* CSR form of a dense row-major float64 users x items block without numpy's nonzero() scan (0.42 s per 754 MB
 * matrix): hyprec_host_row_counts counts the non-zeros of every row, the caller turns the counts into indptr, and
 * hyprec_host_dense_to_csr fills indices / values of rows [0, rows) at indptr[r] (indptr is the block's own slice of
 * the matrix-wide array; indices / values are the matrix-wide arrays).  Callers run row blocks on several threads. */
int hyprec_host_row_counts(const double* x, int64_t rows, int64_t cols, int64_t* counts);
int hyprec_host_dense_to_csr(const double* x, int64_t rows, int64_t cols, const int64_t* indptr, int32_t* indices,
                             double* values);
/* Number of non-zero entries of a float64 buffer (no device involved; callers scan row blocks of a dense users x items
 * matrix on several host threads -- the check that a fold matrix still holds exactly the recorded ratings). */
int64_t hyprec_host_count_nonzero(const double* x, int64_t count);
int hyprec_host_kfold_lists(uint32_t* mt_key, int* mt_pos, int64_t n_users, int64_t n_items, int k_folds,
                            const int64_t* indptr, const int32_t* indices, int64_t capacity, int64_t* slot_ptr,
                            int64_t* slot_positives, int64_t* ids64, int32_t* ids32);

#ifdef __cplusplus
}
#endif
#endif /* HYPREC_B200_H */
