// Host-only entry points of the C ABI (include/hyprec_b200.h, "host-side split"): no CUDA call, no handle.  Included
// by api.cu inside its extern "C" block (uses its error-string helper `fail`).
// (api.cu includes host_split.cuh -- templates, no C linkage -- with its other headers, before the block.)
#pragma once

int hyprec_host_naive_picks(uint32_t* mt_key, int* mt_pos, int64_t n_users, const int64_t* indptr, const int64_t* sizes,
                            int64_t* picks) {
    if (!mt_key || !mt_pos || !indptr || !sizes || !picks) return fail(HYPREC_E_INVALID, "host_naive_picks: null argument");
    if (n_users < 0 || *mt_pos < 0 || *mt_pos > 624) return fail(HYPREC_E_INVALID, "host_naive_picks: bad sizes or generator position");
    for (int64_t u = 0; u < n_users; ++u) {
        const int64_t len = indptr[u + 1] - indptr[u];
        if (len < 0 || len >= ((int64_t)1 << 32) || sizes[u] < 0 || (len == 0 && sizes[u] != 0))
            return fail(HYPREC_E_INVALID, "host_naive_picks: bad row length or sample size");
    }
    hyprec_host::naive_picks(mt_key, mt_pos, n_users, indptr, sizes, picks);
    return HYPREC_OK;
}

int hyprec_host_dense_to_csr(const double* x, int64_t rows, int64_t cols, const int64_t* indptr, int32_t* indices,
                             double* values) {
    if (!x || !indptr || !indices || !values || rows < 0 || cols < 0 || cols >= ((int64_t)1 << 31))
        return fail(HYPREC_E_INVALID, "host_dense_to_csr: bad argument");
    for (int64_t r = 0; r < rows; ++r) {
        const double* row = x + r * cols;
        int64_t at = indptr[r];
        const int64_t end = indptr[r + 1];
        for (int64_t c = 0; c < cols; ++c)
            if (row[c] != 0.0) {
                if (at >= end) return fail(HYPREC_E_INVALID, "host_dense_to_csr: more non-zeros than indptr allows");
                indices[at] = (int32_t)c;
                values[at] = row[c];
                ++at;
            }
        if (at != end) return fail(HYPREC_E_INVALID, "host_dense_to_csr: fewer non-zeros than indptr says");
    }
    return HYPREC_OK;
}

int hyprec_host_row_counts(const double* x, int64_t rows, int64_t cols, int64_t* counts) {
    if (!x || !counts || rows < 0 || cols < 0) return fail(HYPREC_E_INVALID, "host_row_counts: bad argument");
    for (int64_t r = 0; r < rows; ++r) {
        const double* row = x + r * cols;
        int64_t c = 0;
        for (int64_t j = 0; j < cols; ++j) c += (row[j] != 0.0);
        counts[r] = c;
    }
    return HYPREC_OK;
}

int64_t hyprec_host_count_nonzero(const double* x, int64_t count) {
    int64_t c = 0;
    if (x)
        for (int64_t i = 0; i < count; ++i) c += (x[i] != 0.0);
    return c;
}

int hyprec_host_kfold_lists(uint32_t* mt_key, int* mt_pos, int64_t n_users, int64_t n_items, int k_folds,
                            const int64_t* indptr, const int32_t* indices, int64_t capacity, int64_t* slot_ptr,
                            int64_t* slot_positives, int64_t* ids64, int32_t* ids32) {
    if (!mt_key || !mt_pos || !indptr || !slot_ptr || !slot_positives || (!ids64 && !ids32))
        return fail(HYPREC_E_INVALID, "host_kfold_lists: null argument");
    if (n_users < 0 || n_items <= 0 || n_items >= ((int64_t)1 << 31) || k_folds < 1 || *mt_pos < 0 || *mt_pos > 624)
        return fail(HYPREC_E_INVALID, "host_kfold_lists: bad sizes or generator position");
    for (int64_t u = 0; u < n_users; ++u) {
        if (indptr[u + 1] < indptr[u]) return fail(HYPREC_E_INVALID, "host_kfold_lists: indptr not monotone");
        for (int64_t e = indptr[u]; e < indptr[u + 1]; ++e) {
            if (indices[e] < 0 || indices[e] >= n_items) return fail(HYPREC_E_INVALID, "host_kfold_lists: item id out of range");
            if (e > indptr[u] && indices[e] <= indices[e - 1])
                return fail(HYPREC_E_INVALID, "host_kfold_lists: item ids must be strictly increasing in a row");
        }
    }
    int n_threads = (int)std::min<unsigned>(16u, std::max<unsigned>(1u, std::thread::hardware_concurrency()));
    const char* ht = getenv("HYPREC_HOST_THREADS");
    if (ht && atoi(ht) > 0) n_threads = atoi(ht);
    const int rc = hyprec_host::kfold_lists(mt_key, mt_pos, n_users, n_items, k_folds, indptr, indices, capacity,
                                            slot_ptr, slot_positives, ids64, ids32, n_threads);
    if (rc == 1)
        return fail(HYPREC_E_UNSUPPORTED, "host_kfold_lists: a fold's share of a user's positives exceeds n_items / k_folds "
                                          "(negative slice bounds in evaluator.py:184-189); generator state untouched");
    if (rc == 2) return fail(HYPREC_E_INVALID, "host_kfold_lists: capacity too small");
    return HYPREC_OK;
}

