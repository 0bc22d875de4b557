"""
Host-side mirror of the reference ``Evaluator`` (/root/reference/lib/evaluator.py:10-341): same
constructor, attributes and method names, so the reference's callers (RecommenderSystem,
GridSearch, runnables.py) and its tests can use either class.

What differs is how the work is done, not what comes out:

* the k-fold / naive splits consume the global ``numpy.random`` stream exactly as the reference
  does (seed 42 unless ``random_seed``; one ``shuffle(rated)`` then one ``shuffle(non_rated)`` per
  user in user order, evaluator.py:144-171; one ``choice`` per user, :92-93) so the splits are
  bit-identical, but fold matrices are assembled with array indexing instead of per-user
  Python ``set`` arithmetic (evaluator.py:206-213 costs ~20 s per fold at citeulike-a size);
* CSR views of ratings / train / test are kept next to the dense matrices the API exposes;
* the metric methods compute each per-user value with the reference's own arithmetic and hand
  the list to the same ``numpy.mean(..., dtype=numpy.float16)`` call (evaluator.py:292, 315, 341),
  which keeps the reported scalars bit-identical;
* ``load_top_recommendations`` and everything that needs scores is normally driven by
  ``hyprec_b200.CollaborativeFiltering.get_evaluation_report`` from device results
  (``set_device_results``); the dense-matrix entry points remain for API compatibility;
* ``ratings`` may be a scipy sparse matrix ("dense API at scale"): the splits then
  consume the random stream exactly as above but hand out scipy CSR train / test matrices built from the
  index lists -- no dense users x items array exists anywhere -- and, for shapes whose per-user candidate
  lists cannot be enumerated (the k-fold lists hold n_items / k_folds ids per user, the naive split's lists
  n_items), ``use_hashed_candidates`` switches the evaluation to device-generated lists.
"""
import os

import numpy
import scipy.sparse as sp

from hyprec_b200.top_recommendations import TopRecommendations, top_k_stream_order


_POOL = None


def host_pool():
    """A few host threads for memory-bound scans of users x items arrays (numpy releases the GIL inside them)."""
    global _POOL
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=max(1, min(8, os.cpu_count() or 1)), thread_name_prefix="hyprec-host")
    return _POOL


def count_nonzero(matrix):
    """numpy.count_nonzero of a dense float64 matrix, row blocks counted by the library's host code on the host
    threads (numpy's own scan holds the GIL: 0.12 s per 754 MB citeulike-a fold matrix, and the class checks every
    fold matrix it is handed this way)."""
    rows = matrix.shape[0] if matrix.ndim == 2 else 0
    if matrix.size < (1 << 22) or rows < 16 or matrix.dtype != numpy.float64 or not matrix.flags.c_contiguous:
        return int(numpy.count_nonzero(matrix))
    try:
        from hyprec_b200 import _native
        _native.load_library()
    except (RuntimeError, OSError):
        return int(numpy.count_nonzero(matrix))
    step = -(-rows // 16)
    return int(sum(host_pool().map(_native.host_count_nonzero, [matrix[at:at + step] for at in range(0, rows, step)])))


def dense_copy(matrix):
    """numpy.array(matrix, dtype=float64) with the row blocks copied (and their fresh pages faulted in) on the host
    threads: 0.21 s -> 0.035 s for a citeulike-a fold matrix."""
    matrix = numpy.asarray(matrix)
    if matrix.ndim != 2 or matrix.size < (1 << 22):
        return numpy.array(matrix, dtype=numpy.float64)
    out = numpy.empty(matrix.shape, dtype=numpy.float64)
    step = -(-matrix.shape[0] // 16)
    list(host_pool().map(lambda at: numpy.copyto(out[at:at + step], matrix[at:at + step]), range(0, matrix.shape[0], step)))
    return out


def dense_to_csr(matrix):
    """scipy CSR (float64, no stored zeros, sorted column ids) of a dense users x items array.  Large float64 arrays
    are scanned by the library's host code on the host threads (numpy's nonzero(): 0.42 s per citeulike-a matrix)."""
    matrix = numpy.asarray(matrix)
    if matrix.ndim == 2 and matrix.size >= (1 << 22) and matrix.dtype == numpy.float64 and matrix.flags.c_contiguous:
        try:
            from hyprec_b200 import _native
            indptr, indices, values = _native.host_dense_to_csr(matrix, host_pool())
            return sp.csr_matrix((values, indices, indptr), shape=matrix.shape)
        except (RuntimeError, OSError, AttributeError):
            pass
    csr = sp.csr_matrix(matrix, dtype=numpy.float64)
    csr.eliminate_zeros()
    csr.sort_indices()
    return csr


class Evaluator(object):
    def __init__(self, ratings, abstracts_preprocessor=None, random_seed=False, verbose=False):
        self.sparse = sp.issparse(ratings)
        if self.sparse:
            ratings = sp.csr_matrix(ratings, dtype=numpy.float64)
            ratings.eliminate_zeros()
            ratings.sort_indices()
        self.ratings = ratings
        self.hashed_candidates = None            # (n_folds, seed) once use_hashed_candidates() was called
        self.n_users, self.n_items = ratings.shape
        if abstracts_preprocessor:
            self.abstracts_preprocessor = abstracts_preprocessor
        self.random_seed = random_seed
        self._verbose = verbose
        self.k_folds = None
        if self._verbose:
            print('%d users and %d items' % (self.n_users, self.n_items))
        self.recommendation_indices = [[] for _ in range(self.n_users)]
        self.recs_loaded = False
        self._ratings_csr = None
        # per-user hit flags (ratings * rounded along recommendation_indices) and test likes,
        # filled by set_device_results(); None => metric methods fall back to dense inputs.
        self._device_hits = None

    # ------------------------------------------------------------------ plumbing
    def get_abstracts_preprocessor(self):
        return self.abstracts_preprocessor

    def get_ratings(self):
        return self.ratings

    def use_hashed_candidates(self, n_folds=5, seed=0):
        """Rank, for every user, its test positives plus the 1 / n_folds share of its unrated items that a hash of
        (user, item, seed) selects, generated on the device (include/hyprec_b200.h,
        hyprec_set_hashed_candidates) -- for problem sizes whose candidate lists the host cannot enumerate.  No
        counterpart in the reference (its lists come from per-user shuffles, evaluator.py:162-192)."""
        self.hashed_candidates = (int(n_folds), int(seed))

    def ratings_csr(self):
        if self.sparse:
            return self.ratings
        if self._ratings_csr is None:
            self._ratings_csr = dense_to_csr(self.ratings)
        return self._ratings_csr

    def set_kfolds(self, kfolds):
        self.k_folds = kfolds
        self.test_percentage = 1.0 / self.k_folds

    # ------------------------------------------------------------------ naive split (:67-118)
    def naive_split(self, type='user'):
        if type == 'user':
            return self.naive_split_users()
        return self.naive_split_items()

    def naive_split_users(self):
        if self.random_seed is False:
            numpy.random.seed(42)
        csr = self.ratings_csr()
        picks = self._native_naive_picks(csr)
        if picks is not None:
            # every draw of the loop below, made by the library's host code; the cells follow by indexing
            rows_of = numpy.repeat(numpy.arange(self.n_users, dtype=numpy.int64), numpy.diff(csr.indptr))
            test_keys = rows_of[picks] * self.n_items + csr.indices[picks].astype(numpy.int64)
            if self.sparse:
                train, test = self._split_sparse(test_keys)
                self.test_indices = test
                return train, test
            rows, cols = rows_of[picks], csr.indices[picks]
            test = numpy.zeros(self.ratings.shape)
            train = dense_copy(self.ratings) if self.ratings.dtype == numpy.float64 else self.ratings.copy()
            train[rows, cols] = 0.
            test[rows, cols] = csr.data[picks]
            self.test_indices = test
            self._remember_csr(train, test, test_keys)
            return train, test
        keys = []
        if self.sparse:
            for user in range(self.n_users):
                non_zeros = csr.indices[csr.indptr[user]:csr.indptr[user + 1]]
                picked = numpy.random.choice(non_zeros, size=int(self.test_percentage * len(non_zeros)))   # :92-93
                keys.append(user * self.n_items + picked.astype(numpy.int64))
            train, test = self._split_sparse(numpy.concatenate(keys) if keys else numpy.zeros(0, dtype=numpy.int64))
            self.test_indices = test
            return train, test
        test = numpy.zeros(self.ratings.shape)
        train = self.ratings.copy()
        for user in range(self.n_users):
            non_zeros = csr.indices[csr.indptr[user]:csr.indptr[user + 1]]
            # sampling WITH replacement, evaluator.py:92-93
            picked = numpy.random.choice(non_zeros, size=int(self.test_percentage * len(non_zeros)))
            train[user, picked] = 0.
            test[user, picked] = self.ratings[user, picked]
            keys.append(user * self.n_items + picked.astype(numpy.int64))
        self.test_indices = test
        self._remember_csr(train, test, numpy.concatenate(keys) if keys else numpy.zeros(0, dtype=numpy.int64))
        return train, test

    def _native_naive_picks(self, csr):
        """CSR entry numbers of every cell the naive split samples (csrc/host_split.cuh: the per-user
        numpy.random.choice calls of evaluator.py:92-93 restated in C++, same stream, same draws), or None when it does
        not apply (HYPREC_NATIVE_SPLIT=0, library not built)."""
        if os.environ.get("HYPREC_NATIVE_SPLIT", "1") == "0" or self.n_users == 0:
            return None
        lengths = numpy.diff(csr.indptr).astype(numpy.int64)
        sizes = numpy.array([int(self.test_percentage * int(n)) for n in numpy.unique(lengths)], dtype=numpy.int64)
        sizes = sizes[numpy.searchsorted(numpy.unique(lengths), lengths)]        # int(p * len), evaluator.py:93
        try:
            from hyprec_b200 import _native
            return _native.host_naive_picks(csr.indptr, sizes)
        except (RuntimeError, OSError, AttributeError):
            return None

    def naive_split_items(self):
        if self.random_seed is False:
            numpy.random.seed(42)
        columns = numpy.random.choice(list(range(self.n_items)),
                                      size=int(self.test_percentage * self.n_items))
        train = self.ratings.copy()
        test = numpy.zeros(self.ratings.shape)
        train[:, columns] = 0
        test[:, columns] = self.ratings[:, columns]
        return train, test

    # ------------------------------------------------------------------ k-fold (:120-213)
    def get_kfold_indices(self):
        """Flat list of n_users * k_folds id arrays; slot ``u * k_folds + f`` holds user u's
        fold-f candidates: its share of shuffled positives followed by shuffled non-rated items,
        int(n_items / k_folds) long (evaluator.py:176-192).  Normally produced by the library's host code
        (``_native_kfold_lists``: int32 views of one buffer); the loop below is the same thing step by step."""
        if self.random_seed is False:
            numpy.random.seed(42)
        csr = self.ratings_csr()
        n, k = self.n_items, self.k_folds
        native = self._native_kfold_lists(csr)
        if native is not None:
            return native
        all_items = numpy.arange(n)
        out = []
        n_positives = []        # per slot: how many entries at the head of the list are rated cells
        for user in range(self.n_users):
            rated = csr.indices[csr.indptr[user]:csr.indptr[user + 1]].astype(numpy.int64)
            mask = numpy.ones(n, dtype=bool)
            mask[rated] = False
            non_rated = all_items[mask]
            numpy.random.shuffle(rated)                       # :162
            share = round((1.0 / k) * len(rated))             # :165 (banker's rounding)
            numpy.random.shuffle(non_rated)                   # :171
            fills = []
            start = 0
            for f in range(k):
                positives = rated[start:] if f == k - 1 else rated[start:start + share]
                start += share
                fills.append(int((n / k) - len(positives)))   # :184
                if f > 0 and fills[f] != fills[f - 1]:        # :185-187
                    extra = non_rated[f * fills[f - 1]:(fills[f - 1] * f) + fills[f]]
                else:                                         # :189
                    extra = non_rated[f * fills[f]:fills[f] * (f + 1)]
                out.append(numpy.append(numpy.array(positives), extra))
                n_positives.append(len(positives))
        self.test_indices = out
        self._slot_positives = (out, numpy.asarray(n_positives, dtype=numpy.int64))
        return out

    def _native_kfold_lists(self, csr):
        """The same lists from the library's host code (csrc/host_split.cuh: numpy's MT19937 stream and legacy shuffle
        restated in C++, ~4x faster than 2 x n_users numpy.random.shuffle calls plus the Python slicing).  None when
        it does not apply (HYPREC_NATIVE_SPLIT=0, library not built, a user whose lists need negative slice bounds):
        the loop below then runs, from the same generator state."""
        if os.environ.get("HYPREC_NATIVE_SPLIT", "1") == "0" or self.n_items >= (1 << 31):
            return None
        try:
            from hyprec_b200 import _native
            got = _native.host_kfold_lists(csr.indptr, csr.indices, self.n_users, self.n_items, self.k_folds)
        except (RuntimeError, OSError, AttributeError):
            return None
        if got is None:
            return None
        slot_ptr, n_positives, ids = got
        lengths = numpy.diff(slot_ptr)
        if len(lengths) and (lengths == lengths[0]).all() and lengths[0] > 0:
            packed = ids[:len(lengths) * int(lengths[0])].reshape(len(lengths), int(lengths[0]))
            out = list(packed)                      # row views: one buffer behind all lists
        else:
            out = [ids[slot_ptr[s]:slot_ptr[s + 1]] for s in range(len(lengths))]
            packed = None
        self.test_indices = out
        self._slot_positives = (out, n_positives)
        self._packed = (out, packed)
        self._flat = (out, slot_ptr, ids)
        return out

    def get_fold(self, fold_num, fold_test_indices):
        slots = [fold_num + user * self.k_folds for user in range(self.n_users)]
        known = getattr(self, '_slot_positives', None)
        if known is not None and known[0] is fold_test_indices:
            # lists made by get_kfold_indices: only their heads (the fold's share of the user's positives) name rated
            # cells, the sampled non-rated items behind them change nothing (evaluator.py:206-213)
            picked = [fold_test_indices[s][:known[1][s]] for s in slots]
        else:
            picked = [fold_test_indices[s] for s in slots]
        return self.generate_kfold_matrix(picked)

    def generate_kfold_matrix(self, test_indices):
        """train = ratings with the listed cells zeroed, test = ratings on the listed cells
        (evaluator.py:206-213).  Only listed cells that hold a rating change anything, so they are found with one
        gather and written with one scatter; the CSR forms of both matrices fall out of the same cell set and are
        kept for ``csr_view`` (the device wants CSR; re-scanning two dense m x n matrices per fold costs more than
        the epochs)."""
        lengths = [len(x) for x in test_indices]
        rows = numpy.repeat(numpy.arange(self.n_users, dtype=numpy.int64), lengths)
        cols = (numpy.concatenate([numpy.asarray(x, dtype=numpy.int64) for x in test_indices])
                if len(rows) else numpy.zeros(0, dtype=numpy.int64))
        if self.sparse:
            return self._split_sparse(rows * self.n_items + cols)
        ratings = numpy.asarray(self.ratings)
        values = ratings[rows, cols]
        rated = values != 0
        rows, cols, values = rows[rated], cols[rated], values[rated]
        train = dense_copy(ratings)
        test = numpy.zeros(ratings.shape)
        test[rows, cols] = values
        train[rows, cols] = 0
        self._remember_csr(train, test, rows * self.n_items + cols)
        return train, test

    # ------------------------------------------------------------------ CSR views of the dense matrices
    def _split_triplets(self, test_keys, only_rated):
        """CSR triplets of (train, test) given the cells (key = user * n_items + item) that move to the test matrix;
        ``only_rated``: the keys may name cells without a rating (k-fold lists), which change nothing."""
        full = self.ratings_csr()
        full_rows = numpy.repeat(numpy.arange(self.n_users, dtype=numpy.int64), numpy.diff(full.indptr))
        full_keys = full_rows * self.n_items + full.indices
        in_test = numpy.zeros(full.nnz, dtype=bool)
        if len(test_keys) and full.nnz:
            # (no numpy.unique over the keys: 19 M of them per fold at the citeulike-a shape, ~15 s in numpy 2.3;
            # a repeated key just sets the same flag twice)
            keys = numpy.asarray(test_keys, dtype=numpy.int64)
            at = numpy.minimum(numpy.searchsorted(full_keys, keys), full.nnz - 1)
            if only_rated:
                at = at[full_keys[at] == keys]
            in_test[at] = True
        out = []
        for mask in (~in_test, in_test):
            indptr = numpy.zeros(self.n_users + 1, dtype=numpy.int64)
            numpy.cumsum(numpy.bincount(full_rows[mask], minlength=self.n_users), out=indptr[1:])
            out.append((indptr, full.indices[mask].astype(numpy.int32), full.data[mask].astype(numpy.float64)))
        return out

    def _remember_csr(self, train, test, test_keys):
        views = self._split_triplets(test_keys, only_rated=False)
        # (the views of the latest TWO folds are kept, the class prepares fold f + 1 while fold f is on the device --
        # unless two more dense users x items matrices alive would be felt, where it does not prefetch either)
        keep = 4 if 16 * self.n_users * self.n_items <= (4 << 30) else 2
        self._csr_views = (getattr(self, '_csr_views', []) + [(train, views[0]), (test, views[1])])[-keep:]

    def _split_sparse(self, test_keys):
        """train / test as scipy CSR matrices, straight from the index lists (sparse ``ratings``)."""
        views = self._split_triplets(test_keys, only_rated=True)
        mats = [sp.csr_matrix((v[2], v[1], v[0]), shape=(self.n_users, self.n_items)) for v in views]
        # (the views of the latest TWO folds are kept: the class prepares fold f + 1 while fold f is on the device)
        self._csr_views = (getattr(self, '_csr_views', []) + [(mats[0], views[0]), (mats[1], views[1])])[-4:]
        return mats[0], mats[1]

    def csr_view(self, matrix):
        """(indptr int64, indices int32, values float64) of ``matrix`` when it is the ratings matrix or one of the
        matrices of the latest fold AND still holds exactly those entries (same number of non-zeros, same values
        at the recorded cells: an in-place edit by the caller is noticed); None otherwise."""
        triplet = None
        if matrix is self.ratings:
            csr = self.ratings_csr()
            triplet = (csr.indptr.astype(numpy.int64), csr.indices.astype(numpy.int32), csr.data.astype(numpy.float64))
        for dense, view in getattr(self, '_csr_views', []):
            if matrix is dense:
                triplet = view
        if triplet is None:
            return None
        indptr, indices, values = triplet
        if sp.issparse(matrix):             # built here from these very arrays; a caller's edit shows in the count
            return triplet if matrix.nnz == len(indices) else None
        if count_nonzero(matrix) != len(indices):
            return None
        rows = numpy.repeat(numpy.arange(matrix.shape[0], dtype=numpy.int64), numpy.diff(indptr))
        if not numpy.array_equal(matrix[rows, indices], values):
            return None
        return triplet

    # ------------------------------------------------------------------ top-k (:215-234)
    def candidate_lists(self, fold):
        """The per-user candidate streams ``load_top_recommendations`` walks (evaluator.py:225),
        including the reference's ``user * (1 + fold)`` slot arithmetic and,
        for the naive split, the dense test rows whose 0./1. values are cast to item ids
        (App. A Q7).  Returns (indptr int64[m+1], ids int32[total])."""
        if sp.issparse(self.test_indices):
            # naive split of sparse ratings: the reference walks the DENSE test row of every user and casts its 0. / 1.
            # values to item ids (App. A Q7) -- n_items entries per user
            if self.n_users * self.n_items > (1 << 31):
                raise MemoryError("the naive split's candidate lists hold n_items entries per user (%d x %d here): call "
                                  "Evaluator.use_hashed_candidates() to rank device-generated lists instead" %
                                  (self.n_users, self.n_items))
            test = self.test_indices
            ids = numpy.zeros(self.n_users * self.n_items, dtype=numpy.int32)
            rows_of = numpy.repeat(numpy.arange(self.n_users, dtype=numpy.int64), numpy.diff(test.indptr))
            ids[rows_of * self.n_items + test.indices] = test.data.astype(numpy.int32)
            slots = numpy.arange(self.n_users, dtype=numpy.int64) * (1 + fold)
            if fold != 0:
                ids = ids.reshape(self.n_users, self.n_items)[slots].ravel()
            return numpy.arange(self.n_users + 1, dtype=numpy.int64) * self.n_items, ids
        packed = self._packed_lists()
        if packed is not None:
            # equally long k-fold lists, kept as one int32 matrix: the fold's streams are a row gather
            slots = numpy.arange(self.n_users, dtype=numpy.int64) * (1 + fold)
            ids = packed[slots].ravel()
            return numpy.arange(self.n_users + 1, dtype=numpy.int64) * packed.shape[1], ids
        flat = getattr(self, '_flat', None)
        if flat is not None and flat[0] is self.test_indices:
            # ragged lists from the native split, still one buffer: gather the fold's slots by range
            slots = numpy.arange(self.n_users, dtype=numpy.int64) * (1 + fold)
            starts, lens = flat[1][slots], flat[1][slots + 1] - flat[1][slots]
            indptr = numpy.zeros(self.n_users + 1, dtype=numpy.int64)
            numpy.cumsum(lens, out=indptr[1:])
            where = numpy.repeat(starts - indptr[:-1], lens) + numpy.arange(indptr[-1], dtype=numpy.int64)
            return indptr, flat[2][where]
        rows = [numpy.asarray(self.test_indices[user * (1 + fold)]).astype(numpy.int64)
                for user in range(self.n_users)]
        indptr = numpy.zeros(self.n_users + 1, dtype=numpy.int64)
        numpy.cumsum([len(r) for r in rows], out=indptr[1:])
        ids = numpy.concatenate(rows).astype(numpy.int32) if rows else numpy.zeros(0, numpy.int32)
        return indptr, ids

    def _packed_lists(self):
        """test_indices as one (slots x length) int32 matrix when it is get_kfold_indices' list of equally long
        arrays (n_items / k_folds ids each, evaluator.py:184), else None.  Built once per list object."""
        lists = self.test_indices
        if not isinstance(lists, list) or not lists:
            return None
        cached = getattr(self, '_packed', None)
        if cached is not None and cached[0] is lists:
            return cached[1]
        length = len(lists[0])
        packed = None
        if all(len(x) == length for x in lists):
            packed = numpy.empty((len(lists), length), dtype=numpy.int32)
            for i, x in enumerate(lists):
                packed[i] = x
        self._packed = (lists, packed)
        return packed

    def load_top_recommendations(self, n_recommendations, predictions, test_data, fold):
        cand_indptr, cand_ids = self.candidate_lists(fold)
        for user in range(self.n_users):
            cand = cand_ids[cand_indptr[user]:cand_indptr[user + 1]].astype(numpy.int64)
            ids, _ = top_k_stream_order(cand, numpy.asarray(predictions[user])[cand], n_recommendations)
            self.recommendation_indices[user] = ids
        self.recs_loaded = True
        self._device_hits = None
        return self.recommendation_indices

    def set_device_results(self, recommendation_indices, hits):
        """Install top-k lists and hit flags computed on the device."""
        self.recommendation_indices = recommendation_indices
        self._device_hits = hits
        self.recs_loaded = True

    # ------------------------------------------------------------------ metrics (:236-341)
    def get_rmse(self, predicted, actual=None):
        if actual is None:
            actual = self.ratings
        if sp.issparse(actual):
            actual = numpy.asarray(actual.todense())
        rss = 0
        for i in range(predicted.shape[0]):
            rss += numpy.sum((predicted[i] - actual[i]) ** 2)
        return numpy.sqrt(float(rss) / numpy.size(predicted))

    def calculate_recall(self, ratings, predictions):
        if sp.issparse(ratings):
            rows, cols = ratings.nonzero()
            return sum(numpy.asarray(predictions)[rows, cols]) / numpy.float64(ratings.sum())
        denom = sum(sum(ratings))
        return sum(predictions[ratings.nonzero()]) / denom

    def _hits(self, rounded_predictions):
        if self._device_hits is not None and rounded_predictions is None:
            return self._device_hits
        def row(user):
            if self.sparse:
                return numpy.asarray(self.ratings[user].todense()).ravel()
            return numpy.asarray(self.ratings[user])
        return [row(user)[numpy.asarray(recs, dtype=numpy.int64)] *
                numpy.asarray(rounded_predictions[user])[numpy.asarray(recs, dtype=numpy.int64)]
                for user, recs in enumerate(self.recommendation_indices)]

    def recall_at_x(self, x, predictions, ratings, rounded_predictions):
        hits = self._hits(rounded_predictions)
        likes = numpy.asarray(ratings.sum(axis=1)).ravel()
        recalls = [numpy.sum(hits[u][:x]) / (min(x, likes[u]) * 1.0)
                   for u in range(len(hits)) if likes[u] != 0]
        return numpy.mean(recalls, dtype=numpy.float16)

    def calculate_ndcg(self, n_recommendations, predictions, test_data, rounded_predictions):
        hits = self._hits(rounded_predictions)
        logs = numpy.log2(numpy.arange(max(n_recommendations, 1)) + 2)
        ndcgs = []
        for h in hits:
            top = min(len(h), n_recommendations)
            if top == 0:
                continue
            # sequential accumulation, as the reference's += loop (:308-312)
            dcg = numpy.cumsum(numpy.asarray(h[:top], dtype=numpy.float64) / logs[:top])[-1]
            idcg = numpy.cumsum(1 / logs[:top])[-1]
            if idcg != 0:
                ndcgs.append(dcg / idcg)
        return numpy.mean(ndcgs, dtype=numpy.float16)

    def calculate_mrr(self, n_recommendations, predictions, test_data, rounded_predictions):
        hits = self._hits(rounded_predictions)
        mrr_list = []
        for h in hits:
            first = numpy.nonzero(numpy.asarray(h[:n_recommendations]) == 1)[0]
            mrr_list.append(1.0 / (first[0] + 1) if len(first) else 0)
        return numpy.mean(mrr_list, dtype=numpy.float16)
