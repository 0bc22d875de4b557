"""
Drop-in replacement for the reference's ``CollaborativeFiltering``
(/root/reference/lib/collaborative_filtering.py:13-302) whose heavy lifting runs on a B200
through the C ABI of include/hyprec_b200.h.

Same constructor (positional order of collaborative_filtering.py:18-20), attributes
(``user_vecs``, ``item_vecs``, ``train_data``, ``test_data``, ``n_factors``, ``hyperparameters``,
``predictions``, ``document_distribution`` ...) and methods, so ``RecommenderSystem``
(lib/recommender_system.py:97-101), ``GridSearch`` (lib/grid_search.py:63-75) and ``runnables.py``
(:115-118) can use it unchanged; ``hyprec_b200.install()`` rebinds the reference's module
attribute to this class.

What runs where:
  * ``partial_train`` / ``als_step``      -> hyprec_train / hyprec_als_half_step (device)
  * ``get_predictions``                   -> hyprec_predict_dense (device GEMM, dense result on host)
  * ``get_evaluation_report``             -> hyprec_eval_topk + hyprec_rmse (device); the float16 means
                                             are taken on the host with the reference's numpy call
  * splits, configuration, checkpoints    -> host (evaluator / initializer objects passed in)
There is no CPU fallback: without the CUDA library or a GPU the first device call raises.

Environment knobs (the JSON config surface is unchanged): ``HYPREC_PRECISION`` = fp64 (default,
parity with numpy float64) | fp32; ``HYPREC_DEVICE`` = CUDA device index (default: LOCAL_RANK or 0).

Multi-GPU: when the process runs under ``torchrun`` with an initialised ``torch.distributed`` process
group (one process per GPU), every rank executes the same host code on the same data; the epochs shard
users / items across ranks (``hyprec_b200.distributed.ShardedALS``), the evaluation shards users and
gathers the per-user integers, and every rank ends up with identical factors and an identical report.
"""
import os
import time
import zlib

import numpy
import scipy.sparse as sp

from hyprec_b200 import _native, metrics
from hyprec_b200.abstract_recommender import AbstractRecommender, REPORT_FORMAT


def _precision_from_env():
    name = os.environ.get("HYPREC_PRECISION", "fp64").lower()
    if name in ("fp64", "f64", "float64", "0"):
        return _native.F64
    if name in ("fp32", "f32", "float32", "1"):
        return _native.F32
    raise ValueError("HYPREC_PRECISION must be fp64 or fp32, got %r" % name)


def _total(matrix):
    """``sum(sum(matrix))`` of the reference (abstract_recommender.py:115) for dense or sparse matrices."""
    if sp.issparse(matrix):
        return numpy.float64(matrix.sum())
    return sum(sum(matrix))


# get_predictions() hands out a dense users x items float64 array; above this size it raises instead
DENSE_PREDICTIONS_LIMIT_BYTES = int(os.environ.get("HYPREC_DENSE_LIMIT_BYTES", 8 << 30))


def _csr_triplet(matrix):
    """(indptr int64, indices int32, values float64) of a dense or scipy-sparse matrix, zeros
    dropped, column ids sorted."""
    if sp.issparse(matrix):
        csr = sp.csr_matrix(matrix, dtype=numpy.float64)
        csr.eliminate_zeros()
        csr.sort_indices()
    else:
        from hyprec_b200.evaluator import dense_to_csr
        csr = dense_to_csr(numpy.asarray(matrix, dtype=numpy.float64))
    return csr.indptr.astype(numpy.int64), csr.indices.astype(numpy.int32), csr.data.astype(numpy.float64)


def _crc_pool():
    from hyprec_b200.evaluator import host_pool
    return host_pool()


class CollaborativeFiltering(AbstractRecommender):
    n_recommendations = 200          # abstract_recommender.py:111
    recall_cutoff = 200              # abstract_recommender.py:114

    def __init__(self, initializer, evaluator, hyperparameters, options,
                 verbose=False, load_matrices=True, dump_matrices=True, train_more=True,
                 is_hybrid=False, update_with_items=False, init_with_content=True):
        self.initializer = initializer
        self.evaluator = evaluator
        self.ratings = evaluator.get_ratings()
        self.n_users, self.n_items = self.ratings.shape
        self.k_folds = None
        self.prediction_fold = -1

        self._verbose = verbose
        self._load_matrices = load_matrices
        self._dump_matrices = dump_matrices
        self._train_more = train_more
        self._is_hybrid = is_hybrid
        self._update_with_items = update_with_items
        self._split_type = 'user'
        self._init_with_content = init_with_content

        self.document_distribution = None
        self._handle = None
        self._precision = _precision_from_env()
        self._train_uploaded = None      # the train_data object the device CSR was built from
        self._eval_uploaded = None       # (ratings, test_data) objects of the uploaded eval CSRs
        self._train_nnz = 0
        self._test_nnz = 0
        self._cands_uploaded = None      # (test_indices object, fold) of the candidate lists on the device
        self._train_pattern = None       # (indptr, indices) of the uploaded train / test CSRs
        self._test_pattern = None
        self._sync_token = None
        self.last_report_details = None  # per-user device outputs of the latest report

        self.set_hyperparameters(hyperparameters)
        self.set_options(options)

    # ------------------------------------------------------------------ configuration
    def set_hyperparameters(self, hyperparameters):
        self.n_factors = hyperparameters['n_factors']
        self._lambda = hyperparameters['_lambda']
        self.predictions = None
        self.hyperparameters = hyperparameters.copy()

    def set_item_based_recommender(self, recommender):
        self.item_based_recommender = recommender

    def _get_options_suffix(self):
        suffix = ('i' if self._init_with_content else '') + ('u' if self._update_with_items else '')
        return '_' + suffix if suffix else ''

    # ------------------------------------------------------------------ device state
    def device(self):
        """The C-ABI handle (created on first use; raises without the CUDA library / a GPU)."""
        if self._handle is None:
            index = int(os.environ.get("HYPREC_DEVICE", os.environ.get("LOCAL_RANK", "0")))
            self._handle = _native.Handle(index, self._precision)
        return self._handle

    def _world(self):
        """(rank, world_size) of an initialised torch.distributed group, else (0, 1).  An object
        marked ``_single_rank`` (a grid-search worker that owns one whole model, grid_search.py) stays
        on its own GPU even inside a process group."""
        if getattr(self, '_single_rank', False):
            return 0, 1
        import sys
        if 'torch.distributed' not in sys.modules:      # nobody can have initialised a group (importing torch costs ~2 s)
            return 0, 1
        try:
            import torch.distributed as dist
        except ImportError:
            return 0, 1
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist.get_rank(), dist.get_world_size()
        return 0, 1

    def close(self):
        if self._handle is not None:
            self._handle.close()
            self._handle = None
            self._train_uploaded = self._eval_uploaded = self._sync_token = self._cands_uploaded = None
        pool = self.__dict__.pop('_ahead_pool', None)
        if pool is not None:
            pool.shutdown(wait=True)
        self._cands_ready = None

    def set_data(self, train_data, test_data):
        AbstractRecommender.set_data(self, train_data, test_data)
        self._sync_token = None

    def _host_csr(self, matrix):
        """CSR triplet of a dense / sparse users x items matrix: the evaluator's own view when it built the
        matrix (hyprec_b200.Evaluator.csr_view, validated against the array), else a scan."""
        return self._facts(matrix)[0]

    def _facts(self, matrix):
        """(CSR triplet, sum of all entries, row sums) of one of the API's users x items matrices, computed once per
        matrix OBJECT: ratings / train / test are scanned by every report of the reference (test_data.sum(),
        sum(sum(ratings)), ...), which costs more than the device work at the citeulike shape and is repeated for
        every model of a sweep.  A cached entry is re-used only for the same object, and only while a sample of
        its cells still reads as recorded (an in-place edit by the caller drops it)."""
        cache = self.__dict__.setdefault('_matrix_facts', [])
        for at, entry in enumerate(cache):
            if entry[0] is matrix:
                triplet = entry[1]
                if sp.issparse(matrix):
                    if matrix.nnz == len(triplet[1]):
                        return entry[1:4]
                elif len(triplet[1]) == 0 or numpy.array_equal(matrix[entry[4], triplet[1][entry[5]]], triplet[2][entry[5]]):
                    return entry[1:4]
                del cache[at]             # (list.remove would compare the arrays inside the tuples)
                break
        view = getattr(self.evaluator, 'csr_view', None)
        triplet = view(matrix) if view is not None and (isinstance(matrix, numpy.ndarray) or sp.issparse(matrix)) else None
        if triplet is None:
            triplet = _csr_triplet(matrix)
        indptr, indices, values = triplet
        rows = numpy.repeat(numpy.arange(len(indptr) - 1, dtype=numpy.int64), numpy.diff(indptr))
        row_sums = numpy.bincount(rows, weights=values, minlength=len(indptr) - 1) if len(values) else numpy.zeros(len(indptr) - 1)
        pick = numpy.linspace(0, max(len(indices) - 1, 0), num=min(len(indices), 4096)).astype(numpy.int64)
        entry = (matrix, triplet, numpy.float64(values.sum()), row_sums, rows[pick], pick)
        cache.append(entry)
        if len(cache) > 8:
            cache.pop(0)
        return entry[1:4]

    def _upload_train(self):
        h = self.device()
        if self._train_uploaded is not self.train_data:
            indptr, indices, values = self._host_csr(self.train_data)
            h.set_train_csr(self.n_users, self.n_items, indptr, indices, values)
            self._train_nnz = len(indices)
            self._train_pattern = (indptr, indices)      # order of the device's train flags
            self._train_uploaded = self.train_data
            self._eval_uploaded = None
            self._cands_uploaded = None
            self._sync_token = None
        return h

    def _upload_eval(self):
        h = self._upload_train()
        key = (self.ratings, self.test_data)
        if self._eval_uploaded is None or self._eval_uploaded[0] is not key[0] or self._eval_uploaded[1] is not key[1]:
            full = self._host_csr(self.ratings)
            test = self._host_csr(self.test_data)
            h.set_eval_csr(full, test)
            # hit values (rating x rounded prediction) fit one byte when the ratings are small integers -- the
            # reference's data is 0 / 1 -- and then travel that way (hyprec_eval_topk_u8)
            values = full[2]
            self._hits_u8 = bool(len(values) == 0 or (values.min() >= 0 and values.max() <= 255 and
                                                      numpy.array_equal(values, numpy.floor(values))))
            self._test_nnz = len(test[1])
            self._test_pattern = (test[0], test[1])      # order of the device's test flags
            self._eval_uploaded = key
        return h

    @staticmethod
    def _fingerprint(user_vecs, item_vecs):
        """Witness that the host factor arrays still hold what the device holds: CRC-32s of the whole buffer in 4 MB
        pieces (position-sensitive), so an in-place edit anywhere -- an external als_step loop, SDAE writing rows, a row
        permutation -- is noticed and the factors are uploaded again (the sampled witness of round 1 missed edits
        between its ~500 probes).  The pieces are checksummed on a few threads (zlib releases the GIL): ~2 ms for the
        36 MB of the citeulike-a factors instead of ~9."""
        def pieces(a):
            a = numpy.asarray(a)
            if not a.flags.c_contiguous:
                a = numpy.ascontiguousarray(a)
            raw = memoryview(a.reshape(-1).view(numpy.uint8)) if a.size else memoryview(b"")
            step = 4 << 20
            return (a.shape, a.dtype.str), [raw[at:at + step] for at in range(0, len(raw), step)]
        (head_u, chunks_u), (head_v, chunks_v) = pieces(user_vecs), pieces(item_vecs)
        chunks = chunks_u + chunks_v
        if len(chunks) > 2:
            sums = list(_crc_pool().map(zlib.crc32, chunks))
        else:
            sums = [zlib.crc32(c) for c in chunks]
        return head_u, head_v, tuple(sums)

    def _upload_factors(self, force=False):
        h = self._upload_train()
        token = self._fingerprint(self.user_vecs, self.item_vecs)
        if force or token != self._sync_token:
            h.set_factors(self.user_vecs, self.item_vecs)
            self._sync_token = token
        return h

    def _download_factors(self, h):
        """Write the device factors INTO the existing host arrays: ``item_vecs`` may alias the
        caller's theta, which the reference overwrites in place."""
        def target(arr):
            ok = arr.dtype == numpy.float64 and arr.flags.c_contiguous and arr.flags.writeable
            return arr if ok else numpy.empty(arr.shape)
        u_out, v_out = target(self.user_vecs), target(self.item_vecs)
        h.get_factors(u_out, v_out)
        if u_out is not self.user_vecs:
            self.user_vecs[...] = u_out
        if v_out is not self.item_vecs:
            self.item_vecs[...] = v_out
        self._sync_token = self._fingerprint(self.user_vecs, self.item_vecs)

    # ------------------------------------------------------------------ ALS
    def build_confidence_matrix(self, index, type='user'):
        """collaborative_filtering.py:120-144: 1.0 on train non-zeros of the row/column, else 0.1."""
        train = self.train_data
        line = train[index] if type == 'user' else train[:, index]
        if sp.issparse(line):
            line = numpy.asarray(line.todense()).ravel()
        return numpy.where(numpy.asarray(line) != 0, 1.0, 0.1)

    def als_step(self, latent_vectors, fixed_vecs, ratings, _lambda, type='user'):
        """One half-step (collaborative_filtering.py:69-99): ``latent_vectors`` is overwritten in
        place and returned.  As in the reference the confidence pattern comes from
        ``self.train_data``; ``ratings`` must be that same matrix (the only way the reference and
        SDAE.py:218-219 call it)."""
        if ratings is not self.train_data:
            same = ratings.shape == self.train_data.shape and (
                (ratings != self.train_data).nnz == 0 if sp.issparse(ratings) or sp.issparse(self.train_data)
                else numpy.array_equal(ratings, self.train_data))
            if not same:
                raise NotImplementedError("als_step: ratings other than self.train_data are not supported "
                                          "on the device path")
        if type not in ('user', 'item'):
            return latent_vectors       # the reference does nothing for other values (:79-99)
        h = self._upload_train()
        if type == 'user':
            h.set_factors(latent_vectors, fixed_vecs)
            h.set_item_prior(None)
            h.als_half_step(_native.SIDE_USER, _lambda)
            solved, _ = h.get_factors(None, False)
        else:
            h.set_factors(fixed_vecs, latent_vectors)
            prior = self.document_distribution if self._update_with_items else None
            h.set_item_prior(prior)
            h.als_half_step(_native.SIDE_ITEM, _lambda)
            _, solved = h.get_factors(False, None)
        latent_vectors[...] = solved
        self._sync_token = None
        return latent_vectors

    def train(self, item_vecs=None):
        """collaborative_filtering.py:102-118."""
        self.document_distribution = item_vecs.copy() if item_vecs is not None else None
        if self.splitting_method == 'naive':
            self.set_data(*self.evaluator.naive_split(self._split_type))
            self.hyperparameters['fold'] = 0
            return self.train_one_fold(item_vecs)
        self.fold_test_indices = self.evaluator.get_kfold_indices()
        return self.train_k_fold(item_vecs)

    def train_k_fold(self, item_vecs=None):
        """collaborative_filtering.py:147-161."""
        all_errors = []
        ahead = None
        for current_k in range(self.k_folds):
            data = ahead.result() if ahead is not None else self.evaluator.get_fold(current_k, self.fold_test_indices)
            ahead = self._prepare_ahead(current_k + 1) if current_k + 1 < self.k_folds else None
            self.set_data(*data)
            self.hyperparameters['fold'] = current_k
            all_errors.append(self.train_one_fold(item_vecs))
            self.predictions = None
        return numpy.mean(all_errors, axis=0)

    def _prepare_ahead(self, fold):
        """Host work of fold ``fold`` that depends on nothing the current fold computes -- its train / test matrices
        (evaluator.py:120-134, 197-213) and the candidate streams its report will rank -- done on a host thread while
        the current fold's epochs and report run on the device (the library calls release the GIL).  Nothing here
        touches the numpy.random stream: the initial draws stay on the caller's thread, in the reference's order.
        Returns a future of get_fold's (train, test), or None when prefetching is off (HYPREC_PREFETCH=0) or the
        evaluator is not this package's."""
        if os.environ.get("HYPREC_PREFETCH", "1") == "0" or not hasattr(self.evaluator, 'candidate_lists') \
                or getattr(self.evaluator, 'hashed_candidates', None) is not None:
            return None
        if not sp.issparse(self.ratings) and 16 * self.n_users * self.n_items > (4 << 30):
            return None                     # two more dense users x items matrices alive at once: not worth the memory
        lists = self.fold_test_indices

        def work():
            data = self.evaluator.get_fold(fold, lists)
            if self.evaluator.test_indices is lists and self._world()[1] == 1:
                self._cands_ready = (lists, fold, self.evaluator.candidate_lists(fold))
            return data
        from concurrent.futures import ThreadPoolExecutor
        pool = self.__dict__.get('_ahead_pool')
        if pool is None:
            pool = self._ahead_pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="hyprec-fold")
        try:
            return pool.submit(work)
        except RuntimeError:                # shut down by close() of a shallow copy of this object
            pool = self._ahead_pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="hyprec-fold")
            return pool.submit(work)

    def train_one_fold(self, item_vecs=None):
        """collaborative_filtering.py:164-212: initialise (random / theta / checkpoint), train,
        optionally save, report.  The order of the ``numpy.random`` draws is the reference's
        (user matrix first)."""
        return self._run_fold(self._init_fold(item_vecs))

    def _init_fold(self, item_vecs=None):
        """Host half of train_one_fold (:166-192): everything that touches the global numpy.random
        stream or the checkpoint files.  Returns ``matrices_found``."""
        user_shape, item_shape = (self.n_users, self.n_factors), (self.n_items, self.n_factors)
        matrices_found = False
        if self._load_matrices is False:
            self.user_vecs = numpy.random.random(user_shape)
            use_content = item_vecs is not None and self._init_with_content and item_vecs.shape == item_shape
            self.item_vecs = item_vecs if use_content else numpy.random.random(item_shape)
        else:
            suffix = self._get_options_suffix()
            users_found, self.user_vecs = self.initializer.load_matrix(self.hyperparameters, 'user_vecs' + suffix,
                                                                       user_shape)
            if self._verbose and users_found:
                print("User distributions files were found.")
            items_found, self.item_vecs = self.initializer.load_matrix(self.hyperparameters, 'item_vecs' + suffix,
                                                                       item_shape)
            if self._verbose and items_found:
                print("Document distributions files were found.")
            if not items_found and item_vecs is not None:
                items_found = True
                self.item_vecs = item_vecs
            matrices_found = users_found and items_found
        self._sync_token = None
        return matrices_found

    def _run_fold(self, matrices_found):
        """Device half of train_one_fold (:194-212): train (or not), save, report."""
        if not matrices_found:
            if self._verbose and self._load_matrices:
                print("User and Document distributions files were not found, will train collaborative.")
            self.partial_train()
        elif self._train_more:
            if self._verbose and self._load_matrices:
                print("User and Document distributions files found, will train model further.")
            self.partial_train()
        elif self._verbose and self._load_matrices:
            print("User and Document distributions files found, will not train the model further.")

        if self._dump_matrices:
            self.initializer.set_config(self.hyperparameters, self.n_iter)
            self.initializer.save_matrix(self.user_vecs, 'user_vecs' + self._get_options_suffix())
            self.initializer.save_matrix(self.item_vecs, 'item_vecs' + self._get_options_suffix())
        return self.get_evaluation_report()

    def partial_train(self):
        """collaborative_filtering.py:224-259, device-resident: factors go up once, the epochs
        (user half-step, item half-step, RMSE, ``error >= old_error`` break) run inside
        hyprec_train, and the result is written back into the host arrays."""
        h = self._upload_factors(force=True)
        use_prior = self._update_with_items and self.document_distribution is not None
        h.set_item_prior(self.document_distribution if use_prior else None)
        fold = self.hyperparameters['fold'] + 1 if 'fold' in self.hyperparameters else 0
        prefix = 'Fold:{:02d} '.format(fold) if fold != 0 else ''
        line = prefix + 'Epoch:{epoch:02d} Loss:{loss:1.4e} Time:{time:.3f}s'
        error = numpy.inf
        if self._verbose:
            error = h.rmse()
            print(line.format(epoch=0, loss=error, time=0))
        state = {'previous': error, 'stalled': False}
        rank, world = self._world()
        chatty = self._verbose and rank == 0

        def on_epoch(epoch, rmse, seconds):
            if chatty:
                print(line.format(epoch=epoch, loss=rmse, time=seconds))
            state['stalled'] = rmse >= state['previous']
            state['previous'] = rmse

        if world == 1:
            h.train(self.n_iter, self._lambda, error, on_epoch)
        else:
            self._train_sharded(h, rank, world, error, on_epoch)
        if chatty and state['stalled']:
            print("Local Optimum was found in the last iteration, breaking.")
        self._download_factors(h)

    def _shard_bounds(self, world):
        from hyprec_b200 import distributed
        indptr, indices, _ = self._host_csr(self.train_data)
        users = distributed.partition_rows(indptr, world, self.n_factors)
        items = distributed.partition_rows(distributed.csc_indptr(indptr, indices, self.n_items), world, self.n_factors)
        return users, items

    def _train_sharded(self, h, rank, world, error, on_epoch):
        """The epoch loop of partial_train over row shards.  With NVLink peer memory the library runs the whole
        loop (hyprec_train: solved rows stored into every replica from the kernels' epilogues, Gram and RMSE
        reduced through the mailbox, the same early-stop decision on every rank); otherwise one
        torch.distributed exchange of the solved rows after each half-step."""
        from hyprec_b200 import distributed
        users, items = self._shard_bounds(world)
        device = int(os.environ.get("HYPREC_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        als = distributed.make_sharded(h, device, rank, world, users, items, self.n_users, self.n_items)
        if isinstance(als, distributed.NativeShardedALS):
            try:
                h.train(self.n_iter, self._lambda, error, on_epoch)
            finally:
                distributed.detach_exchange(h)
            return
        try:
            for epoch in range(1, self.n_iter + 1):
                old_error = error
                t0 = time.time()
                als.half_step(_native.SIDE_USER, self._lambda)
                als.half_step(_native.SIDE_ITEM, self._lambda)
                t1 = time.time()
                error = als.rmse()
                on_epoch(epoch, error, t1 - t0)
                if error >= old_error:
                    break
        finally:
            h.set_shard(0, -1, 0, -1)          # back to "all rows" for later single-rank calls

    # ------------------------------------------------------------------ prediction
    def get_predictions(self):
        """collaborative_filtering.py:262-284; the dense product is computed on the device."""
        if self.predictions is None or not self.prediction_fold == self.hyperparameters['fold']:
            need = 8 * self.n_users * self.n_items
            if need > DENSE_PREDICTIONS_LIMIT_BYTES:
                raise MemoryError("get_predictions(): the dense %d x %d float64 score matrix needs %.1f GB (limit "
                                  "HYPREC_DENSE_LIMIT_BYTES = %.1f GB); the evaluation report, recommend_items and "
                                  "dump_recommendations do not need it" %
                                  (self.n_users, self.n_items, need / 1e9, DENSE_PREDICTIONS_LIMIT_BYTES / 1e9))
            h = self._upload_factors()
            collaborative_predictions = h.predict_dense()
            if self._is_hybrid:
                from hyprec_b200.linear_regression import LinearRegression
                self.item_based_recommender.set_data(self.train_data, self.test_data)
                regr = LinearRegression(self.train_data, self.test_data,
                                        self.item_based_recommender.get_predictions(), collaborative_predictions)
                self.predictions = regr.train()
                if self._verbose:
                    print("returned linear regression ratings")
            else:
                self.predictions = collaborative_predictions
            self.prediction_fold = self.hyperparameters['fold']
        return self.predictions

    def predict(self, user, item):
        return self.user_vecs[user, :].dot(self.item_vecs[item, :].T)

    def recommend_items(self, user_id, num_recommendations=10):
        """abstract_recommender.py:189-204 without the dense matrix: full-catalogue top-N of one user
        on the device, returned ascending like TopRecommendations' lists."""
        if self._is_hybrid:
            return AbstractRecommender.recommend_items(self, user_id, num_recommendations)
        h = self._upload_factors()
        if self.n_users == 1:
            self._drop_candidates(h)         # the one-user range would match the fold's cached lists
        out = h.eval_topk(num_recommendations, None, None, user_id, user_id + 1, want_counts=False)
        keep = out["topk_idx"][0] >= 0
        ids = [int(i) for i in out["topk_idx"][0][keep]]
        vals = [float(v) for v in out["topk_val"][0][keep]]
        return zip(reversed(ids), reversed(vals))

    def _drop_candidates(self, h):
        """Forget the fold's candidate lists on the device: a NULL-list top-k over the same user range would
        use them instead of the whole catalogue (include/hyprec_b200.h, hyprec_set_candidates)."""
        if self._cands_uploaded is not None:
            h.set_candidates(None, None)
            self._cands_uploaded = None

    def top_recommendations(self, num_recommendations=10, user_lo=0, user_hi=None):
        """Full-catalogue top-N of users [user_lo, user_hi) in ONE device pass (what
        abstract_recommender.py:189-204 computes user by user): (ids int32, scores float64), users x N,
        descending score with TopRecommendations' tie order, -1 padded."""
        h = self._upload_factors()
        self._drop_candidates(h)
        user_hi = self.n_users if user_hi is None else user_hi
        out = h.eval_topk(num_recommendations, None, None, user_lo, user_hi, want_counts=False)
        return out["topk_idx"], out["topk_val"]

    def dump_recommendations(self, num_recommendations=10, path=None):
        """abstract_recommender.py:148-171 with one batched device pass instead of n_users x n_items Python
        inserts.  ``path`` defaults to ``matrices/<results_file_name>`` like the reference
        (``results_file_name`` is 'top_recommendations.dat' unless the owner set one,
        recommender_system.py:27, 59)."""
        if self._is_hybrid:
            return AbstractRecommender.dump_recommendations(self, num_recommendations)
        from hyprec_b200 import formats
        ids, vals = self.top_recommendations(num_recommendations)
        if path is None:
            base_dir = os.path.dirname(os.path.realpath(__file__))
            name = getattr(self, 'results_file_name', 'top_recommendations.dat')
            path = os.path.join(os.path.dirname(base_dir), 'matrices/%s' % name)
            os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            f.writelines(formats.recommendation_lines(ids, vals))
        if self._verbose:
            print("dumped top recommendations to %s" % path)
        return path

    # ------------------------------------------------------------------ evaluation
    def candidate_lists(self, fold):
        """The candidate stream of every user exactly as load_top_recommendations reads it:
        ``evaluator.test_indices[user * (1 + fold)]`` cast to int (evaluator.py:225-228; SURVEY.md
        App. A Q6 / Q7 describe what that slot holds for k-fold and naive splits)."""
        if hasattr(self.evaluator, 'candidate_lists'):         # hyprec_b200.Evaluator (dense or sparse test_indices)
            ready = self.__dict__.get('_cands_ready')
            if ready is not None and ready[0] is self.evaluator.test_indices and ready[1] == fold:
                self._cands_ready = None                        # gathered while the previous fold was on the device
                return ready[2]
            return self.evaluator.candidate_lists(fold)
        rows = [numpy.asarray(self.evaluator.test_indices[user * (1 + fold)]).astype(numpy.int64)
                for user in range(self.n_users)]
        indptr = numpy.zeros(self.n_users + 1, dtype=numpy.int64)
        numpy.cumsum([len(r) for r in rows], out=indptr[1:])
        ids = numpy.concatenate(rows).astype(numpy.int32) if len(rows) else numpy.zeros(0, numpy.int32)
        return indptr, ids

    def _eval_sharded(self, h, rank, world, fold):
        """Users sharded across ranks (V is replicated); the per-user outputs are gathered so that
        every rank holds the full arrays."""
        import torch.distributed as dist
        users, _ = self._shard_bounds(world)
        lo, hi = users[rank], users[rank + 1]
        train_ptr = self._host_csr(self.train_data)[0]
        test_ptr = self._host_csr(self.test_data)[0]
        hashed = getattr(self.evaluator, 'hashed_candidates', None)
        if hashed is not None:
            h.set_hashed_candidates(fold % hashed[0], hashed[0], hashed[1])
            cp = ci = None
        else:
            cand_indptr, cand_ids = self.candidate_lists(fold)
            cp, ci = cand_indptr[lo:hi + 1] - cand_indptr[lo], cand_ids[cand_indptr[lo]:cand_indptr[hi]]
        if hashed is not None:
            part = self._eval_blocks(h, lo, hi, hashed[0], train_ptr, test_ptr)
        else:
            part = h.eval_topk(self.n_recommendations, cp, ci, lo, hi,
                               n_train_nz=int(train_ptr[hi] - train_ptr[lo]), n_test_nz=int(test_ptr[hi] - test_ptr[lo]))
        self._cands_uploaded = None
        gathered = [None] * world
        dist.all_gather_object(gathered, part)
        return {key: numpy.concatenate([g[key] for g in gathered]) for key in part}

    def _naive_lists(self, h, blended):
        """Evaluation under the NAIVE split: ``evaluator.test_indices`` is the test MATRIX, whose rows the reference
        walks as candidate streams of item ids 0 / 1 (evaluator.py:97, :225-228).  The lists
        then follow from two scores per user in closed form (top_recommendations.naive_split_recommendations); the
        device supplies thresholds, rounded counts and the flags at the positives, and no n_items-long stream is
        built or uploaded.  Returns the eval_topk-shaped dict, or None when a test value is not 0 / 1."""
        from hyprec_b200.top_recommendations import naive_split_recommendations
        test_ptr, test_idx, test_val = self._host_csr(self.test_data)
        w_u, w_v = blended if blended is not None else (self.user_vecs, self.item_vecs)
        if self.n_items < 2:
            return None
        pair = numpy.asarray(w_u, dtype=numpy.float64).dot(numpy.asarray(w_v[:2], dtype=numpy.float64).T)
        got = naive_split_recommendations(pair[:, 0], pair[:, 1], test_ptr, test_idx, test_val, self.n_items,
                                          self.n_recommendations)
        if got is None:
            return None
        ids, lengths = got
        rank, world = self._world()
        if world == 1:
            out = h.eval_topk(self.n_recommendations, None, None, 0, self.n_users, want_topk=False,
                              n_train_nz=self._train_nnz, n_test_nz=self._test_nnz)
        else:
            import torch.distributed as dist
            users, _ = self._shard_bounds(world)
            lo, hi = users[rank], users[rank + 1]
            train_ptr = self._host_csr(self.train_data)[0]
            part = h.eval_topk(self.n_recommendations, None, None, lo, hi, want_topk=False,
                               n_train_nz=int(train_ptr[hi] - train_ptr[lo]), n_test_nz=int(test_ptr[hi] - test_ptr[lo]))
            gathered = [None] * world
            dist.all_gather_object(gathered, {k: v for k, v in part.items() if v is not None})
            out = {key: numpy.concatenate([g[key] for g in gathered]) for key in gathered[0]}
        full_ptr, full_idx, full_val = self._host_csr(self.ratings)
        rated = numpy.zeros((self.n_users, 2))                     # ratings[u][0], ratings[u][1]
        rows = numpy.repeat(numpy.arange(self.n_users), numpy.diff(full_ptr))
        low = full_idx < 2
        rated[rows[low], full_idx[low]] = full_val[low]
        rounded = (pair >= numpy.asarray(out["thresholds"])[:, None]).astype(numpy.float64)
        safe = numpy.maximum(ids, 0)
        users = numpy.arange(self.n_users)[:, None]
        out["topk_idx"] = ids
        out["topk_val"] = numpy.where(ids >= 0, pair[users, safe], 0.0)
        out["topk_hit"] = numpy.where(ids >= 0, rated[users, safe] * rounded[users, safe], 0.0)
        return out

    def _eval_blocks(self, h, lo, hi, n_folds, train_ptr=None, test_ptr=None):
        """Device-generated candidate lists hold ~n_items / n_folds ids per user: users [lo, hi) are evaluated in
        blocks whose lists stay below ~2 GB on the device, and the per-user outputs are concatenated."""
        train_ptr = self._train_pattern[0] if train_ptr is None else train_ptr
        test_ptr = self._test_pattern[0] if test_ptr is None else test_ptr
        per_user = max(1, 4 * self.n_items // max(n_folds, 1))
        block = int(max(128, min(hi - lo, (2 << 30) // per_user)))
        parts = []
        for b_lo in range(lo, hi, block):
            b_hi = min(hi, b_lo + block)
            parts.append(h.eval_topk(self.n_recommendations, None, None, b_lo, b_hi,
                                     n_train_nz=int(train_ptr[b_hi] - train_ptr[b_lo]),
                                     n_test_nz=int(test_ptr[b_hi] - test_ptr[b_lo])))
        if len(parts) == 1:
            return parts[0]
        return {key: numpy.concatenate([p[key] for p in parts]) for key in parts[0]}

    def recall_at_cutoffs(self, cutoffs=(50, 100, 150, 200, 250, 300)):
        """recall@x for every x of ``cutoffs`` from ONE device pass (BASELINE.json configs[1]: recall@{50..300}).
        The reference loads 200 recommendations and reports recall@200 only (abstract_recommender.py:111-114);
        here the streamed top-max(cutoffs) list of each user is loaded once and each cut-off is evaluated with
        the reference's own arithmetic (evaluator.py:281-292).  Returns {x: numpy.float16}."""
        cutoffs = [int(x) for x in cutoffs]
        saved = self.n_recommendations
        # the longer lists are this call's own: what the evaluator and last_report_details held before (the
        # reference's 200-long lists of the last report) is put back, so later metric calls see the same data
        ev = self.evaluator
        before = (ev.recommendation_indices, ev.recs_loaded, getattr(ev, '_device_hits', None), self.last_report_details)
        self.n_recommendations = max([saved] + cutoffs)
        try:
            verbose, self._verbose = self._verbose, False
            self.get_evaluation_report()
            hits = self.last_report_details["topk_hit"]
        finally:
            self.n_recommendations, self._verbose = saved, verbose
            ev.recommendation_indices, ev.recs_loaded, self.last_report_details = before[0], before[1], before[3]
            if hasattr(ev, '_device_hits'):
                ev._device_hits = before[2]
        test_likes = self._facts(self.test_data)[2]
        return {x: metrics.recall_at_x(x, hits, test_likes) for x in cutoffs}

    def _report_factors(self):
        """Hybrid recommender (``is_hybrid``): the blended scores of collaborative_filtering.py:271-279 as one
        low-rank product ``W_u W_v^T`` (hyprec_b200/hybrid.py), so that the report runs on the device like the
        plain collaborative one instead of on three dense m x n matrices.  The default whenever the content model
        exposes the topic matrix its predictions are built from (``document_distribution``, content_based.py:88-117;
        checked against the dense host flow at the citeulike-a shape with k' + k = 400,
        tests/test_gpu_bench_parity.py); ``HYPREC_HYBRID_LOWRANK=0`` or a content model without that matrix fall
        back to the reference's dense host flow (None)."""
        if os.environ.get("HYPREC_HYBRID_LOWRANK", "1") == "0":
            return None
        content = getattr(self, 'item_based_recommender', None)
        theta = getattr(content, 'document_distribution', None)
        if theta is None:                   # the content model would draw a random topic matrix (content_based.py:105-106)
            return None
        from hyprec_b200 import hybrid
        indptr, indices, values = self._host_csr(self.train_data)
        train = sp.csr_matrix((values, indices, indptr), shape=(self.n_users, self.n_items))
        w_u, w_v, _, _ = hybrid.blended_factors(train, theta, self.user_vecs, self.item_vecs)
        return w_u, w_v

    def get_evaluation_report(self):
        """abstract_recommender.py:100-131.  Scores never leave the GPU: the device returns row
        thresholds, per-row counts of rounded ones, the streamed top-200 of every user's candidate
        list with hit flags, and the rounded prediction at every train / test positive; the metric
        arithmetic is then the reference's."""
        blended = self._report_factors() if self._is_hybrid else None
        if self._is_hybrid and blended is None:
            return AbstractRecommender.get_evaluation_report(self)
        fold = self.hyperparameters['fold']
        h = self._upload_eval()
        if blended is None:
            self._upload_factors()
        else:                               # hybrid scores as one low-rank product (hyprec_b200/hybrid.py)
            h.set_factors(*blended)
            self._sync_token = None         # the device no longer holds user_vecs / item_vecs
        rank, world = self._world()
        hashed = getattr(self.evaluator, 'hashed_candidates', None)
        naive = None
        if hashed is None and fold == 0 and not isinstance(self.evaluator.test_indices, list) and \
                getattr(self.evaluator.test_indices, 'ndim', 0) == 2:
            naive = self._naive_lists(h, blended)
        if naive is not None:
            out = naive
        elif world == 1:
            key = (self.evaluator.test_indices, fold)
            if hashed is not None:
                # device-generated lists (Evaluator.use_hashed_candidates): nothing to enumerate on the host
                h.set_candidates(None, None)
                h.set_hashed_candidates(fold % hashed[0], hashed[0], hashed[1])
                self._cands_uploaded = None
            elif self._cands_uploaded is None or self._cands_uploaded[0] is not key[0] or self._cands_uploaded[1] != fold:
                h.set_hashed_candidates(0, 0)
                h.set_candidates(*self.candidate_lists(fold))        # once per fold
                self._cands_uploaded = key
            if hashed is not None:
                out = self._eval_blocks(h, 0, self.n_users, hashed[0])
            else:
                out = h.eval_topk(self.n_recommendations, None, None, 0, self.n_users,
                                  n_train_nz=self._train_nnz, n_test_nz=self._test_nnz,
                                  compact=getattr(self, '_hits_u8', False))
        else:
            out = self._eval_sharded(h, rank, world, fold)
        rmse = h.rmse()
        if out["topk_hit"].dtype != numpy.float64:      # one byte per hit on the wire; the metric arithmetic is float64
            out["topk_hit"] = out["topk_hit"].astype(numpy.float64)

        lengths = (out["topk_idx"] >= 0).sum(axis=1)
        recs = [row[:n].tolist() for row, n in zip(out["topk_idx"], lengths)]
        hits = out["topk_hit"]
        self.evaluator.recommendation_indices = recs
        self.evaluator.recs_loaded = True
        if hasattr(self.evaluator, "set_device_results"):
            self.evaluator.set_device_results(recs, [hits[u][:lengths[u]] for u in range(self.n_users)])

        # (sums over the recorded non-zeros: what test_data.sum() / train_data.sum() / sum(sum(ratings)) of the
        # reference add up, abstract_recommender.py:104-115, without three passes over dense m x n arrays)
        _, test_sum, test_likes = self._facts(self.test_data)
        _, train_sum, _ = self._facts(self.train_data)
        with numpy.errstate(divide='ignore', invalid='ignore'):
            # sum(rounded[ratings.nonzero()]) / sum(sum(ratings))  (evaluator.py:263-266)
            train_recall = numpy.float64(out["train_flags"].sum(dtype=numpy.int64)) / numpy.float64(train_sum)
            test_recall = numpy.float64(out["test_flags"].sum(dtype=numpy.int64)) / numpy.float64(test_sum)
            ratio = numpy.float64(out["ones_per_row"].sum()) / self._facts(self.ratings)[1]
        recall_at_x = metrics.recall_at_x(self.recall_cutoff, hits, test_likes)
        mrr_at_five = metrics.mrr_at(5, hits, lengths)
        ndcg_at_five = metrics.ndcg_at(5, hits, lengths)
        mrr_at_ten = metrics.mrr_at(10, hits, lengths)
        ndcg_at_ten = metrics.ndcg_at(10, hits, lengths)
        out["lengths"] = lengths
        self.last_report_details = out
        report = (test_sum, train_sum, rmse, train_recall, test_recall, recall_at_x, ratio, mrr_at_five,
                  ndcg_at_five, mrr_at_ten, ndcg_at_ten)
        if self._verbose:
            print(REPORT_FORMAT.format(*report))
        return report
