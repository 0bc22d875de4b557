"""
Mirror of the reference's hyper-parameter sweep (/root/reference/lib/grid_search.py:13-131): same
constructor, ``get_all_combinations`` / ``get_key`` / ``train`` / ``dump_csv`` / ``get_all_errors``,
same result rows, same best-configuration rule (first strictly smaller ``1 - test_recall`` wins).

What differs is how a device-backed recommender (``hyprec_b200.CollaborativeFiltering``) is driven
(BASELINE.json configs[4]):

* no dense matrices: the reference pulls ``get_predictions()`` / ``rounded_predictions()`` (two m x n
  float64 arrays per configuration) only to count rounded ones at the positives of the full ratings and of
  the naive test split (grid_search.py:72-75).  The device already returns the rounded prediction at every
  train / test positive, and train + test positives = all positives, so both recalls are sums of
  those flags;
* models are independent, so ``workers`` of them are in flight at once, each on its own C-ABI handle
  (own streams, own device buffers): the device work of one model overlaps the host work (split,
  initial factors, metric arithmetic) of the others.  Everything that consumes the global
  ``numpy.random`` stream runs under one lock and in the reference's order per model: with the
  evaluator's default fixed seed every model is initialised from the same post-split generator state
  as in the reference (seed 42 -> naive split -> user matrix -> item matrix, App. A Q4), so the sweep
  returns the same rows whatever ``workers`` is;
* under ``torchrun`` (one process per GPU) the configurations are dealt round-robin to the ranks, each
  model trains on one GPU, and the rows are all-gathered: every rank returns the full table.

Any other recommender (no device handle, hybrid, checkpoint files in play, k-fold splits) goes through
the reference's sequential flow unchanged.
"""
import copy
import csv
import itertools as it
import os
import threading

import numpy

from hyprec_b200.evaluator import Evaluator

HEADER = ['n_factors', '_lambda', 'rmse', 'train_recall', 'test_recall', 'recall_at_200', 'ratio',
          'mrr @ 5', 'ndcg @ 5', 'mrr @ 10', 'ndcg @ 10']            # grid_search.py:61-62

_RNG_LOCK = threading.Lock()


def _cell_keys(indptr, indices, n_items):
    rows = numpy.repeat(numpy.arange(len(indptr) - 1, dtype=numpy.int64), numpy.diff(indptr))
    return rows * n_items + numpy.asarray(indices, dtype=numpy.int64)


class CellSet(object):
    """The non-zero cells of a users x items matrix as sorted ``row * n_items + col`` keys, with the denominator
    ``sum(sum(matrix))`` of evaluator.py:266.  Built once per sweep for the matrices every model is judged on --
    from the evaluator's CSR view of the matrix when it has one (``triplet``: no scan of the dense array)."""

    def __init__(self, matrix, triplet=None):
        self._lookup = None                       # ((pattern arrays), positions) of the latest flags_at
        if triplet is not None:
            indptr, indices, values = triplet
            self.n_items = matrix.shape[1]
            self.keys = _cell_keys(indptr, indices, self.n_items)
            # sum(sum(matrix)): column sums accumulated row by row, then added up column by column, in float64 --
            # bincount adds a column's entries in CSR (= row) order, and the zeros of the dense rows add nothing
            column_sums = numpy.bincount(numpy.asarray(indices, dtype=numpy.int64), weights=values, minlength=self.n_items)
            self.total = sum(column_sums) if len(indices) else sum(numpy.zeros(self.n_items))
            return
        matrix = numpy.asarray(matrix)
        self.n_items = matrix.shape[1]
        rows, cols = numpy.nonzero(matrix)
        self.keys = rows.astype(numpy.int64) * self.n_items + cols.astype(numpy.int64)
        self.total = sum(sum(matrix))


def flags_at(cells, train_pattern, train_flags, test_pattern, test_flags):
    """Rounded predictions (0/1) at the cells of ``cells`` (a CellSet or a matrix), looked up among the cells the
    device reported (the train and the test positives).  Returns None when some cell was not reported.
    Where a CellSet's cells sit among the reported ones depends on the two patterns only, which every model of a
    sweep shares: the positions are computed once per CellSet and pattern pair."""
    if not isinstance(cells, CellSet):
        cells = CellSet(cells)
    flags = numpy.concatenate([numpy.asarray(train_flags, dtype=numpy.uint8),
                               numpy.asarray(test_flags, dtype=numpy.uint8)])
    known = cells._lookup
    arrays = (train_pattern[0], train_pattern[1], test_pattern[0], test_pattern[1])
    if known is not None and all(a is b for a, b in zip(known[0], arrays)):
        return None if known[1] is None else flags[known[1]]
    keys = numpy.concatenate([_cell_keys(train_pattern[0], train_pattern[1], cells.n_items),
                              _cell_keys(test_pattern[0], test_pattern[1], cells.n_items)])
    order = numpy.argsort(keys, kind='stable')
    keys = keys[order]
    wanted = cells.keys
    if len(wanted) == 0:
        return numpy.zeros(0, dtype=numpy.uint8)
    pos = numpy.searchsorted(keys, wanted)
    if len(keys) == 0 or pos.max() >= len(keys) or not numpy.array_equal(keys[pos], wanted):
        cells._lookup = (arrays, None)
        return None
    where = order[pos]
    cells._lookup = (arrays, where)
    return flags[where]


def recall_from_flags(cells, flags):
    """evaluator.py:263-266 -- ``sum(rounded[ratings.nonzero()]) / sum(sum(ratings))``."""
    if not isinstance(cells, CellSet):
        cells = CellSet(cells)
    with numpy.errstate(divide='ignore', invalid='ignore'):
        return numpy.float64(int(flags.sum(dtype=numpy.int64))) / cells.total


class GridSearch(object):
    def __init__(self, recommender, hyperparameters, verbose=True, report_name='grid_search_results', workers=None):
        self.recommender = recommender
        self.hyperparameters = hyperparameters
        self._verbose = verbose
        self.evaluator = Evaluator(recommender.get_ratings())
        self.all_errors = dict()
        self.results_file_name = report_name + '.csv'
        if workers is None:
            workers = int(os.environ.get("HYPREC_GRID_WORKERS", "3"))
        self.workers = max(1, int(workers))

    # ------------------------------------------------------------------ combinatorics (:33-46, :90-108)
    def get_all_combinations(self):
        names = sorted(self.hyperparameters)
        return [dict(zip(names, prod)) for prod in it.product(*(self.hyperparameters[name] for name in names))]

    def get_key(self, config):
        return ','.join('%s:%s' % (key, config[key]) for key in sorted(config))

    def get_all_errors(self):
        return self.all_errors

    def dump_csv(self, all_results):
        """grid_search.py:110-122: appended to ``matrices/<report_name>.csv`` next to the package."""
        base_dir = os.path.dirname(os.path.realpath(__file__))
        path = os.path.join(os.path.dirname(base_dir), 'matrices/%s' % self.results_file_name)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "a") as f:
            csv.writer(f).writerows(all_results)
        if self._verbose:
            print("dumped to %s" % path)

    # ------------------------------------------------------------------ the sweep (:48-88)
    def _device_backed(self):
        rec = self.recommender
        return (hasattr(rec, 'device') and hasattr(rec, '_init_fold') and not getattr(rec, '_is_hybrid', False)
                and rec.splitting_method == 'naive' and not rec._load_matrices and not rec._dump_matrices)

    def train(self):
        combos = self.get_all_combinations()
        if self._device_backed():
            outcomes = self._sweep_device(combos)
        else:
            outcomes = self._sweep_sequential(combos)
        best_error = numpy.inf
        best_params = dict()
        all_results = [list(HEADER)]
        # under torchrun every rank holds the whole table; the reference writes it once: rank 0 does
        rank = self.recommender._world()[0] if hasattr(self.recommender, '_world') else 0
        for hyperparameters, (report, train_recall, test_recall) in zip(combos, outcomes):
            all_results.append([hyperparameters['n_factors'], hyperparameters['_lambda']] + list(report))
            if self._verbose and rank == 0:
                print('Train error: %f, Test error: %f' % (train_recall, test_recall))
            if 1 - test_recall < best_error:
                best_params = hyperparameters
                best_error = 1 - test_recall
            key = self.get_key(hyperparameters)
            self.all_errors[key] = dict(train_recall=train_recall, test_recall=test_recall)
        if rank == 0:
            self.dump_csv(all_results)
            if self._verbose:
                print("Best config: %s" % best_params)
        return best_params, all_results

    def _sweep_sequential(self, combos):
        """The reference's loop, verbatim in behaviour (dense predictions on the host)."""
        rec = self.recommender
        train, test = rec.evaluator.naive_split(rec._split_type)
        outcomes = []
        for hyperparameters in combos:
            if self._verbose:
                print("Running config: %s" % hyperparameters)
            rec.set_hyperparameters(hyperparameters)
            rec.train()
            report = rec.get_evaluation_report()
            rounded = rec.rounded_predictions()
            test_recall = self.evaluator.calculate_recall(test, rounded)
            train_recall = self.evaluator.calculate_recall(rec.get_ratings(), rounded)
            outcomes.append((report, train_recall, test_recall))
        return outcomes

    # ------------------------------------------------------------------ device-backed sweep
    def _make_worker(self):
        """A recommender of the same configuration with its own evaluator and its own device handle."""
        rec = copy.copy(self.recommender)
        rec.evaluator = copy.copy(self.recommender.evaluator)
        rec.evaluator.recommendation_indices = [[] for _ in range(rec.evaluator.n_users)]   # not the shared list
        rec.evaluator.recs_loaded = False
        if hasattr(rec.evaluator, '_device_hits'):
            rec.evaluator._device_hits = None
        rec.hyperparameters = dict(self.recommender.hyperparameters)
        rec._matrix_facts = list(getattr(self.recommender, '_matrix_facts', []))    # entries shared, list private
        rec._handle = None
        rec._train_uploaded = rec._eval_uploaded = rec._cands_uploaded = rec._sync_token = None
        rec._train_pattern = rec._test_pattern = None
        rec.last_report_details = None
        rec.predictions = None
        rec._single_rank = True
        rec.__dict__.pop('_ahead_pool', None)          # (a k-fold run's host thread belongs to the caller's object)
        rec._cands_ready = None
        return rec

    def _prepare(self, rec, hyperparameters, shared):
        """Host half of ``recommender.train()`` for the naive split (collaborative_filtering.py:102-113,
        :166-178), under the RNG lock.  With a fixed-seed evaluator the split is the same for every model:
        it is computed once and every model starts its draws from the generator state right after it."""
        with _RNG_LOCK:
            rec.set_hyperparameters(hyperparameters)
            rec.document_distribution = None
            if rec.evaluator.random_seed is False and shared.get('split') is not None:
                train, test, state = shared['split']
                rec.evaluator.test_indices = test
                numpy.random.set_state(state)
            else:
                train, test = rec.evaluator.naive_split(rec._split_type)
                if rec.evaluator.random_seed is False:
                    shared['split'] = (train, test, numpy.random.get_state())
            rec.set_data(train, test)
            rec.hyperparameters['fold'] = 0
            found = rec._init_fold(None)
            shared['last_state'][self.get_key(hyperparameters)] = numpy.random.get_state()
        return found

    def _evaluate(self, rec, hyperparameters, shared, naive_test):
        """One model: train, report, the two recalls of grid_search.py:74-75 from the device's flags."""
        found = self._prepare(rec, hyperparameters, shared)
        report = rec._run_fold(found)
        if rec._verbose:                 # the reference asks for the report a second time (:69)
            from hyprec_b200.abstract_recommender import REPORT_FORMAT
            print(REPORT_FORMAT.format(*report))
        details = rec.last_report_details
        cells = shared['cells']          # (naive test split, full ratings) of grid_search.py:74-75
        lookups = [flags_at(c, rec._train_pattern, details["train_flags"], rec._test_pattern, details["test_flags"])
                   for c in cells]
        if lookups[0] is None or lookups[1] is None:
            rounded = rec.rounded_predictions()
            return report, self.evaluator.calculate_recall(rec.get_ratings(), rounded), \
                self.evaluator.calculate_recall(naive_test, rounded)
        return report, recall_from_flags(cells[1], lookups[1]), recall_from_flags(cells[0], lookups[0])

    def _sweep_device(self, combos):
        rec0 = self.recommender
        rank, world = rec0._world() if hasattr(rec0, '_world') else (0, 1)
        with _RNG_LOCK:
            naive_train, naive_test = rec0.evaluator.naive_split(rec0._split_type)       # :59
            view = getattr(rec0.evaluator, 'csr_view', lambda matrix: None)
            shared = dict(split=None, last_state={},
                          cells=(CellSet(naive_test, view(naive_test)), CellSet(rec0.get_ratings(), view(rec0.get_ratings()))))
            if rec0.evaluator.random_seed is False:
                shared['split'] = (naive_train, naive_test, numpy.random.get_state())
        mine = list(range(rank, len(combos), world))
        results, owner = {}, {}
        n_workers = min(self.workers, max(1, len(mine)))
        workers = [self._make_worker() for _ in range(n_workers)]
        cursor = iter(mine)
        cursor_lock = threading.Lock()
        failures = []

        def loop(worker):
            while not failures:
                with cursor_lock:
                    index = next(cursor, None)
                if index is None:
                    return
                if self._verbose:
                    print("Running config: %s" % combos[index])
                try:
                    owner[index] = worker
                    results[index] = self._evaluate(worker, combos[index], shared, naive_test)
                except BaseException as exc:      # re-raised in the caller's thread
                    failures.append(exc)
                    return

        try:
            if n_workers == 1:
                loop(workers[0])
            else:
                threads = [threading.Thread(target=loop, args=(w,)) for w in workers]
                for t in threads:
                    t.start()
                for t in threads:
                    t.join()
            if failures:
                raise failures[0]
            last_owner = owner.get(len(combos) - 1)
            self._adopt(last_owner, combos, shared)
        finally:
            for w in workers:
                w.close()
        if world > 1:
            import torch.distributed as dist
            if dist.get_backend() == "nccl":        # object collectives stage through the current CUDA device
                import torch
                torch.cuda.set_device(int(os.environ.get("HYPREC_DEVICE", os.environ.get("LOCAL_RANK", "0"))))
            gathered = [None] * world
            dist.all_gather_object(gathered, {i: tuple(results[i]) for i in results})
            for part in gathered:
                results.update(part)
            # every rank leaves train() holding the last configuration's model, as the reference's loop does: its
            # owner hands the factors and the report details round (the split is the same on every rank)
            if combos:
                src = (len(combos) - 1) % world
                payload = [(rec0.user_vecs, rec0.item_vecs, rec0.last_report_details, rec0.evaluator.recommendation_indices,
                            rec0.evaluator.recs_loaded) if rank == src else None]
                dist.broadcast_object_list(payload, src=src)
                if rank != src:
                    (rec0.user_vecs, rec0.item_vecs, rec0.last_report_details, rec0.evaluator.recommendation_indices,
                     rec0.evaluator.recs_loaded) = payload[0]
                    if shared['split'] is not None:
                        rec0.hyperparameters['fold'] = 0
                        rec0.document_distribution = None
                        rec0.set_data(shared['split'][0], shared['split'][1])
                        rec0.evaluator.test_indices = shared['split'][1]
        return [results[i] for i in range(len(combos))]

    def _adopt(self, worker, combos, shared):
        """Leave the caller's recommender (and the global generator) as the reference's loop would: holding
        the last configuration's model."""
        if not combos:
            return
        state = shared['last_state'].get(self.get_key(combos[-1]))
        if state is not None:
            numpy.random.set_state(state)
        rec = self.recommender
        rec.set_hyperparameters(combos[-1])
        if worker is None:
            return
        rec.hyperparameters['fold'] = 0
        rec.document_distribution = None
        rec.set_data(worker.train_data, worker.test_data)
        rec.evaluator.test_indices = worker.evaluator.test_indices
        rec.evaluator.recommendation_indices = worker.evaluator.recommendation_indices
        rec.evaluator.recs_loaded = worker.evaluator.recs_loaded
        rec.user_vecs, rec.item_vecs = worker.user_vecs, worker.item_vecs
        rec.last_report_details = worker.last_report_details
