"""Test infrastructure for tests/test_grid_search.py: a ``CollaborativeFiltering`` whose two device
entry points (the epochs and the evaluation report) are answered by the oracle, so that the sweep's
host logic -- generator hand-over between models, worker threads, flag look-ups, rank striding -- runs
in the CPU-only suite.  Never a product path: the product class has no CPU fallback."""
import os

import numpy

from hyprec_b200.abstract_recommender import AbstractRecommender
from hyprec_b200.collaborative_filtering import CollaborativeFiltering, _csr_triplet
from oracle import als as oals

GRID = {'_lambda': [0.01, 0.1, 1.0], 'n_factors': [4, 8]}


class OracleBacked(CollaborativeFiltering):
    def partial_train(self):
        oals.partial_train(self.user_vecs, self.item_vecs, oals.to_csr(self.train_data), self._lambda, self.n_iter)

    def get_predictions(self):
        return self.user_vecs.dot(self.item_vecs.T)

    def get_evaluation_report(self):
        report = AbstractRecommender.get_evaluation_report(self)
        rounded = self.rounded_predictions()
        details = {}
        for name, matrix in (("train", self.train_data), ("test", self.test_data)):
            indptr, indices, _ = _csr_triplet(matrix)
            rows = numpy.repeat(numpy.arange(matrix.shape[0]), numpy.diff(indptr))
            details[name + "_flags"] = rounded[rows, indices].astype(numpy.uint8)
            setattr(self, "_%s_pattern" % name, (indptr, indices))
        self.last_report_details = details
        return report

    def close(self):
        pass


def make_recommender(cls, ratings, n_iter=5, random_seed=False, **flags):
    from hyprec_b200 import Evaluator, ModelInitializer
    hyper = {'n_factors': 4, '_lambda': 0.01}
    options = {'k_folds': 1, 'n_iterations': n_iter}
    kwargs = dict(load_matrices=False, dump_matrices=False)
    kwargs.update(flags)
    return cls(ModelInitializer(hyper.copy(), n_iter), Evaluator(ratings, random_seed=random_seed), hyper, options,
               **kwargs)


def sweep_under_gloo(rank, world, port, ratings, out):
    import torch.distributed as dist
    from hyprec_b200.grid_search import GridSearch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        search = GridSearch(make_recommender(OracleBacked, ratings), GRID, verbose=False, workers=2)
        dumps = []
        search.dump_csv = lambda rows: dumps.append(len(rows))
        best, rows = search.train()
        rec = search.recommender
        out[rank] = (best, [[float(x) for x in row] for row in rows[1:]], len(dumps),
                     (numpy.asarray(rec.user_vecs).copy(), numpy.asarray(rec.item_vecs).copy()))
    finally:
        dist.destroy_process_group()
