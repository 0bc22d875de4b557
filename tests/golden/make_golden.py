"""
Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
(/root/reference, imported through oracle/reference_loader.py) in the build container.
/root/reference does not exist on the GPU box, so the vectors are committed; re-run this script
only when the cases change:

    python tests/golden/make_golden.py

Cases (all with load_matrices=False, dump_matrices=False so nothing is read from / written to disk):
  als_kfold_4x30      the reference's own test fixture (tests/collaborative_tests.py:16-36):
                      ratings[u][a] = int(not (a+u) % 3), 4 users x 30 items, k=5, 3 folds, 20 iterations
  als_kfold_toy       synthetic 40 x 120, k=8, 5 folds, 6 iterations
  als_theta_toy       same data, item factors initialised from theta, update_with_items=True
                      (recommender_system.py:184-193 seam; collaborative_filtering.py:93-96)
  als_naive_toy       k_folds=1 naive split (candidate lists degenerate to ids 0/1, App. A Q7)
  top_recommendations streams with ties through util/top_recommendations.py
  metrics             the 5 x 9 fixture of tests/metrics_tests.py:22-57 with its known answers
Each ALS case stores, per fold: train, test, the initial factors (captured by a subclass hook
around partial_train -- the reference class itself is untouched), the trained factors, the
11-tuple report, the recommendation lists and the candidate lists.
"""
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.reference_loader import load_reference  # noqa: E402
from hyprec_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def recording_class(ref):
    class Recording(ref.CollaborativeFiltering):
        def __init__(self, *args, **kwargs):
            self.folds = []
            ref.CollaborativeFiltering.__init__(self, *args, **kwargs)

        def partial_train(self):
            rec = dict(train=self.train_data.copy(), test=self.test_data.copy(),
                       u0=self.user_vecs.copy(), v0=self.item_vecs.copy())
            ref.CollaborativeFiltering.partial_train(self)
            rec.update(u=self.user_vecs.copy(), v=self.item_vecs.copy())
            self.folds.append(rec)

        def get_evaluation_report(self):
            report = ref.CollaborativeFiltering.get_evaluation_report(self)
            rec = self.folds[-1]
            rec["report"] = numpy.array([float(x) for x in report])
            recs = self.evaluator.recommendation_indices
            rec["rec_len"] = numpy.array([len(r) for r in recs], dtype=numpy.int64)
            rec["rec_ids"] = numpy.array([i for r in recs for i in r], dtype=numpy.int64)
            fold = self.hyperparameters['fold']
            cands = [numpy.asarray(self.evaluator.test_indices[u * (1 + fold)]).astype(numpy.int64)
                     for u in range(self.n_users)]
            rec["cand_len"] = numpy.array([len(c) for c in cands], dtype=numpy.int64)
            rec["cand_ids"] = numpy.concatenate(cands)
            rec["rounded_ones"] = self.rounded_predictions().sum(axis=1).astype(numpy.int64)
            return report
    return Recording


def run_als_case(ref, name, ratings, k, lam, n_iter, k_folds, theta=None, update_with_items=False):
    hyper = {'n_factors': k, '_lambda': lam}
    options = {'k_folds': k_folds, 'n_iterations': n_iter}
    evaluator = ref.Evaluator(ratings)
    initializer = ref.ModelInitializer(hyper.copy(), n_iter)
    cls = recording_class(ref)
    cf = cls(initializer, evaluator, hyper, options, load_matrices=False, dump_matrices=False,
             update_with_items=update_with_items)
    mean_report = cf.train(None if theta is None else theta.copy())
    out = dict(ratings=ratings, k=k, lam=lam, n_iter=n_iter, k_folds=k_folds,
               update_with_items=int(update_with_items), mean_report=numpy.asarray(mean_report, dtype=numpy.float64),
               n_folds_run=len(cf.folds))
    if theta is not None:
        out["theta"] = theta
    if k_folds > 1:
        flat = cf.fold_test_indices
        out["kfold_len"] = numpy.array([len(x) for x in flat], dtype=numpy.int64)
        out["kfold_ids"] = numpy.concatenate([numpy.asarray(x).astype(numpy.int64) for x in flat])
    for f, rec in enumerate(cf.folds):
        for key, val in rec.items():
            out["f%d_%s" % (f, key)] = val
    numpy.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "folds", len(cf.folds), "report", numpy.round(out["mean_report"], 4))


def run_top_recommendations(ref):
    rs = numpy.random.RandomState(3)
    out = {}
    cases = [(2, [1, 1, 2, 1]), (3, [0, 0, 0, 0, 0, 0]), (3, [1, 2, 2, 2, 3]), (3, [3, 2, 2, 2]),
             (4, [5, 1, 5, 1, 5, 1, 5, 5]), (5, [2, 2]), (1, [7, 7, 8, 8, 7])]
    for _ in range(40):
        length = int(rs.randint(1, 60))
        cases.append((int(rs.randint(1, 20)), list(rs.randint(0, 6, size=length) / 2.0)))
    for c, (cap, stream) in enumerate(cases):
        top = ref.TopRecommendations(cap)
        for idx, val in enumerate(stream):
            top.insert(idx, float(val))
        out["c%d_cap" % c] = cap
        out["c%d_stream" % c] = numpy.asarray(stream, dtype=numpy.float64)
        out["c%d_result" % c] = numpy.asarray(list(reversed(top.get_indices())), dtype=numpy.int64)
    out["n_cases"] = len(cases)
    numpy.savez_compressed(os.path.join(OUT, "top_recommendations.npz"), **out)
    print("top_recommendations", len(cases), "cases")


def run_metrics(ref):
    # fixture of /root/reference/tests/metrics_tests.py:22-52 (list-of-lists: the ragged numpy.array
    # in the original is an error on numpy >= 1.24)
    ratings = numpy.array([[1, 1, 0, 0, 1, 0, 1, 0, 0], [0, 0, 1, 1, 0, 0, 0, 1, 0], [1, 1, 0, 1, 0, 0, 1, 0, 1],
                           [1, 0, 0, 0, 1, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0, 0, 0, 1]])
    expected = numpy.array([[0, 1, 0, 0, 0, 0, 0, 0, 0], [0, 0, 1, 0, 0, 0, 0, 0, 0], [0, 0, 0, 1, 0, 0, 0, 0, 1],
                            [1, 0, 0, 0, 0, 0, 0, 0, 0], [0, 1, 0, 0, 0, 0, 0, 0, 0]])
    recs = [[1], [3, 2], [4, 6, 3, 0, 8], [0], [0]]
    ev = ref.Evaluator(ratings)
    ev.recs_loaded = True
    ev.recommendation_indices = recs
    out = dict(ratings=ratings, expected=expected, rec_len=numpy.array([len(r) for r in recs]),
               rec_ids=numpy.array([i for r in recs for i in r]))
    for n in (1, 4, 5):
        out["mrr_%d" % n] = numpy.float64(ev.calculate_mrr(n, None, ratings, expected))
        out["ndcg_%d" % n] = numpy.float64(ev.calculate_ndcg(n, None, ratings, expected))
    out["recall_3"] = numpy.float64(ev.recall_at_x(3, None, ratings, expected))
    numpy.savez_compressed(os.path.join(OUT, "metrics.npz"), **out)
    print("metrics", {k: float(v) for k, v in out.items() if k[0] in "mn" and "_" in k})


def main():
    ref = load_reference()
    fixture = numpy.array([[int(not bool((a + u) % 3)) for a in range(30)] for u in range(4)])
    run_als_case(ref, "als_kfold_4x30", fixture, k=5, lam=0.01, n_iter=20, k_folds=3)
    toy = numpy.asarray(synth.ratings_csr("toy").todense()).astype(numpy.int64)
    toy[3, :] = 0          # a user without any rating: zero factors, all-tied scores (App. A Q10)
    run_als_case(ref, "als_kfold_toy", toy, k=8, lam=0.01, n_iter=6, k_folds=5)
    theta = synth.dirichlet_theta(toy.shape[1], 8, seed=7)
    run_als_case(ref, "als_theta_toy", toy, k=8, lam=0.1, n_iter=4, k_folds=5, theta=theta, update_with_items=True)
    run_als_case(ref, "als_naive_toy", toy, k=8, lam=0.01, n_iter=5, k_folds=1)
    run_top_recommendations(ref)
    run_metrics(ref)


if __name__ == "__main__":
    main()
