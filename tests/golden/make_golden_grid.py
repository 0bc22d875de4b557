"""
Golden vectors of the hyper-parameter sweep: the UNMODIFIED reference ``GridSearch``
(/root/reference/lib/grid_search.py:48-88) driving the UNMODIFIED reference ``CollaborativeFiltering``
on the toy 40 x 120 ratings (one user without ratings), naive split, 5 iterations,
``_lambda`` in {0.01, 0.1, 1.0} x ``n_factors`` in {4, 8}.  Build container only:

    python tests/golden/make_golden_grid.py        ->  tests/golden/grid_naive_toy.npz

Stored: the result rows (13 numbers per configuration -- n_factors, _lambda and the 11-tuple report; the
reference's header names only 11 of them, grid_search.py:61-68), the best configuration, the
train / test recalls of ``get_all_errors()`` and the generator's position after the sweep (one
``numpy.random.random()`` draw).  ``dump_csv`` is redirected to a scratch directory by a subclass (the
reference tree is read-only); nothing of the reference is copied or modified.
"""
import os
import sys
import tempfile

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.reference_loader import load_reference  # noqa: E402
from hyprec_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
GRID = {'_lambda': [0.01, 0.1, 1.0], 'n_factors': [4, 8]}


def toy_ratings():
    toy = numpy.asarray(synth.ratings_csr("toy").todense()).astype(numpy.int64)
    toy[3, :] = 0
    return toy


def main():
    ref = load_reference()
    scratch = tempfile.mkdtemp()

    class ScratchGridSearch(ref.GridSearch):
        def dump_csv(self, all_results):
            import csv
            with open(os.path.join(scratch, self.results_file_name), "a") as f:
                csv.writer(f).writerows(all_results)

    ratings = toy_ratings()
    hyper = {'n_factors': 4, '_lambda': 0.01}
    options = {'k_folds': 1, 'n_iterations': 5}
    evaluator = ref.Evaluator(ratings)
    cf = ref.CollaborativeFiltering(ref.ModelInitializer(hyper.copy(), 5), evaluator, hyper, options,
                                    load_matrices=False, dump_matrices=False)
    search = ScratchGridSearch(cf, GRID, verbose=False)
    best, rows = search.train()
    after = numpy.random.random()
    combos = search.get_all_combinations()
    errors = search.get_all_errors()
    out = dict(ratings=ratings, n_iter=5,
               lambdas=numpy.array(GRID['_lambda']), factors=numpy.array(GRID['n_factors']),
               header=numpy.array(rows[0]), rows=numpy.array([[float(x) for x in row] for row in rows[1:]]),
               best=numpy.array([best['_lambda'], best['n_factors']], dtype=numpy.float64),
               keys=numpy.array([search.get_key(c) for c in combos]),
               train_recall=numpy.array([float(errors[search.get_key(c)]['train_recall']) for c in combos]),
               test_recall=numpy.array([float(errors[search.get_key(c)]['test_recall']) for c in combos]),
               rng_after=after, final_user_vecs=cf.user_vecs, final_item_vecs=cf.item_vecs)
    numpy.savez_compressed(os.path.join(OUT, "grid_naive_toy.npz"), **out)
    print("rows", out["rows"].shape, "best", best)
    print(numpy.round(out["rows"], 4))


if __name__ == "__main__":
    main()
