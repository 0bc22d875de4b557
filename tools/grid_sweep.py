#!/usr/bin/env python
"""BASELINE.json configs[4]: the hyper-parameter sweep lambda in {0.01, 0.1, 1, 10} x n_factors in {50, 100, 200}
on the citeulike-a shape with the naive split (as GridSearch.train does, /root/reference/lib/grid_search.py:59),
5 iterations per model.  Times hyprec_b200.GridSearch with --workers models in flight (and, with --also-single, with
one worker, checking that both return the same rows).  One JSON line.

    python tools/grid_sweep.py [--workload citeulike-a|small] [--workers 3] [--iterations 5] [--also-single]
"""
import argparse
import json
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="citeulike-a")
    ap.add_argument("--workers", type=int, default=3)
    ap.add_argument("--iterations", type=int, default=5)
    ap.add_argument("--also-single", action="store_true")
    ap.add_argument("--lambdas", default="0.01,0.1,1,10")
    ap.add_argument("--factors", default="50,100,200")
    args = ap.parse_args()
    from hyprec_b200 import CollaborativeFiltering, Evaluator, GridSearch, ModelInitializer, synth
    grid = {'_lambda': [float(x) for x in args.lambdas.split(",")], 'n_factors': [int(x) for x in args.factors.split(",")]}
    ratings = numpy.asarray(synth.ratings_csr(args.workload).todense(), dtype=numpy.float64)
    hyper = {'n_factors': grid['n_factors'][0], '_lambda': grid['_lambda'][0]}
    options = {'k_folds': 1, 'n_iterations': args.iterations}
    timings, tables = {}, {}
    t_setup = time.perf_counter()
    runs = ([1] if args.also_single else []) + [args.workers]
    for workers in runs:
        rec = CollaborativeFiltering(ModelInitializer(hyper.copy(), args.iterations), Evaluator(ratings), hyper, options,
                                     load_matrices=False, dump_matrices=False)
        search = GridSearch(rec, grid, verbose=False, workers=workers)
        search.dump_csv = lambda rows: None
        t0 = time.perf_counter()
        best, rows = search.train()
        timings[workers] = time.perf_counter() - t0
        tables[workers] = (best, numpy.array([[float(x) for x in row] for row in rows[1:]]))
        rec.close()
    same = None
    if args.also_single:
        same = bool(tables[1][0] == tables[args.workers][0] and
                    numpy.array_equal(tables[1][1], tables[args.workers][1], equal_nan=True))
    best, table = tables[args.workers]
    n_models = len(table)
    print(json.dumps(dict(workload=args.workload, users=int(ratings.shape[0]), items=int(ratings.shape[1]), models=n_models,
                          iterations=args.iterations, grid=grid, seconds_1_worker=timings.get(1),
                          seconds_n_workers=timings[args.workers], workers=args.workers,
                          seconds_per_model=timings[args.workers] / n_models, identical_rows=same,
                          best=best, recall_at_200=[float(x) for x in table[:, 7]], rmse=[float(x) for x in table[:, 4]])))


if __name__ == "__main__":
    main()
