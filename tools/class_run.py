#!/usr/bin/env python
"""What `python3 runnables.py collaborative` does (/root/reference/runnables.py:111-123), through the reference-facing
classes of this package: Evaluator(ratings) -> CollaborativeFiltering(...).train() = k-fold split, per fold random
initialisation, n_iterations ALS epochs and the evaluation report, then the mean 11-tuple.  Wall clock of the whole call
and of its host / device parts, one JSON line.

    python tools/class_run.py [--workload citeulike-a] [--iterations 5] [--folds 5] [--factors 200] [--sparse]
    python tools/class_run.py --scale 0.25 --folds 1 --factors 256      # the scaled shape through the class: sparse
        ratings, naive split, device-generated candidate lists (no dense users x items array can exist there)
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
    ap.add_argument("--iterations", type=int, default=5)
    ap.add_argument("--folds", type=int, default=5)
    ap.add_argument("--factors", type=int, default=200)
    ap.add_argument("--lambda", dest="lam", type=float, default=0.01)
    ap.add_argument("--sparse", action="store_true", help="hand the ratings over as a scipy CSR matrix")
    ap.add_argument("--scale", type=float, default=0.0, help="> 0: the scaled synthetic shape (sparse, hashed candidates)")
    args = ap.parse_args()
    from hyprec_b200 import CollaborativeFiltering, Evaluator, ModelInitializer, _native, synth
    import scipy.sparse as sp
    if args.scale > 0:
        data = synth.scaled_ratings(args.scale)
        indptr, indices = data["full"]
        ratings = sp.csr_matrix((numpy.ones(len(indices)), indices, indptr), shape=(data["m"], data["n"]))
        args.workload = "scaled-%g" % args.scale
    else:
        ratings = synth.ratings_csr(args.workload)
        if not args.sparse:
            ratings = numpy.asarray(ratings.todense(), dtype=numpy.float64)
    hyper = {'n_factors': args.factors, '_lambda': args.lam}
    options = {'k_folds': args.folds, 'n_iterations': args.iterations}
    evaluator = Evaluator(ratings)
    if args.scale > 0:
        evaluator.use_hashed_candidates(5, 20170303)
    cf = CollaborativeFiltering(ModelInitializer(hyper.copy(), args.iterations), evaluator, hyper, options,
                                load_matrices=False, dump_matrices=False)
    cf.device()                                   # context creation is not part of the run
    phases = {}

    def timed(obj, name, label):
        inner = getattr(obj, name)

        def wrapper(*a, **kw):
            t0 = time.perf_counter()
            try:
                return inner(*a, **kw)
            finally:
                phases[label] = phases.get(label, 0.0) + time.perf_counter() - t0
        setattr(obj, name, wrapper)

    timed(evaluator, "get_kfold_indices", "split_indices_s")
    timed(evaluator, "naive_split", "naive_split_s")
    timed(evaluator, "get_fold", "fold_matrices_s")
    timed(cf, "partial_train", "partial_train_s")
    timed(cf, "get_evaluation_report", "evaluation_report_s")
    timed(cf, "_init_fold", "init_factors_s")
    # inside partial_train / get_evaluation_report
    timed(cf, "_upload_train", "  upload_train_s")
    timed(cf, "_upload_eval", "  upload_eval_incl_train_s")
    timed(cf, "_download_factors", "  download_factors_s")
    timed(cf, "candidate_lists", "  candidate_lists_s")
    handle = cf.device()
    for name in ("set_factors", "train", "set_candidates", "eval_topk", "rmse"):
        timed(handle, name, "  handle.%s_s" % name)
    t0 = time.perf_counter()
    report = cf.train()
    total = time.perf_counter() - t0
    h = cf.device()
    device_ms = {name: h.kernel_ms(t)[0] for name, t in
                 dict(gram=_native.T_GRAM, solve=_native.T_SOLVE, prep=_native.T_PREP, rmse=_native.T_RMSE,
                      count=_native.T_COUNT, topk=_native.T_TOPK, flags=_native.T_FLAGS).items()}
    launches = h.launch_count
    cf.close()
    print(json.dumps(dict(workload=args.workload, sparse_ratings=bool(sp.issparse(ratings)),
                          precision=os.environ.get("HYPREC_PRECISION", "fp64"), users=int(ratings.shape[0]), items=int(ratings.shape[1]),
                          folds=args.folds, iterations=args.iterations, n_factors=args.factors, train_seconds=total,
                          phases_s=phases, device_kernel_ms_total=sum(device_ms.values()), device_kernel_ms=device_ms,
                          gpu_launches=int(launches), report=[float(x) for x in report])))


if __name__ == "__main__":
    main()
