#!/usr/bin/env python
"""Host-side cost of the reference-facing call without a GPU: CollaborativeFiltering.train() (k-fold split, per fold
initial draws, uploads, report arithmetic) with a NULL stand-in for the C-ABI handle -- every device call returns at
once with right-shaped zeros -- so what is timed is exactly the Python / numpy / host-C++ part that tools/class_run.py
shows around the ~0.1 s of kernels.  Not a parity tool and not a bench: it tells where the host seconds go.

    python tools/host_profile.py [--workload citeulike-a] [--dense] [--profile]
"""
import argparse
import cProfile
import os
import pstats
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class NullHandle(object):
    launch_count = 0

    def __init__(self, device=0, precision=0):
        self.m = self.n = self.k = 0
        self.precision = precision

    def set_train_csr(self, m, n, *rest):
        self.m, self.n = m, n

    def set_factors(self, u, v):
        self.k = u.shape[1]
        self._u, self._v = u, v

    def get_factors(self, u=None, v=None):
        if u is None:
            u = numpy.array(self._u)
        if v is None:
            v = numpy.array(self._v)
        return u, v

    def train(self, n_iter, lam, start_error=None, on_epoch=None):
        return n_iter, 0.5

    def rmse(self):
        return 0.5

    def eval_topk(self, capacity, cp=None, ci=None, user_lo=0, user_hi=None, want_counts=True, want_topk=True,
                  n_train_nz=None, n_test_nz=None, out=None, **kw):
        nu = (self.m if user_hi is None else user_hi) - user_lo
        return dict(thresholds=numpy.zeros(nu), ones_per_row=numpy.zeros(nu, numpy.int64),
                    topk_idx=numpy.zeros((nu, capacity), numpy.int32), topk_val=numpy.zeros((nu, capacity)),
                    topk_hit=numpy.zeros((nu, capacity)),
                    train_flags=numpy.zeros(n_train_nz or 0, numpy.uint8), test_flags=numpy.zeros(n_test_nz or 0, numpy.uint8))

    def kernel_ms(self, which):
        return 0.0, 0, 0.0

    def __getattr__(self, name):
        return lambda *args, **kwargs: None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="citeulike-a")
    ap.add_argument("--dense", action="store_true")
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--folds", type=int, default=5)
    ap.add_argument("--factors", type=int, default=200)
    args = ap.parse_args()
    from hyprec_b200 import _native
    _native.Handle = NullHandle
    from hyprec_b200 import CollaborativeFiltering, Evaluator, ModelInitializer, synth
    ratings = synth.ratings_csr(args.workload)
    if args.dense:
        ratings = numpy.asarray(ratings.todense(), dtype=numpy.float64)
    hyper = {'n_factors': args.factors, '_lambda': 0.01}
    options = {'k_folds': args.folds, 'n_iterations': 5}
    evaluator = Evaluator(ratings)
    cf = CollaborativeFiltering(ModelInitializer(hyper.copy(), 5), evaluator, hyper, options,
                                load_matrices=False, dump_matrices=False)
    t0 = time.perf_counter()
    if args.profile:
        prof = cProfile.Profile()
        prof.runcall(cf.train)
        pstats.Stats(prof).sort_stats("cumtime").print_stats(45)
    else:
        cf.train()
    print("host seconds of train(): %.3f" % (time.perf_counter() - t0))


if __name__ == "__main__":
    main()
