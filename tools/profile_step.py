"""One epoch + one evaluation at the citeulike-a shape, for ncu / compute-sanitizer captures
(quick setup: random 80/20 split instead of the reference's k-fold machinery).

    python tools/profile_step.py [--precision fp64|fp32] [--epochs 2] [--shape citeulike-a] [--k 200]
"""
import argparse
import os
import sys
import time

import numpy
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyprec_b200 import _native, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="fp64")
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--shape", default="citeulike-a")
    ap.add_argument("--k", type=int, default=200)
    ap.add_argument("--no-eval", action="store_true")
    args = ap.parse_args()
    full = synth.ratings_csr(args.shape)
    m, n = full.shape
    rs = numpy.random.RandomState(1)
    coo = full.tocoo()
    keep = rs.rand(full.nnz) < 0.8
    train = sp.csr_matrix((coo.data[keep], (coo.row[keep], coo.col[keep])), shape=full.shape)
    test = sp.csr_matrix((coo.data[~keep], (coo.row[~keep], coo.col[~keep])), shape=full.shape)
    for mat in (train, test):
        mat.sort_indices()
    h = _native.Handle(0, _native.F64 if args.precision == "fp64" else _native.F32)
    h.set_train_csr(m, n, train.indptr, train.indices, train.data)
    h.set_eval_csr((full.indptr, full.indices, full.data), (test.indptr, test.indices, test.data))
    h.set_factors(rs.rand(m, args.k), rs.rand(n, args.k))
    n_cand = n // 5
    cands = numpy.stack([rs.permutation(n)[:n_cand] for _ in range(m)]).astype(numpy.int32)
    cp = numpy.arange(m + 1, dtype=numpy.int64) * n_cand
    h.set_candidates(cp, cands.ravel())          # per fold, like bench.py
    for e in range(args.epochs):
        t0 = time.perf_counter()
        h.als_half_step(_native.SIDE_USER, 0.01)
        h.als_half_step(_native.SIDE_ITEM, 0.01)
        rmse = h.rmse()
        t1 = time.perf_counter()
        if not args.no_eval:
            h.eval_topk(200, None, None, n_train_nz=train.nnz, n_test_nz=test.nnz)
        t2 = time.perf_counter()
        print("epoch %d rmse %.6f epoch %.2f ms eval %.2f ms" % (e, rmse, 1e3 * (t1 - t0), 1e3 * (t2 - t1)))
    if not args.no_eval:
        print("eval stats", h.eval_stats())
    names = ["gram", "solve", "rmse", "sweep", "topk", "flags", "count", "prep", "comm"]
    for i, name in enumerate(names):
        total, calls, last = h.kernel_ms(i)
        print("%-6s total %.3f ms in %d regions" % (name, total, calls))
    h.close()


if __name__ == "__main__":
    main()
