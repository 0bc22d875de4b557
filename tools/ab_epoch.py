#!/usr/bin/env python
"""A/B of library knobs on one epoch of a citeulike-shaped problem, all variants in ONE process: for every variant the
environment is set, a fresh handle is created (the knobs are read at hyprec_create), 3 warm-up epochs run, then
`--epochs` epochs are timed between two CUDA events on the library's stream (L2 flushed before each), and the trained
factors are compared with the first variant's (same arithmetic, so they must agree to rounding).  One JSON line per
variant.

    python tools/ab_epoch.py [--shape citeulike-a] [--k 200] [--epochs 20] [--precision fp64] \
        "name:ENV=value,ENV2=value" ...           (values may contain ';' where the knob wants ',')
"""
import argparse
import json
import os
import sys

import numpy
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyprec_b200 import _native, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="citeulike-a")
    ap.add_argument("--k", type=int, default=200)
    ap.add_argument("--epochs", type=int, default=20)
    ap.add_argument("--precision", default="fp64")
    ap.add_argument("variants", nargs="*")
    args = ap.parse_args()
    full = synth.ratings_csr(args.shape)
    m, n = full.shape
    rs = numpy.random.RandomState(1)
    coo = full.tocoo()
    keep = rs.rand(full.nnz) < 0.8
    train = sp.csr_matrix((coo.data[keep], (coo.row[keep], coo.col[keep])), shape=full.shape)
    train.sort_indices()
    u0, v0 = rs.rand(m, args.k), rs.rand(n, args.k)
    names = ["gram", "solve", "rmse", "sweep", "topk", "flags", "count", "prep", "comm"]
    first = None
    for spec in ["default:"] + list(args.variants):
        name, _, envs = spec.partition(":")
        pairs = [e.split("=", 1) for e in envs.split(",") if e]
        saved = {key: os.environ.get(key) for key, _ in pairs}
        for key, value in pairs:
            os.environ[key] = value.replace(";", ",")
        try:
            h = _native.Handle(0, _native.F64 if args.precision == "fp64" else _native.F32)
            h.set_train_csr(m, n, train.indptr, train.indices, train.data)
            h.set_factors(u0, v0)
            for _ in range(3):
                h.als_half_step(_native.SIDE_USER, 0.01)
                h.als_half_step(_native.SIDE_ITEM, 0.01)
                h.rmse()
            u1, v1 = h.get_factors()
            h.set_factors(u0, v0)
            h.reset_kernel_ms()
            times = []
            for _ in range(args.epochs):
                h.flush_l2()
                h.mark(0)
                h.als_half_step(_native.SIDE_USER, 0.01)
                h.als_half_step(_native.SIDE_ITEM, 0.01)
                rmse = h.rmse()
                h.mark(1)
                times.append(h.elapsed_ms(0, 1))
            parts = {nm: round(h.kernel_ms(i)[0] / args.epochs, 4) for i, nm in enumerate(names) if h.kernel_ms(i)[1]}
            if first is None:
                first = (u1, v1)
            diff = max(float(numpy.abs(u1 - first[0]).max()), float(numpy.abs(v1 - first[1]).max()))
            print(json.dumps(dict(variant=name, env=dict(pairs), epoch_ms=round(float(numpy.mean(times)), 4),
                                  epoch_ms_min=round(float(numpy.min(times)), 4), rmse=rmse, kernel_ms=parts,
                                  max_abs_diff_vs_default=diff)), flush=True)
            h.close()
        except Exception as exc:
            print(json.dumps(dict(variant=name, env=dict(pairs), error=repr(exc))), flush=True)
        finally:
            for key, value in saved.items():
                if value is None:
                    os.environ.pop(key, None)
                else:
                    os.environ[key] = value


if __name__ == "__main__":
    main()
