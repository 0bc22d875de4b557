"""citeulike-t shape, float32, one user half-step: row errors against the oracle by degree range."""
import os, sys
import numpy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from hyprec_b200 import _native
from oracle import als as oals

wl = bench.build_workload("citeulike-t")
m, n, k, lam = wl["m"], wl["n"], wl["k"], wl["lam"]
train = wl["train_csr"]
u0, v0 = wl["u0"], wl["v0"]
deg = numpy.diff(train.indptr)
rows = numpy.concatenate([numpy.nonzero((deg >= a) & (deg <= b))[0][:12] for a, b in ((1, 31), (32, 63), (64, 127), (128, 191), (192, 100000))])
want = oals.als_half_step(u0.copy(), v0, train, lam, rows=rows)
for rep in range(2):
    h = _native.Handle(0, 1)
    h.set_train_csr(m, n, train.indptr, train.indices, train.data)
    h.set_factors(u0, v0)
    h.als_half_step(_native.SIDE_USER, lam)
    du, _ = h.get_factors()
    h.close()
    err = numpy.linalg.norm(du[rows] - want[rows], axis=1) / numpy.linalg.norm(want[rows], axis=1)
    for a, b in ((1, 31), (32, 63), (64, 127), (128, 191), (192, 100000)):
        sel = (deg[rows] >= a) & (deg[rows] <= b)
        print(os.environ.get("TAG", ""), "rep", rep, "deg %d..%d rows %d max err %.3e wrong %d" % (a, b, sel.sum(), err[sel].max() if sel.any() else 0, int((err[sel] > 1e-3).sum())),
              "degs of wrong", deg[rows][sel][err[sel] > 1e-3][:8], flush=True)
