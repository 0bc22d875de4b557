"""Which partial sums of y = Q_r^T t reach the output?  citeulike-t, float32, user half-step, rows of 128..191 positives."""
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
rows = numpy.nonzero((deg >= 128) & (deg <= 191))[0]
h = _native.Handle(0, 1)
h.set_train_csr(m, n, train.indptr, train.indices, train.data)
h.set_factors(u0, v0)
h.als_half_step(_native.SIDE_USER, lam)
du, _ = h.get_factors()
h.close()
v32 = v0.astype(numpy.float32).astype(numpy.float64)
B = 0.1 * v32.T.dot(v32) + lam * numpy.eye(k)
L = numpy.linalg.cholesky(B)
Q = numpy.linalg.solve(L, v32.T).T            # Q = V L^-T
for r in rows[:6]:
    cols = train.indices[train.indptr[r]:train.indptr[r + 1]]
    qr = Q[cols]
    rv = train.data[train.indptr[r]:train.indptr[r + 1]].astype(numpy.float64)
    t = numpy.linalg.solve(numpy.eye(len(cols)) + 0.9 * qr.dot(qr.T), rv)
    y_want = qr.T.dot(t)
    y_got = L.T.dot(du[r].astype(numpy.float64))
    parts = [qr[g::4].T.dot(t[g::4]) for g in range(4)]
    coef = numpy.linalg.lstsq(numpy.stack(parts, axis=1), y_got, rcond=None)[0]
    resid = numpy.linalg.norm(numpy.stack(parts, axis=1).dot(coef) - y_got) / numpy.linalg.norm(y_want)
    print("row", r, "deg", len(cols), "err %.3f" % (numpy.linalg.norm(y_got - y_want) / numpy.linalg.norm(y_want)),
          "y_got as a combination of the four group sums: coef", numpy.round(coef, 3), "residual %.2e" % resid,
          "| per 64 columns |got|/|want|:", [round(float(numpy.linalg.norm(y_got[c:c + 64]) / numpy.linalg.norm(y_want[c:c + 64])), 3) for c in range(0, k, 64)], flush=True)
