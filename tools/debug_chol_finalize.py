"""Per-row / per-column-block error of one float32 user half-step against the oracle, rows of 129..191 positives."""
import os, sys
import numpy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hyprec_b200 import _native
from oracle import als as oals

for k in (256, 200, 224, 192, 160):
    rs = numpy.random.RandomState(k)
    m, n = 40, 600
    train = numpy.zeros((m, n))
    for u in range(m):
        d = 129 + (u * 3) % 63
        train[u, rs.permutation(n)[:d]] = 1.0
    csr = oals.to_csr(train)
    u0, v0 = rs.rand(m, k), rs.rand(n, k)
    h = _native.Handle(0, 1)
    h.set_train_csr(m, n, csr.indptr, csr.indices, csr.data)
    h.set_factors(u0, v0)
    h.als_half_step(_native.SIDE_USER, 0.01)
    du, _ = h.get_factors()
    h.close()
    want = oals.als_half_step(u0.copy(), v0, csr, 0.01)
    err = numpy.linalg.norm(du - want, axis=1) / numpy.linalg.norm(want, axis=1)
    blocks = [numpy.abs(du[:, c:c + 64] - want[:, c:c + 64]).max() for c in range(0, k, 64)]
    print("k", k, "max row err %.3e" % err.max(), "rows wrong", int((err > 1e-3).sum()), "of", m, "col-block abs err", ["%.2e" % b for b in blocks], flush=True)
