"""Validate the tcgen05 score sweep against float64 and time it (run under `timeout`)."""
import os
import sys
import time

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyprec_b200 import _native  # noqa: E402


def check(m, n, k, seed=0):
    rs = numpy.random.RandomState(seed)
    u = rs.rand(m, k) - 0.3
    v = rs.rand(n, k) - 0.4
    h = _native.Handle(0, _native.F64)
    ip = numpy.arange(m + 1, dtype=numpy.int64)
    h.set_train_csr(m, n, ip, (numpy.arange(m) % n).astype(numpy.int32), None)
    h.set_factors(u, v)
    for _ in range(3):
        t0 = time.perf_counter()
        s = h.score_dense_f32()
        t1 = time.perf_counter()
    exact = u.dot(v.T)
    scale = numpy.linalg.norm(u, axis=1)[:, None] * numpy.linalg.norm(v, axis=1)[None, :]
    ratio = numpy.abs(s - exact) / scale
    total, calls, last = h.kernel_ms(_native.T_SWEEP)
    print("m=%d n=%d k=%d  max|err|/(|u||v|) = %.3e  mean %.3e  max abs err %.3e  sweep %.3f ms (wall %.1f ms)" %
          (m, n, k, ratio.max(), ratio.mean(), numpy.abs(s - exact).max(), last, 1e3 * (t1 - t0)))
    h.close()
    return ratio.max()


if __name__ == "__main__":
    worst = 0.0
    for shape in ((128, 128, 32), (130, 257, 200), (300, 900, 48), (1000, 3000, 200), (5551, 16980, 200)):
        worst = max(worst, check(*shape))
    print("worst ratio %.3e" % worst)
