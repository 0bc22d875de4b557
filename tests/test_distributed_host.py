"""N > 1 host logic on CPU: cost-balanced row partition, in-place shard exchange, sharded epochs
with partial-sum RMSE and the shared early-stop decision -- world_size 2 over gloo.  The engine is
an oracle-backed stand-in for the C-ABI handle (tests may use the oracle; the product engine is
hyprec_b200.distributed.NativeEngine)."""
import os
import socket

import numpy
import pytest
import scipy.sparse as sp

from hyprec_b200 import distributed as hd
from oracle import als as oals


def test_partition_is_contiguous_and_balanced():
    rs = numpy.random.RandomState(0)
    deg = rs.randint(0, 50, size=1000)
    indptr = numpy.concatenate(([0], numpy.cumsum(deg)))
    for world in (1, 2, 3, 8):
        b = hd.partition_rows(indptr, world, 16)
        assert b[0] == 0 and b[-1] == 1000 and all(b[i] <= b[i + 1] for i in range(world))
        cost = hd.row_costs(indptr, 16)
        shares = [cost[b[i]:b[i + 1]].sum() for i in range(world)]
        assert max(shares) <= 1.05 * (sum(shares) / world) + cost.max()
    assert hd.partition_rows(numpy.array([0, 3]), 4, 8)[-1] == 1      # fewer rows than ranks
    csr = sp.random(30, 40, density=0.2, format="csr", random_state=1)
    assert numpy.array_equal(hd.csc_indptr(csr.indptr, csr.indices, 40), csr.T.tocsr().indptr)


class OracleEngine(object):
    """Same interface as NativeEngine, arithmetic from oracle.als, tensors share memory with numpy."""

    def __init__(self, train, u0, v0):
        import torch
        self.csr = oals.to_csr(train)
        self.csc = oals.to_csr(train.T)
        self.u, self.v = u0.copy(), v0.copy()
        self.tu, self.tv = torch.from_numpy(self.u), torch.from_numpy(self.v)

    def set_shard(self, user_range, item_range):
        self.user_range, self.item_range = user_range, item_range

    def factor_tensors(self):
        return self.tu, self.tv

    def half_step(self, side, lam):
        if side == 0:
            oals.als_half_step(self.u, self.v, self.csr, lam, rows=numpy.arange(*self.user_range))
        else:
            oals.als_half_step(self.v, self.u, self.csc, lam, rows=numpy.arange(*self.item_range))

    def after_exchange(self):
        pass

    def rmse_terms(self):
        lo, hi = self.user_range
        sub = self.csr[lo:hi].tocoo()
        dots = numpy.einsum("ij,ij->i", self.u[lo + sub.row], self.v[sub.col])
        frob = numpy.sum(self.u.T.dot(self.u) * self.v.T.dot(self.v))
        return frob, float(numpy.dot(sub.data, dots)), float(numpy.dot(sub.data, sub.data))


def _worker(rank, world, port, train, u0, v0, lam, n_iter, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        k = u0.shape[1]
        csr = oals.to_csr(train)
        ub = hd.partition_rows(csr.indptr, world, k)
        ib = hd.partition_rows(hd.csc_indptr(csr.indptr, csr.indices, train.shape[1]), world, k)
        engine = OracleEngine(train, u0, v0)
        als = hd.ShardedALS(engine, ub, ib, rank, train.shape[0], train.shape[1])
        history = als.train(n_iter, lam)
        out[rank] = (engine.u.copy(), engine.v.copy(), history)
    finally:
        dist.destroy_process_group()


def test_sharded_training_equals_single_process_gloo():
    import torch.multiprocessing as mp
    rs = numpy.random.RandomState(4)
    m, n, k, lam, n_iter = 37, 90, 6, 0.01, 4
    train = (rs.rand(m, n) < 0.15).astype(numpy.float64)
    train[5, :] = 0
    u0, v0 = rs.rand(m, k), rs.rand(n, k)
    u, v = u0.copy(), v0.copy()
    epochs, want_hist = oals.partial_train(u, v, oals.to_csr(train), lam, n_iter)

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    manager = mp.Manager()
    out = manager.dict()
    world = 2
    procs = [mp.get_context("spawn").Process(target=_worker, args=(r, world, port, train, u0, v0, lam, n_iter, out))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    for r in range(world):
        ru, rv, hist = out[r]
        assert numpy.allclose(ru, u, rtol=1e-10, atol=1e-13)       # every replica holds the full result
        assert numpy.allclose(rv, v, rtol=1e-10, atol=1e-13)
        assert len(hist) == epochs and numpy.allclose(hist, want_hist, rtol=1e-12)
