"""GPU tests of the sharded path: the exchange inside the library (csrc/comm.cuh -- peer-memory stores from
the kernels' epilogues, mailbox Gram reduction, device barriers), driven through the C ABI.

* one rank with the exchange attached == the plain handle;
* TWO ranks on ONE GPU (two handles, one host thread each, arenas attached as plain pointers with
  hyprec_comm_attach_local): the whole protocol -- partition, peer stores, mailbox, barriers, sharded RMSE, the
  shared early stop of hyprec_train -- against the single-handle run and the float64 oracle;
* two PROCESSES on two GPUs (torchrun, CUDA IPC + NVLink; skipped on a one-GPU box): tools/dist_exchange_check.py.

Reference arithmetic: /root/reference/lib/collaborative_filtering.py:79-98 (als_step), :241-259 (epoch loop).
"""
import os
import subprocess
import sys
import threading

import numpy
import pytest
import scipy.sparse as sp

from oracle import als as oals

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    from hyprec_b200 import _native
    return _native


def problem(m, n, k, nnz, seed, heavy_item=None):
    from hyprec_b200 import synth
    full = synth.ratings_csr("toy", n_users=m, n_items=n, positives=nnz, min_degree=1, seed=seed)
    train = oals.to_csr(full)
    if heavy_item is not None:          # one item almost every user rated: a long row on the item side
        dense = numpy.asarray(train.todense())
        dense[::2, heavy_item] = 1.0
        dense[1::3, heavy_item] = 1.0
        train = oals.to_csr(dense)
    rs = numpy.random.RandomState(seed + 1)
    return train, rs.rand(m, k), rs.rand(n, k)


def loaded(native, train, u0, v0, precision):
    h = native.Handle(0, precision)
    h.set_train_csr(train.shape[0], train.shape[1], train.indptr, train.indices, train.data)
    h.set_factors(u0, v0)
    return h


def bounds_for(train, world, k):
    from hyprec_b200 import distributed
    ub = distributed.partition_rows(train.indptr, world, k)
    ib = distributed.partition_rows(distributed.csc_indptr(train.indptr, train.indices, train.shape[1]), world, k)
    return ub, ib


@pytest.mark.parametrize("precision,k", [(0, 24), (0, 200), (1, 64), (1, 32)])
def test_one_rank_exchange_equals_plain_handle(native, precision, k):
    train, u0, v0 = problem(120, 420, k, 5000, seed=5)
    plain = loaded(native, train, u0, v0, precision)
    shard = loaded(native, train, u0, v0, precision)
    blob = shard.comm_export(0, 1)
    shard.comm_attach([blob], [0, train.shape[0]], [0, train.shape[1]])
    for _ in range(2):
        for side in (0, 1):
            plain.als_half_step(side, 0.01)
            shard.als_half_step(side, 0.01)
        r0, r1 = plain.rmse(), shard.rmse()
        assert abs(r0 - r1) <= (1e-12 if precision == 0 else 2e-6) * max(1.0, abs(r0))
    pu, pv = plain.get_factors()
    su, sv = shard.get_factors()
    tol = 1e-10 if precision == 0 else 5e-5
    assert numpy.linalg.norm(pu - su) <= tol * numpy.linalg.norm(pu)
    assert numpy.linalg.norm(pv - sv) <= tol * numpy.linalg.norm(pv)
    stats = shard.comm_stats()
    assert stats["attached"] and stats["barriers"] > 0
    shard.comm_detach()
    shard.als_half_step(0, 0.01)            # back to a plain handle
    plain.als_half_step(0, 0.01)
    assert numpy.linalg.norm(plain.get_factors()[0] - shard.get_factors()[0]) <= tol * numpy.linalg.norm(pu)
    plain.close()
    shard.close()


def run_ranks(native, train, u0, v0, precision, world, k, n_iter, bounds=None, prior=None):
    """`world` handles on GPU 0, one host thread each, through hyprec_train; returns per-rank (history, U, V)."""
    ub, ib = bounds if bounds is not None else bounds_for(train, world, k)
    handles = [loaded(native, train, u0, v0, precision) for _ in range(world)]
    for r, h in enumerate(handles):
        if prior is not None:
            h.set_item_prior(prior)
        # One epoch as a single-rank exchange first: it sizes every scratch buffer.  The ranks of this test share ONE
        # device, where a device allocation in the middle of a step (an implicit synchronisation point of the CUDA
        # runtime) would put this rank's next kernels behind the other rank's barrier kernel, which waits for them.
        # Ranks on their own GPUs (the real case) never meet this.
        h.comm_attach([h.comm_export(0, 1)], [0, train.shape[0]], [0, train.shape[1]])
        h.als_half_step(0, 0.01)
        h.als_half_step(1, 0.01)
        h.rmse()
        h.comm_detach()
        h.set_factors(u0, v0)
        h.comm_export(r, world)
    arenas = [h.comm_arena()[0] for h in handles]
    for h in handles:
        h.comm_attach_local(arenas, [0] * world, ub, ib)
    results, errors = [None] * world, []

    def work(r):
        try:
            history = []
            handles[r].train(n_iter, 0.01, numpy.inf, lambda e, rmse, s: history.append(rmse))
            results[r] = history
        except Exception as exc:      # surfaced below: a dead rank makes the others time out
            errors.append((r, exc))

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    assert not errors, errors
    out = []
    for r, h in enumerate(handles):
        u, v = h.get_factors()
        out.append((results[r], u, v, h.comm_stats()))
    for h in handles:
        h.comm_detach()
    for h in handles:
        h.close()
    return out, (ub, ib)


@pytest.mark.parametrize("precision,k,world", [(0, 24, 2), (0, 200, 2), (1, 64, 2), (1, 64, 3), (1, 32, 2)])
def test_ranks_on_one_gpu_reproduce_the_single_handle_run(native, precision, k, world):
    train, u0, v0 = problem(150, 500, k, 7000, seed=9, heavy_item=3)
    n_iter = 3
    single = loaded(native, train, u0, v0, precision)
    want_hist = []
    single.train(n_iter, 0.01, numpy.inf, lambda e, rmse, s: want_hist.append(rmse))
    wu, wv = single.get_factors()
    single.close()
    ranks, (ub, ib) = run_ranks(native, train, u0, v0, precision, world, k, n_iter)
    tol = 1e-9 if precision == 0 else 1e-4
    for history, u, v, stats in ranks:
        assert history == ranks[0][0]                       # bit-identical RMSE on every rank: the same early stop
        assert len(history) == len(want_hist)
        assert numpy.allclose(history, want_hist, rtol=1e-9 if precision == 0 else 5e-6, atol=0)
        assert numpy.array_equal(u, ranks[0][1]) and numpy.array_equal(v, ranks[0][2])     # identical replicas
        assert numpy.linalg.norm(u - wu) <= tol * numpy.linalg.norm(wu)
        assert numpy.linalg.norm(v - wv) <= tol * numpy.linalg.norm(wv)
        assert stats["bytes_pushed"] > 0
    # and against the float64 oracle (same tolerance as the single-GPU parity tests)
    ou, ov = u0.copy(), v0.copy()
    oals.partial_train(ou, ov, train, 0.01, n_iter)
    otol = 1e-8 if precision == 0 else 2e-4
    assert numpy.linalg.norm(ranks[0][1] - ou) <= otol * numpy.linalg.norm(ou)
    assert numpy.linalg.norm(ranks[0][2] - ov) <= otol * numpy.linalg.norm(ov)


def test_skewed_item_shard_and_empty_shard_fp32(native):
    """float32, k = 32 (tensor-memory kernels for every bucket but no k x k tensor path), an item
    shard that is empty on one rank -- every rank must take the same RMSE branch."""
    k = 32
    train, u0, v0 = problem(100, 300, k, 4000, seed=21, heavy_item=0)
    m, n = train.shape
    single = loaded(native, train, u0, v0, 1)
    want = []
    single.train(2, 0.01, numpy.inf, lambda e, rmse, s: want.append(rmse))
    single.close()
    ranks, _ = run_ranks(native, train, u0, v0, 1, 2, k, 2, bounds=([0, 10, m], [0, n, n]))
    assert ranks[0][0] == ranks[1][0]
    assert numpy.allclose(ranks[0][0], want, rtol=5e-6, atol=0)


def test_item_prior_sharded_fp64(native):
    """update_with_items (collaborative_filtering.py:93-96) under the sharded exchange."""
    from hyprec_b200 import synth
    k = 40
    train, u0, _ = problem(90, 260, k, 3000, seed=4)
    theta = synth.dirichlet_theta(train.shape[1], k, seed=3)
    single = loaded(native, train, u0, theta.copy(), 0)
    single.set_item_prior(theta)
    single.train(2, 0.01)
    wu, wv = single.get_factors()
    single.close()
    ranks, _ = run_ranks(native, train, u0, theta.copy(), 0, 2, k, 2, prior=theta)
    assert numpy.allclose(ranks[0][1], wu, rtol=1e-8, atol=1e-11) and numpy.allclose(ranks[1][2], wv, rtol=1e-8, atol=1e-11)
    ou, ov = u0.copy(), theta.copy()
    oals.partial_train(ou, ov, train, 0.01, 2, prior=theta)
    assert numpy.allclose(ranks[0][2], ov, rtol=1e-8, atol=1e-11)


def test_device_initialised_factors_match_the_restated_generator(native):
    train, _, _ = problem(70, 90, 16, 900, seed=2)
    for precision in (0, 1):
        h = native.Handle(0, precision)
        h.set_train_csr(70, 90, train.indptr, train.indices, train.data)
        h.init_factors_uniform(16, 4)
        u, v = h.get_factors()
        assert numpy.array_equal(u, oals.uniform_factors(70, 16, 4))
        assert numpy.array_equal(v, oals.uniform_factors(90, 16, 5))
        rows = numpy.array([3, 69, 0])
        assert numpy.array_equal(h.get_factor_rows(0, rows), u[rows])
        assert numpy.array_equal(h.get_factor_rows(1, rows), v[rows])
        assert numpy.array_equal(oals.uniform_factors(90, 16, 5, rows=rows), v[rows])
        h.close()


def test_two_processes_two_gpus_ipc():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "dist_exchange_check.py")]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "exchange: nvlink-peer-memory" in proc.stdout, proc.stdout[-2000:]
