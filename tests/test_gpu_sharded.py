"""GPU tests of the sharded path: the exchange inside the library (csrc/comm.cuh -- peer-memory stores from
the kernels' epilogues, mailbox Gram reduction, device barriers), driven through the C ABI.

* one rank with the exchange attached == the plain handle;
* two processes on two GPUs over NVLink (torchrun, CUDA IPC arenas; skipped on a one-GPU box -- ranks that wait on
  one another must not share a GPU, /opt/skills/guides/B200_PROFILING.md): the whole protocol -- partition, peer
  stores, mailbox, barriers, sharded RMSE, the shared early stop of hyprec_train, an empty item shard, the item
  prior -- against the float64 oracle, plus the torch.distributed fallback and the reference-facing class on the
  golden runs (tools/dist_exchange_check.py; its log on two B200s: profiles/r02_dist_exchange_check_2gpu.log).
  The protocol's kernels also run on the CPU emulator with three ranks as three host threads
  (tests/test_emu_kernels.py::test_exchange_protocol_three_ranks).

Reference arithmetic: /root/reference/lib/collaborative_filtering.py:79-98 (als_step), :241-259 (epoch loop).
"""
import os
import subprocess
import sys

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


def _torchrun(extra, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "dist_exchange_check.py")] + extra
    return subprocess.run(cmd, capture_output=True, text=True, timeout=280, cwd=ROOT)


def test_two_processes_two_gpus_ipc():
    """The same over NVLink between two GPUs (NCCL group; also checks the torch.distributed fallback path)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    proc = _torchrun([], 29533)
    assert proc.returncode == 0, proc.stdout[-4000:] + proc.stderr[-3000:]
    assert "exchange: nvlink-peer-memory" in proc.stdout and "MISMATCH" not in proc.stdout, proc.stdout[-3000:]
