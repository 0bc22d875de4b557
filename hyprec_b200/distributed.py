"""
Multi-GPU plumbing: one process per GPU (torchrun), rows of each half-step sharded across ranks,
one exchange of the freshly solved factor rows after every half-step.

The reference has no distributed code at all; what shards is the independence of rows inside
``als_step`` (/root/reference/lib/collaborative_filtering.py:82-86 reads only ``fixed_vecs``).
Each rank keeps a full replica of U and V and solves a contiguous, cost-balanced range of users
(user half-step) and of items (item half-step).

Two exchange paths:

* ``attach_exchange`` + ``NativeShardedALS`` (the product path on GPUs): the library itself stores every
  solved / whitened row into all replicas over NVLink peer memory from the epilogue of the kernel that
  produced it, reduces the k x k Gram and the RMSE sums through a mailbox and orders the ranks with
  device-side barriers (csrc/comm.cuh).  ``torch.distributed`` only carries the CUDA IPC handles at
  set-up; nothing Python-side runs inside an epoch.
* ``ShardedALS`` over an engine with ``half_step`` / ``factor_tensors`` / ``rmse_terms``: the exchange as
  ``torch.distributed`` broadcasts (NCCL on GPUs -- the fallback when peer memory cannot be mapped -- and
  gloo in the CPU tests, where an oracle-backed engine stands in for the device).

Evaluation shards users; per-user integers are gathered on rank 0.
"""
import numpy


# --------------------------------------------------------------------------------------------------
# partitioning (pure host logic)
# --------------------------------------------------------------------------------------------------
def row_costs(indptr, k):
    """FLOP model of one row update as the kernels run it: rows with deg <= k solve a deg x deg system on
    whitened factors (deg^2 k to build it, deg^3 / 3 to factor it, 4 deg k around it), rows with more
    positives the k x k system (2 deg k^2 + k^3 / 3); a fixed term keeps near-empty rows from being free."""
    deg = numpy.diff(numpy.asarray(indptr, dtype=numpy.int64)).astype(numpy.float64)
    low = deg * deg * k + deg ** 3 / 3.0 + 4.0 * deg * k
    direct = deg * (2.0 * k * k) + (k ** 3) / 3.0
    return numpy.where(deg <= k, low, direct) + 8.0 * k * k


def partition_rows(indptr, world, k):
    """Boundaries b[0..world] of contiguous row ranges with (nearly) equal modelled cost."""
    cost = numpy.cumsum(row_costs(indptr, k))
    total = cost[-1] if len(cost) else 0.0
    n_rows = len(cost)
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(numpy.searchsorted(cost, total * r / world, side="left")))
    bounds.append(n_rows)
    for r in range(1, world + 1):          # monotone, even when rows run out
        bounds[r] = max(bounds[r], bounds[r - 1])
    return bounds


def csc_indptr(indptr, indices, n_cols):
    counts = numpy.bincount(numpy.asarray(indices[:indptr[-1]], dtype=numpy.int64), minlength=n_cols)
    out = numpy.zeros(n_cols + 1, dtype=numpy.int64)
    numpy.cumsum(counts, out=out[1:])
    return out


# --------------------------------------------------------------------------------------------------
# collectives
# --------------------------------------------------------------------------------------------------
def exchange_rows(matrix, bounds, group=None):
    """In-place all-gather of a row-sharded replica: rank r owns rows [bounds[r], bounds[r+1]) and
    broadcasts them into every other rank's copy (ranges may differ in size)."""
    import torch.distributed as dist
    works = []
    for src in range(len(bounds) - 1):
        lo, hi = bounds[src], bounds[src + 1]
        if hi > lo:
            works.append(dist.broadcast(matrix[lo:hi], src=src, group=group, async_op=True))
    for w in works:
        w.wait()


class _DeviceArray(object):
    """Expose a raw device pointer through __cuda_array_interface__ so torch can view it."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = dict(shape=shape, typestr=typestr, data=(int(ptr), False), version=3,
                                             strides=None)


class NativeEngine(object):
    """The C-ABI handle as a ShardedALS engine."""

    def __init__(self, handle, device_index):
        import torch
        self.h = handle
        self.torch = torch
        self.device = torch.device("cuda", device_index)
        self._views = None

    def set_shard(self, user_range, item_range):
        self.h.set_shard(user_range[0], user_range[1], item_range[0], item_range[1])

    def factor_tensors(self):
        if self._views is None:
            u_ptr, v_ptr, elem = self.h.device_factors()
            typestr = "<f8" if elem == 8 else "<f4"
            u = self.torch.as_tensor(_DeviceArray(u_ptr, (self.h.m, self.h.k), typestr), device=self.device)
            v = self.torch.as_tensor(_DeviceArray(v_ptr, (self.h.n, self.h.k), typestr), device=self.device)
            self._views = (u, v)
        else:
            self.h.device_factors()        # tells the library its cached Gram matrices are stale
        return self._views

    def invalidate(self):
        self._views = None

    def half_step(self, side, lam):
        self.h.als_half_step(side, lam)    # returns with the library's stream idle

    def after_exchange(self):
        self.torch.cuda.synchronize(self.device)

    def rmse_terms(self):
        return self.h.rmse_terms()


class ShardedALS(object):
    """User/item-sharded ALS epochs over an engine + a process group."""

    def __init__(self, engine, user_bounds, item_bounds, rank, n_users, n_items, group=None):
        self.engine, self.rank, self.group = engine, rank, group
        self.user_bounds, self.item_bounds = list(user_bounds), list(item_bounds)
        self.n_users, self.n_items = n_users, n_items
        engine.set_shard((self.user_bounds[rank], self.user_bounds[rank + 1]),
                         (self.item_bounds[rank], self.item_bounds[rank + 1]))

    def half_step(self, side, lam):
        self.engine.half_step(side, lam)
        u, v = self.engine.factor_tensors()
        exchange_rows(u if side == 0 else v, self.user_bounds if side == 0 else self.item_bounds, self.group)
        self.engine.after_exchange()

    def rmse(self):
        import torch
        import torch.distributed as dist
        frob, s1, s2 = self.engine.rmse_terms()
        u, _ = self.engine.factor_tensors()
        part = torch.tensor([s1, s2], dtype=torch.float64, device=u.device)
        dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.group)
        s1, s2 = part.tolist()
        rss = max(frob - 2.0 * s1 + s2, 0.0)
        return float(numpy.sqrt(rss / (float(self.n_users) * float(self.n_items))))

    def epoch(self, lam):
        self.half_step(0, lam)
        self.half_step(1, lam)
        return self.rmse()

    def train(self, n_iter, lam, start_error=numpy.inf):
        """partial_train's loop (collaborative_filtering.py:241-259) over sharded epochs; every rank
        sees the same RMSE, so every rank takes the same early-stop decision."""
        error, history = start_error, []
        for _ in range(n_iter):
            old = error
            error = self.epoch(lam)
            history.append(error)
            if error >= old:
                break
        return history


# --------------------------------------------------------------------------------------------------
# exchange inside the library (NVLink peer memory)
# --------------------------------------------------------------------------------------------------
def attach_exchange(handle, rank, world, user_bounds, item_bounds, group=None):
    """Map every rank's exchange arena into every other rank (CUDA IPC handles carried by
    ``all_gather_object``) and hand the row partition to the library.  Collective.  Returns True when every
    rank attached; on any failure every rank detaches and False is returned (callers fall back to ShardedALS)."""
    import torch.distributed as dist
    from hyprec_b200._native import HyprecError
    blob, err = b"", None
    try:
        blob = handle.comm_export(rank, world)
    except HyprecError as exc:
        err = str(exc)
    gathered = [None] * world
    dist.all_gather_object(gathered, (err, blob), group=group)
    failed = [g[0] for g in gathered if g[0] is not None]
    if not failed:
        try:
            handle.comm_attach([g[1] for g in gathered], user_bounds, item_bounds)
        except HyprecError as exc:
            err = str(exc)
    flags = [None] * world
    dist.all_gather_object(flags, err, group=group)
    failed = [f for f in flags if f is not None]
    if failed:
        if rank == 0:
            import sys
            sys.stderr.write("hyprec_b200: peer-memory exchange unavailable (%s); falling back to torch.distributed\n" % failed[0])
        try:
            handle.comm_detach()
        except HyprecError:
            pass
        dist.barrier(group=group)
        return False
    return True


def detach_exchange(handle, group=None):
    """Collective: nobody unmaps or frees an arena a peer may still store into."""
    import torch.distributed as dist
    dist.barrier(group=group)
    handle.comm_detach()
    dist.barrier(group=group)


class NativeShardedALS(object):
    """ShardedALS's interface over a handle with the exchange attached: every call is one C-ABI call."""

    def __init__(self, handle, user_bounds, item_bounds, rank):
        self.h, self.rank = handle, rank
        self.user_bounds, self.item_bounds = list(user_bounds), list(item_bounds)

    def half_step(self, side, lam):
        self.h.als_half_step(side, lam)

    def rmse(self):
        return self.h.rmse()

    def epoch(self, lam):
        self.h.als_half_step(0, lam)
        self.h.als_half_step(1, lam)
        return self.h.rmse()

    def train(self, n_iter, lam, start_error=numpy.inf, callback=None):
        history = []

        def record(epoch, rmse, seconds):
            history.append(rmse)
            if callback is not None:
                callback(epoch, rmse, seconds)

        self.h.train(n_iter, lam, start_error, record)
        return history


def make_sharded(handle, device_index, rank, world, user_bounds, item_bounds, n_users, n_items, group=None,
                 allow_peer_memory=True):
    """The sharded trainer for a handle whose data and factors are loaded: peer-memory exchange when it can be
    set up on every rank, else the torch.distributed exchange."""
    import os
    if allow_peer_memory and os.environ.get("HYPREC_EXCHANGE", "peer") != "nccl":
        if attach_exchange(handle, rank, world, user_bounds, item_bounds, group):
            return NativeShardedALS(handle, user_bounds, item_bounds, rank)
    return ShardedALS(NativeEngine(handle, device_index), user_bounds, item_bounds, rank, n_users, n_items, group)
