"""
Multi-GPU plumbing: one process per GPU (torchrun), rows of each half-step sharded across ranks,
one exchange of the freshly solved factor rows after every half-step.

The reference has no distributed code at all; what shards is the independence of rows inside
``als_step`` (/root/reference/lib/collaborative_filtering.py:82-86 reads only ``fixed_vecs``).
Each rank keeps a full replica of U and V, solves a contiguous, cost-balanced range of users
(user half-step) and of items (item half-step), and the ranges are exchanged through
``torch.distributed`` (NCCL over NVLink on GPUs; gloo in the CPU tests).  The RMSE is reduced
from per-shard partial sums.  Evaluation shards users; per-user integers are gathered on rank 0.

The engine behind ``ShardedALS`` is anything with ``half_step`` / ``factor_tensors`` /
``rmse_terms``: ``NativeEngine`` wraps the C-ABI handle, the CPU tests plug in an oracle-backed one.
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
# bench.py's N > 1 arm
# --------------------------------------------------------------------------------------------------
def bench_sharded(args, wl, rank, world, local, config, hooks=None):
    import time
    import torch
    import torch.distributed as dist
    from hyprec_b200 import _native

    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    precision = _native.F64 if args.precision == "fp64" else _native.F32
    h = _native.Handle(local, precision)
    m, n, k, lam = wl["m"], wl["n"], wl["k"], wl["lam"]
    tr, te, fu = wl["train_csr"], wl["test_csr"], wl["full_csr"]
    h.set_train_csr(m, n, tr.indptr, tr.indices, tr.data)
    h.set_eval_csr((fu.indptr, fu.indices, fu.data), (te.indptr, te.indices, te.data))
    h.set_factors(wl["u0"], wl["v0"])
    user_bounds = partition_rows(tr.indptr, world, k)
    item_bounds = partition_rows(csc_indptr(tr.indptr, tr.indices, n), world, k)
    engine = NativeEngine(h, local)
    als = ShardedALS(engine, user_bounds, item_bounds, rank, m, n)
    ulo, uhi = user_bounds[rank], user_bounds[rank + 1]
    has_eval = wl["cand_ids"] is not None
    n_tr = int(tr.indptr[uhi] - tr.indptr[ulo])
    n_te = int(te.indptr[uhi] - te.indptr[ulo])
    if has_eval:
        cp = wl["cand_indptr"][ulo:uhi + 1] - wl["cand_indptr"][ulo]
        ci = wl["cand_ids"][wl["cand_indptr"][ulo]:wl["cand_indptr"][uhi]]
        h.set_candidates(cp, ci, ulo, uhi)

    def evaluate():
        if has_eval:
            return h.eval_topk(200, None, None, ulo, uhi, n_train_nz=n_tr, n_test_nz=n_te)
        return None

    def step():
        rmse = als.epoch(lam)
        return rmse, evaluate()

    if hooks and "begin" in hooks:
        hooks["begin"]()               # clock sampling starts with the warm-up, not with the setup
    for _ in range(args.warmup):
        h.flush_l2()
        step()
    h.reset_kernel_ms()
    launches0 = h.launch_count
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    epoch_wall = 0.0
    total_wall = 0.0
    for _ in range(args.steps):
        h.flush_l2()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rmse = als.epoch(lam)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        out = evaluate()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        epoch_wall += t1 - t0
        total_wall += t2 - t0
    dist.barrier()
    torch.cuda.synchronize()
    if hooks and "end" in hooks:
        hooks["end"]()
    times = torch.tensor([epoch_wall, total_wall], dtype=torch.float64, device="cuda")
    dist.all_reduce(times, op=dist.ReduceOp.MAX)
    epoch_ms = 1e3 * times[0].item() / args.steps
    step_ms = 1e3 * times[1].item() / args.steps
    solve_ms = h.kernel_ms(_native.T_SOLVE)[0] / args.steps
    launches = torch.tensor([h.launch_count - launches0], dtype=torch.int64, device="cuda")
    dist.all_reduce(launches)
    line = dict(metric="als_epoch_ms", value=epoch_ms, unit="ms", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=step_ms, higher_is_better=False, scaling="strong", vs_baseline=None,
                dtype="f64" if args.precision == "fp64" else "f32", data="synthetic", config=config,
                gpu_launches=int(launches.item()), rmse=rmse,
                eval_users_per_s=(m / ((step_ms - epoch_ms) * 1e-3)) if has_eval else None,
                timing="wall clock around barrier + cuda synchronize on both sides, max over ranks "
                       "(the epoch mixes library-stream kernels and NCCL broadcasts)",
                rank0_solve_kernel_ms=solve_ms,
                e2e=dict(value=epoch_ms, unit="ms", h2d_bytes_per_step=0,
                         d2h_bytes_per_step=int((uhi - ulo) * 200 * 20 + n_tr + n_te),
                         note="factors stay device-resident across sharded epochs; evaluation inputs/outputs "
                              "cross PCIe every step"))
    h.close()
    if has_eval:
        # Reported next to the sharded number, never instead of it: the same epoch as N independent
        # replicas (one whole model per GPU, no data-path collective -- what k-fold cross-validation and
        # GridSearch do with N GPUs at this problem size, where a 3 ms epoch has nothing left to shard).
        line["replicas"] = _bench_replicas(args, wl, local, precision, world)
    dist.barrier()
    dist.destroy_process_group()
    return line


def _bench_replicas(args, wl, local, precision, world):
    import time
    import torch
    import torch.distributed as dist
    from hyprec_b200 import _native
    m, n, lam = wl["m"], wl["n"], wl["lam"]
    tr = wl["train_csr"]
    h = _native.Handle(local, precision)
    h.set_train_csr(m, n, tr.indptr, tr.indices, tr.data)
    h.set_factors(wl["u0"], wl["v0"])

    def step():
        h.als_half_step(_native.SIDE_USER, lam)
        h.als_half_step(_native.SIDE_ITEM, lam)
        h.rmse()

    for _ in range(max(1, args.warmup)):
        step()
    dist.barrier()
    torch.cuda.synchronize()
    wall = 0.0
    for _ in range(args.steps):
        h.flush_l2()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step()
        torch.cuda.synchronize()
        wall += time.perf_counter() - t0
    t = torch.tensor([wall], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    per_replica = 1e3 * t.item() / args.steps
    h.close()
    return dict(value=per_replica / world, unit="ms", scaling="weak", per_replica_epoch_ms=per_replica, models=world,
                note="aggregate ms per epoch when every GPU trains its own model (max over ranks / N); "
                     "no collective in the data path")
