"""Under torchrun (one process per rank): the sharded epoch loop with the exchange inside the library (CUDA IPC
arenas, peer stores from the kernels' epilogues, mailbox reductions, device barriers) against the float64 oracle,
in float64 and float32; then the reference-facing class on the golden runs of the unmodified reference.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_exchange_check.py

One rank per GPU only: ranks wait for one another on the device (barrier kernels), and kernels that wait on one
another must not share a GPU (/opt/skills/guides/B200_PROFILING.md).
"""
import argparse
import os
import sys

import numpy
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from hyprec_b200 import _native, distributed, synth  # noqa: E402
from oracle import als as oals  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true", help="the small cases only")
args = ap.parse_args()

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
ok = True


def say(msg):
    print("rank %d %s" % (rank, msg), flush=True)


def toy(m, n, nnz, seed, heavy_item=None):
    full = synth.ratings_csr("toy", n_users=m, n_items=n, positives=nnz, min_degree=1, seed=seed)
    train = oals.to_csr(full)
    if heavy_item is not None:          # an item most users rated: one long row on the item side
        dense = numpy.asarray(train.todense())
        dense[::2, heavy_item] = 1.0
        dense[1::3, heavy_item] = 1.0
        train = oals.to_csr(dense)
    return train


# (precision, k, tolerance vs the float64 oracle, shape, prior?, skewed shard?)
CASES = [
    (_native.F64, 24, 1e-8, (150, 500, 7000), False, False),
    (_native.F64, 40, 1e-8, (90, 260, 3000), True, False),
    (_native.F32, 32, 2e-4, (100, 300, 4000), False, True),
    (_native.F32, 64, 2e-4, (150, 500, 7000), False, False),
]
if not args.quick:
    CASES += [
        (_native.F64, 200, 1e-8, (700, 1900, 60000), False, False),
        (_native.F32, 256, 2e-4, (700, 1900, 60000), False, False),
    ]

for precision, k, tol, (m, n, nnz), with_prior, skewed in CASES:
    train = toy(m, n, nnz, seed=13 + k, heavy_item=3)
    rs = numpy.random.RandomState(3)
    u0, v0 = rs.rand(m, k), rs.rand(n, k)
    theta = synth.dirichlet_theta(n, k, seed=3) if with_prior else None
    if with_prior:
        v0 = theta.copy()
    if skewed:            # an item shard that is empty on the last rank, an uneven user shard
        ub = [0] + [min(m, 10 * (r + 1)) for r in range(world - 1)] + [m]
        ib = [0] + [n] * world
    else:
        ub = distributed.partition_rows(train.indptr, world, k)
        ib = distributed.partition_rows(distributed.csc_indptr(train.indptr, train.indices, n), world, k)
    results = {}
    for mode in ("peer", "nccl"):
        h = _native.Handle(local, precision)
        h.set_train_csr(m, n, train.indptr, train.indices, train.data)
        h.set_factors(u0, v0)
        if with_prior:
            h.set_item_prior(theta)
        als = distributed.make_sharded(h, local, rank, world, ub, ib, m, n, allow_peer_memory=(mode == "peer"))
        native_path = isinstance(als, distributed.NativeShardedALS)
        if mode == "peer" and rank == 0:
            print("exchange: %s" % ("nvlink-peer-memory" if native_path else "FALLBACK"), flush=True)
        if mode == "peer" and not native_path:
            ok = False
        history = als.train(3, 0.01)
        u, v = h.get_factors()
        results[mode] = (history, u, v)
        if native_path:
            stats = h.comm_stats()
            ok = ok and stats["bytes_pushed"] > 0 and stats["barriers"] > 0
            distributed.detach_exchange(h)
        else:
            h.set_shard(0, -1, 0, -1)
        h.close()
    ou, ov = u0.copy(), v0.copy()
    _, ohist = oals.partial_train(ou, ov, train, 0.01, 3, prior=theta)
    for mode, (history, u, v) in results.items():
        eu = numpy.linalg.norm(u - ou) / numpy.linalg.norm(ou)
        ev = numpy.linalg.norm(v - ov) / numpy.linalg.norm(ov)
        good = eu <= tol and ev <= tol and len(history) == len(ohist) and numpy.allclose(history, ohist, rtol=max(tol, 1e-9))
        say("precision %d k %d prior %d skewed %d %s: rel err U %.2e V %.2e rmse %.12g -> %s" %
            (precision, k, with_prior, skewed, mode, eu, ev, history[-1], "OK" if good else "MISMATCH"))
        ok = ok and good
    # every rank holds the same replicas and took the same decisions (bit for bit)
    gathered = [None] * world
    dist.all_gather_object(gathered, (results["peer"][0], float(results["peer"][1].sum()), float(results["peer"][2].sum())))
    if any(g != gathered[0] for g in gathered):
        say("replicas differ between ranks: %s" % (gathered,))
        ok = False

# the reference-facing class under torchrun: the golden runs of the unmodified reference on every rank
from conftest import golden_folds, load_golden, split_lists  # noqa: E402
from hyprec_b200 import CollaborativeFiltering, Evaluator, ModelInitializer  # noqa: E402
for name in ("als_kfold_toy", "als_theta_toy", "als_naive_toy"):
    g = load_golden(name)
    hyper = {'n_factors': int(g["k"]), '_lambda': float(g["lam"])}
    options = {'k_folds': int(g["k_folds"]), 'n_iterations': int(g["n_iter"])}
    cf = CollaborativeFiltering(ModelInitializer(hyper.copy(), options['n_iterations']), Evaluator(g["ratings"]), hyper,
                                options, load_matrices=False, dump_matrices=False,
                                update_with_items=bool(int(g["update_with_items"])))
    theta = g["theta"].copy() if "theta" in g else None
    report = numpy.asarray(cf.train(theta), dtype=numpy.float64)
    want = g["mean_report"]
    same = numpy.isnan(report) & numpy.isnan(want) | (numpy.abs(report - want) <= 1e-9 * numpy.maximum(1, numpy.abs(want)))
    last = golden_folds(g)[-1]
    recs = [list(r) for r in split_lists(last["rec_ids"], last["rec_len"])]
    good = bool(numpy.all(same)) and numpy.allclose(cf.user_vecs, last["u"], rtol=1e-8, atol=1e-11) and \
        numpy.allclose(cf.item_vecs, last["v"], rtol=1e-8, atol=1e-11) and cf.evaluator.recommendation_indices == recs
    say("class %s: %s" % (name, "OK" if good else "MISMATCH %s vs %s" % (report, want)))
    ok = ok and good
    cf.close()

dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
