"""Under torchrun (one rank per GPU): the reference-facing class trained and evaluated across ranks
reproduces the golden run of the unmodified reference on every rank.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_class_check.py
"""
import os
import sys

import numpy
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import golden_folds, load_golden, split_lists  # noqa: E402
from hyprec_b200 import CollaborativeFiltering, Evaluator, ModelInitializer  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank = dist.get_rank()
ok = True
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
    print("rank %d %s: %s" % (rank, name, "OK" if good else "MISMATCH %s vs %s" % (report, want)))
    ok = ok and good
    cf.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
