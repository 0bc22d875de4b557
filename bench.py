#!/usr/bin/env python
"""
bench.py -- the measurement contract.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload scaled|citeulike-a|citeulike-t|small] [--scale S] [--precision fp64|fp32]

One "step" = one pass of the hot path over the workload: one weighted-ALS epoch (user half-step,
item half-step, RMSE -- /root/reference/lib/collaborative_filtering.py:241-247) followed by the
evaluation report's device work (row thresholds, rounded counts, streamed top-200 per user, hit and
recall flags -- lib/abstract_recommender.py:100-131).

Default workload, every N (BASELINE.json's metric is quoted "at 1/2/4/8 B200", which only configs[3] is):
the scaled synthetic config -- 1M x 2M users x items, 2e8 positives, n_factors=256, float32 -- at --scale
(default 1.0; the scale is part of config.workload).  Rows of each half-step are sharded over the N ranks
(strong scaling: the problem is fixed); the evaluation is timed on a block of users per rank through
device-generated candidate lists.  The citeulike-a record of BASELINE configs[1] (5551 x 16980, k=200,
float64, 5-fold split fold 0, theta-initialised: epoch, evaluation, recall@{50..300}, roofline, CPU
baseline) rides in the same line under "citeulike_a" (measured on rank 0).

Printed JSON (one line, rank 0):
  metric/value   als_epoch_ms: device time of one epoch between two CUDA events on the library's stream
                 (hyprec_mark / hyprec_elapsed_ms), inputs resident in HBM, max over ranks; the same clock at
                 every N
  e2e            the same epoch through the C ABI with HOST buffers: factors host->device from pinned
                 memory, epoch, RMSE + factors device->host
  roofline       dominant kernel family of the epoch (the fused gather + Cholesky row solves): algorithmic
                 FLOPs / event time against the measured peak
  cpu_baseline   the oracle's restatement of the reference arithmetic timed on a bounded row sample
"""
import os
import sys

# torchrun exports OMP_NUM_THREADS=1: the CPU legs (reference arm, cpu_baseline) must use the host's cores,
# and the BLAS reads these variables when numpy is imported
if os.environ.get("HYPREC_BENCH_KEEP_THREADS") != "1":
    for _var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_var] = str(os.cpu_count() or 1)

import argparse   # noqa: E402
import hashlib    # noqa: E402
import json       # noqa: E402
import subprocess  # noqa: E402
import threading  # noqa: E402
import time       # noqa: E402

import numpy      # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (synth shape, n_factors, lambda, k_folds)
    "citeulike-a": ("citeulike-a", 200, 0.01, 5),
    "citeulike-t": ("citeulike-t", 300, 0.01, 5),
    "small": ("small", 50, 0.01, 5),
}
SCALED_SEED = 20170303
FACTOR_SEED = 4


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        counts = [int(p.get("num_threads", 1)) for p in threadpool_info() if p.get("user_api") in ("blas", "openmp")]
        return max(counts) if counts else 1
    except Exception:
        return int(os.environ.get("OMP_NUM_THREADS", "1"))


# --------------------------------------------------------------------------------------------------
# workloads: plain arrays, no scipy matrices at the scaled size
# --------------------------------------------------------------------------------------------------
def build_workload(name):
    """Synthetic ratings, the reference's 5-fold split (fold 0), theta-initialised item factors."""
    from hyprec_b200 import synth
    from hyprec_b200.evaluator import Evaluator
    import scipy.sparse as sp
    shape, k, lam, folds = WORKLOADS[name]
    t0 = time.time()
    ratings_csr = synth.ratings_csr(shape)
    ratings = numpy.asarray(ratings_csr.todense(), dtype=numpy.float64)
    evaluator = Evaluator(ratings)
    evaluator.set_kfolds(folds)
    lists = evaluator.get_kfold_indices()
    train, test = evaluator.get_fold(0, lists)
    cand_indptr, cand_ids = evaluator.candidate_lists(0)
    m, n = ratings.shape
    user_vecs = numpy.random.random((m, k))                  # drawn right after the split (App. A Q4)
    theta = synth.dirichlet_theta(n, k, seed=7)
    wl = dict(name=name, m=m, n=n, k=k, lam=lam, folds=folds, ratings=ratings, train=train, test=test,
              evaluator=evaluator, fold_lists=lists,
              train_csr=sp.csr_matrix(train), test_csr=sp.csr_matrix(test), full_csr=ratings_csr,
              cand_indptr=cand_indptr, cand_ids=cand_ids, u0=user_vecs, v0=theta.copy(), device_init=None)
    for key in ("train_csr", "test_csr", "full_csr"):
        wl[key].sort_indices()
    for key in ("train", "test", "full"):
        wl[key + "_arrays"] = (wl[key + "_csr"].indptr.astype(numpy.int64), wl[key + "_csr"].indices.astype(numpy.int32))
    wl["setup_s"] = time.time() - t0
    return wl


def build_scaled_workload(scale, device=None):
    """BASELINE.json configs[3] shape times `scale` (hyprec_b200.synth.scaled_ratings): CSR arrays only, factors
    drawn on the device (hyprec_init_factors_uniform, seed 4), no dense matrices, no host candidate lists."""
    from hyprec_b200 import synth
    t0 = time.time()
    data = synth.scaled_ratings(scale, device=device, seed=SCALED_SEED)
    wl = dict(name="scaled-%g" % scale, scale=scale, m=data["m"], n=data["n"], k=256, lam=0.01, folds=5, ratings=None,
              train=None, test=None, train_arrays=data["train"], test_arrays=data["test"], full_arrays=data["full"],
              cand_indptr=None, cand_ids=None, u0=None, v0=None, device_init=FACTOR_SEED)
    wl["setup_s"] = time.time() - t0
    return wl


def degrees(wl):
    indptr, indices = wl["train_arrays"]
    user_deg = numpy.diff(indptr).astype(numpy.float64)
    item_deg = numpy.bincount(indices, minlength=wl["n"]).astype(numpy.float64)
    return user_deg, item_deg


def lowrank_cap(k, elem):
    """Rows with more positives than this are solved as k x k systems (mirrors Engine::half_step)."""
    if elem == 8:
        return min(k, 127 if k <= 230 else 191)
    return min(k, 319)


def algorithmic_work(wl, elem=8):
    """FLOPs / bytes per epoch (SURVEY.md section 8d formulas) and the executed (low-rank) FLOP model.
    elem: bytes per factor element (the largest low-rank system depends on what fits shared memory)."""
    m, n, k = wl["m"], wl["n"], wl["k"]
    nnz = int(wl["train_arrays"][0][-1])
    solve_flops = 2 * (2.0 * nnz * k * k) + (m + n) * (k ** 3 / 3.0 + 2.0 * k * k) + 2 * (2.0 * nnz * k)
    cap = lowrank_cap(k, elem)
    lowrank = 0.0
    for deg in degrees(wl):
        light = deg <= cap
        d = deg[light]
        lowrank += float(numpy.sum(d * d * k + d ** 3 / 3.0 + 4.0 * d * k))
        lowrank += float(numpy.sum(2.0 * deg[~light] * k * k + k ** 3 / 3.0 + 2.0 * k * k))
    return dict(nnz=nnz, solve_flops_per_epoch=solve_flops, lowrank_flops_per_epoch=lowrank,
                solve_bytes_per_epoch=lambda s: 2.0 * nnz * (k * s + 4) + 2.0 * (m + n) * k * s + 8.0 * (m + n + 2),
                eval_flops=2.0 * m * n * k)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf32=d["bf16_tflops"] / 2.0, source="measured (MEASURED_PEAKS.json)")
    # the profiling recipe's fallback figures for a B200
    return dict(hbm=6650.0, tf32=1635.0 / 2.0, source="fallback (B200_PROFILING.md)")


def csrc_fingerprint():
    """sha256 over the kernel sources: recorded ncu traffic figures only apply to the kernels they were taken from.
    (host_*.cuh hold host-only code -- the split's random stream -- and launch nothing.)"""
    sha = hashlib.sha256()
    base = os.path.join(ROOT, "hyprec_b200", "csrc")
    for name in sorted(os.listdir(base)):
        if name.startswith("host_"):
            continue
        with open(os.path.join(base, name), "rb") as f:
            sha.update(f.read())
    return sha.hexdigest()[:16]


def recorded_traffic(key):
    """DRAM bytes per launch family from a committed ncu capture (profiles/traffic.json: value, command, commit,
    csrc fingerprint); None when the kernels changed since the capture."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None, "no capture recorded"
    rec = json.load(open(path)).get(key)
    if rec is None:
        return None, "no capture recorded for " + key
    if rec.get("csrc_sha") != csrc_fingerprint():
        return None, "capture %s is of older kernels (commit %s)" % (rec.get("file"), rec.get("commit"))
    return rec["dram_bytes"], "%s, `%s` at commit %s" % (rec.get("file"), rec.get("command"), rec.get("commit"))


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler(object):
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            if len(s) < 9:
                continue
            try:
                sm.append(float(s[1]))
                mx.append(float(s[2]))
                power.append(float(s[3]))
            except ValueError:
                continue
            for name, flag in zip(names, s[5:9]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        # "under load": a sample taken while the GPU idles between two short steps shows the idle clock,
        # which says nothing about the timed kernels -- keep the samples in the upper half of the power range
        loaded = sm
        if power and max(power) > min(power):
            cut = min(power) + 0.5 * (max(power) - min(power))
            loaded = [c for c, p_w in zip(sm, power) if p_w >= cut] or sm
        return dict(sm_mhz=float(numpy.median(loaded)) if loaded else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(power) if power else None, samples=len(sm), samples_under_load=len(loaded),
                    sm_mhz_all_samples=float(numpy.median(sm)) if sm else None, reasons=sorted(reasons))


# --------------------------------------------------------------------------------------------------
# CPU baseline (the oracle timed on the host: the only place bench.py executes oracle/)
# --------------------------------------------------------------------------------------------------
def sub_csr(indptr, indices, rows, n_cols):
    """scipy CSR of the given rows of an (indptr, indices) binary matrix."""
    import scipy.sparse as sp
    counts = indptr[rows + 1] - indptr[rows]
    ptr = numpy.zeros(len(rows) + 1, dtype=numpy.int64)
    numpy.cumsum(counts, out=ptr[1:])
    idx = numpy.concatenate([indices[indptr[r]:indptr[r + 1]] for r in rows]) if len(rows) else numpy.zeros(0, numpy.int32)
    return sp.csr_matrix((numpy.ones(len(idx)), idx, ptr), shape=(len(rows), n_cols))


def sub_csc(indptr, indices, cols, n_rows):
    """scipy CSR (one row per requested column) of the transposed binary matrix, rows ascending."""
    import scipy.sparse as sp
    cols = numpy.asarray(cols)
    where = numpy.full(int(indices.max()) + 1 if len(indices) else 1, -1, dtype=numpy.int64)
    where[cols] = numpy.arange(len(cols))
    hit = numpy.nonzero(where[indices] >= 0)[0]
    slot = where[indices[hit]]
    row_of = numpy.searchsorted(indptr, hit, side="right") - 1
    order = numpy.lexsort((row_of, slot))
    ptr = numpy.zeros(len(cols) + 1, dtype=numpy.int64)
    numpy.cumsum(numpy.bincount(slot, minlength=len(cols)), out=ptr[1:])
    return sp.csr_matrix((numpy.ones(len(hit)), row_of[order].astype(numpy.int32), ptr), shape=(len(cols), n_rows))


def cpu_epoch_sample(wl, u, v, n_rows, literal, seed=0):
    """Time the reference arithmetic on a row sample and extrapolate to one epoch (ms).

    literal=True : oracle.als.als_step_literal -- the reference's per-row dense
                   (k x n)(n x k) product + LAPACK solve (collaborative_filtering.py:82-98), with the
                   pure-Python confidence loop replaced by numpy.where (favours the reference).
    literal=False: the sparse-identity restatement (shared Gram + batched numpy solves); the Gram of the
                   fixed side is computed in full in every half-step, the row systems on the sample."""
    from oracle import als as oals
    rs = numpy.random.RandomState(seed)
    m, n, lam = wl["m"], wl["n"], wl["lam"]
    indptr, indices = wl["train_arrays"]
    users = numpy.sort(rs.choice(m, min(n_rows, m), replace=False))
    items = numpy.sort(rs.choice(n, min(n_rows, n), replace=False))
    if literal:
        train = wl["train"]
        t0 = time.perf_counter()
        oals.als_step_literal(u[users].copy(), v, train[users], lam, "user")
        t_user = time.perf_counter() - t0
        t0 = time.perf_counter()
        oals.als_step_literal(v[items].copy(), u, train[:, items], lam, "item")
        t_item = time.perf_counter() - t0
        t_fixed = 0.0
    else:
        rows_u = sub_csr(indptr, indices, users, n)
        rows_i = sub_csc(indptr, indices, items, m)
        t0 = time.perf_counter()
        gram_v = v.T.dot(v)
        t_gv = time.perf_counter() - t0
        t0 = time.perf_counter()
        oals.als_half_step(u[users].copy(), v, rows_u, lam, gram=gram_v)
        t_user = time.perf_counter() - t0
        t0 = time.perf_counter()
        gram_u = u.T.dot(u)
        t_gu = time.perf_counter() - t0
        t0 = time.perf_counter()
        oals.als_half_step(v[items].copy(), u, rows_i, lam, gram=gram_u)
        t_item = time.perf_counter() - t0
        t_fixed = t_gv + t_gu
    t_rmse = 0.0
    if wl.get("train_csr") is not None and wl["train_csr"].nnz <= 2000000:   # a small term; skipped at scaled sizes
        t0 = time.perf_counter()
        oals.rmse_gram(u, v, wl["train_csr"])
        t_rmse = time.perf_counter() - t0
    epoch_ms = 1e3 * (t_user * m / len(users) + t_item * n / len(items) + t_fixed + t_rmse)
    return epoch_ms, dict(users=len(users), items=len(items), t_user_s=t_user, t_item_s=t_item, t_gram_s=t_fixed,
                          t_rmse_s=t_rmse)


def cpu_eval_sample(wl, n_users, seed=0):
    """Users/s of the literal streamed top-k + thresholds on a user sample (oracle, pure Python)."""
    from oracle import evaluation as oev
    rs = numpy.random.RandomState(seed)
    users = numpy.sort(rs.choice(wl["m"], min(n_users, wl["m"]), replace=False))
    u, v = wl["u0"], wl["v0"]
    t0 = time.perf_counter()
    for uid in users:
        row = u[uid].dot(v.T)
        thr = sum(row) / row.shape[0]                                   # abstract_recommender.py:183
        cand = wl["cand_ids"][wl["cand_indptr"][uid]:wl["cand_indptr"][uid + 1]]
        oev.top_recommendations_stream(cand, row[cand], 200)            # evaluator.py:226-230
        _ = row >= thr
    return len(users) / (time.perf_counter() - t0)


def host_factors(wl, dtype=numpy.float64):
    """Initial factors on the host for the CPU legs: the workload's arrays, or for device-initialised workloads
    uniform [0, 1) float32 values of the same shape (the CPU arithmetic's cost does not depend on the values)."""
    if wl["u0"] is not None:
        return wl["u0"], wl["v0"]
    rs = numpy.random.RandomState(FACTOR_SEED)
    u = rs.random_sample((wl["m"], wl["k"])).astype(numpy.float32).astype(dtype)
    v = rs.random_sample((wl["n"], wl["k"])).astype(numpy.float32).astype(dtype)
    return u, v


def reference_arm(args, wl):
    """--impl reference: the reference's CPU arithmetic for the same metric/config, bounded sample per step."""
    cores = os.cpu_count()
    literal = wl["train"] is not None          # the literal arithmetic needs the dense train matrix
    u, v = host_factors(wl)
    per_step, per_step_vec = [], []
    detail = {}
    for step in range(args.warmup + args.steps):
        ms, detail = cpu_epoch_sample(wl, u, v, args.cpu_rows, literal=literal, seed=step)
        if step >= args.warmup:
            per_step.append(ms)
    if literal:
        for step in range(2):
            ms_v, _ = cpu_epoch_sample(wl, u, v, 8 * args.cpu_rows, literal=False, seed=step)
            per_step_vec.append(ms_v)
    value = float(numpy.mean(per_step))
    eval_rate = cpu_eval_sample(wl, 24) if wl["cand_ids"] is not None else None
    what = ("the literal per-row arithmetic (dense (k x n)(n x k) product + LAPACK solve per row)" if literal else
            "the vectorised sparse-identity restatement (full Gram of the fixed side + batched k x k solves; the literal "
            "arithmetic needs a dense users x items matrix that cannot exist at this size)")
    sample = ("%d users + %d items (random rows, full width) through %s, extrapolated x m/%d and x n/%d; per step" %
              (detail["users"], detail["items"], what, detail["users"], detail["items"]))
    line = dict(impl="reference", metric="als_epoch_ms", value=value, unit="ms", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=value, higher_is_better=False, scaling="strong", vs_baseline=None,
                dtype="f64", data="synthetic", config=config_block(wl, args),
                eval_users_per_s=eval_rate,
                cpu_baseline=dict(value=value, unit="ms", cores=cores, blas_threads=blas_threads(), kind="port", sample=sample,
                                  vectorised_value=float(numpy.mean(per_step_vec)) if per_step_vec else (None if literal else value)),
                e2e=dict(value=value, unit="ms", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def config_block(wl, args):
    nnz = int(wl["train_arrays"][0][-1])
    if wl["cand_ids"] is None:
        what = ("%s: BASELINE configs[3] (1M x 2M, 2e8 positives) x scale %g = %dx%d, %d train positives (random 80/20 split), "
                "n_factors=%d, lambda=%g, uniform float32 initial factors drawn on the device; evaluation timed on a user "
                "block per rank through device-generated candidate lists (test positives + a hashed 1/5 of the unrated items)" %
                (wl["name"], wl.get("scale", 1.0), wl["m"], wl["n"], nnz, wl["k"], wl["lam"]))
    else:
        what = ("%s shape %dx%d, %d train positives (5-fold split, fold 0), n_factors=%d, lambda=%g, "
                "theta-initialised item factors, top-200 over %d candidates/user" %
                (wl["name"], wl["m"], wl["n"], nnz, wl["k"], wl["lam"],
                 int(wl["cand_indptr"][1] - wl["cand_indptr"][0])))
    return dict(workload=what, users=wl["m"], items=wl["n"], n_factors=wl["k"], train_nnz=nnz,
                precision=args.precision, l2="flushed between timed steps (384 MiB memset)",
                parallelism="rows of each half-step sharded over %d GPU(s); solved rows stored into every replica over "
                            "NVLink peer memory" % args.gpus)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
CLASSES = ("gram", "solve", "rmse", "sweep", "topk", "flags", "count", "prep", "comm")


def kernel_classes(_native):
    return dict(gram=_native.T_GRAM, solve=_native.T_SOLVE, prep=_native.T_PREP, rmse=_native.T_RMSE,
                count=_native.T_COUNT, topk=_native.T_TOPK, flags=_native.T_FLAGS, comm=_native.T_COMM)


class SingleALS(object):
    def __init__(self, handle):
        self.h = handle

    def epoch(self, lam):
        self.h.als_half_step(0, lam)
        self.h.als_half_step(1, lam)
        return self.h.rmse()


def class_fold_e2e(wl, args, n_iter=5, reps=3):
    """One fold of hyprec_b200.CollaborativeFiltering.train() (the call runnables.py makes), wall clock."""
    from hyprec_b200 import CollaborativeFiltering, ModelInitializer
    prev = os.environ.get("HYPREC_PRECISION")
    os.environ["HYPREC_PRECISION"] = args.precision
    try:
        hyper = {'n_factors': wl["k"], '_lambda': wl["lam"]}
        options = {'k_folds': wl["folds"], 'n_iterations': n_iter}
        ev = wl["evaluator"]
        cf = CollaborativeFiltering(ModelInitializer(hyper.copy(), n_iter), ev, hyper, options, load_matrices=False,
                                    dump_matrices=False)
        cf.fold_test_indices = wl["fold_lists"]
        times, parts, report = [], {}, None

        def timed(name, fn, *a):
            t0 = time.perf_counter()
            out = fn(*a)
            parts[name] = parts.get(name, 0.0) + time.perf_counter() - t0
            return out

        for rep in range(1 + reps):
            parts = {}
            cf._train_uploaded = cf._eval_uploaded = cf._cands_uploaded = None      # a new fold: everything is uploaded again
            theta = wl["v0"].copy()
            t0 = time.perf_counter()
            cf.set_data(wl["train"], wl["test"])
            cf.hyperparameters['fold'] = 0
            matrices_found = timed("init_factors_s", cf._init_fold, theta) if hasattr(cf, "_init_fold") else None
            if matrices_found is None:
                report = timed("train_one_fold_s", cf.train_one_fold, theta)
            else:
                timed("partial_train_s", cf.partial_train)
                report = timed("evaluation_report_s", cf.get_evaluation_report)
            if rep > 0:
                times.append(time.perf_counter() - t0)
        cf.close()
        return dict(value=1e3 * float(numpy.mean(times)), unit="ms", epochs=n_iter, reps=reps,
                    per_epoch_equivalent_ms=1e3 * float(numpy.mean(times)) / n_iter, last_rep_parts_s=parts,
                    report=[float(x) for x in report],
                    note="set_data + initial factors + %d epochs + evaluation report through the class, host arrays in, "
                         "11-tuple out; train / eval CSR and candidate lists uploaded inside the timed region" % n_iter)
    finally:
        if prev is None:
            os.environ.pop("HYPREC_PRECISION", None)
        else:
            os.environ["HYPREC_PRECISION"] = prev


def solve_roofline(wl, precision, world, solve_ms, lowrank, fma_peak):
    """Roofline view of the row-solve kernels of one epoch on one rank: its 1/world share of the epoch's FLOPs
    (executed low-rank model, or SURVEY 8d's direct formula with HYPREC_LOWRANK=0) over their device time."""
    elem = 8 if precision == "fp64" else 4
    work = algorithmic_work(wl, elem)
    peaks = measured_peaks()
    flops_epoch = work["lowrank_flops_per_epoch"] if lowrank else work["solve_flops_per_epoch"]
    flops = flops_epoch / float(world)
    achieved = flops / (max(solve_ms, 1e-9) * 1e-3) / 1e12
    k = wl["k"]
    tensor_rows = precision == "fp32" and lowrank and k % 4 == 0 and 64 <= k <= 256 and \
        os.environ.get("HYPREC_TC", "1") != "0" and os.environ.get("HYPREC_TC_SOLVE", "1") != "0"
    if tensor_rows:
        kernel = ("tc_chol_rows_kernel + tc_chol_multi_kernel + split_partial_kernel (row systems of one epoch: normal matrix "
                  "accumulated by tcgen05 3xTF32 into tensor memory, 32-column panel Cholesky with tcgen05 trailing updates; rows "
                  "with more positives than n_factors as k x k systems -- the heaviest split over several CTAs -- the rest as "
                  "deg x deg low-rank systems, small ones several per CTA)")
        bound, peak, src = "tensor", peaks["tf32"], peaks["source"] + " bf16_tflops / 2 (dense TF32 rate); 3 TF32 products per algorithmic one"
    else:
        kernel = "als_solve_rows_kernel (row systems of one epoch; %s)" % (
            "deg x deg low-rank systems in concurrent degree buckets + direct k x k for heavy rows"
            if lowrank else "direct k x k systems")
        bound, peak, src = "%s_fma" % precision, fma_peak, "hyprec_probe_fma_tflops on this GPU (FMA chains, best of 4)"
    bytes_epoch = work["solve_bytes_per_epoch"](elem) / float(world)
    gbs = bytes_epoch / (max(solve_ms, 1e-9) * 1e-3) / 1e9
    return dict(kernel=kernel, bound=bound, achieved=achieved, peak=peak, unit="TFLOP/s",
                frac=(achieved / peak) if peak else None, peak_source=src, traffic=None, traffic_source=None,
                flops_per_epoch=flops_epoch, ranks=world,
                flops_model="low-rank (sum d^2 k + d^3/3 + 4 d k; k x k for rows above the cap)" if lowrank else "SURVEY 8d direct formula",
                direct_formula_flops=work["solve_flops_per_epoch"],
                hbm=dict(achieved=gbs, peak=peaks["hbm"], unit="GB/s", frac=gbs / peaks["hbm"], peak_source=peaks["source"],
                         algorithmic_bytes=bytes_epoch))


def sharded_roofline(wl, precision, world, solve_ms):
    """(kept for callers of the round-1 name) rank 0's share of a sharded epoch."""
    return solve_roofline(wl, precision, world, solve_ms, True, 37.0 if precision == "fp64" else 74.0)


def run_workload(args, wl, rank, world, local, with_cpu=True, quiet=False):
    """Our arm on one workload: returns the JSON line (a dict) on rank 0, None elsewhere."""
    from hyprec_b200 import _native, distributed
    from hyprec_b200 import metrics as hm
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()

    def reduce_max(values):
        if dist is None:
            return list(values)
        import torch
        t = torch.tensor(list(values), dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def reduce_sum(values):
        if dist is None:
            return list(values)
        import torch
        t = torch.tensor(list(values), dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.tolist()

    precision = _native.F64 if args.precision == "fp64" else _native.F32
    h = _native.Handle(local, precision)
    m, n, k, lam = wl["m"], wl["n"], wl["k"], wl["lam"]
    (tr_ptr, tr_idx), (te_ptr, te_idx), (fu_ptr, fu_idx) = wl["train_arrays"], wl["test_arrays"], wl["full_arrays"]
    t_up = time.time()
    h.set_train_csr(m, n, tr_ptr, tr_idx, None)
    h.set_eval_csr((fu_ptr, fu_idx, None), (te_ptr, te_idx, None))
    if wl["device_init"] is not None:
        h.init_factors_uniform(k, wl["device_init"])
    else:
        h.set_factors(wl["u0"], wl["v0"])
    upload_s = time.time() - t_up
    classes = kernel_classes(_native)

    # ---- row partition and the exchange
    if world > 1:
        user_bounds = distributed.partition_rows(tr_ptr, world, k)
        item_bounds = distributed.partition_rows(distributed.csc_indptr(tr_ptr, tr_idx, n), world, k)
        als = distributed.make_sharded(h, local, rank, world, user_bounds, item_bounds, m, n)
        exchange = "nvlink-peer-memory" if isinstance(als, distributed.NativeShardedALS) else "torch.distributed broadcasts (fallback)"
    else:
        user_bounds, item_bounds = [0, m], [0, n]
        als = SingleALS(h)
        exchange = None
    ulo, uhi = user_bounds[rank], user_bounds[rank + 1]

    # ---- evaluation block of this rank
    has_lists = wl["cand_ids"] is not None
    if has_lists:
        elo, ehi = ulo, uhi
        cp = wl["cand_indptr"][elo:ehi + 1] - wl["cand_indptr"][elo]
        ci = wl["cand_ids"][wl["cand_indptr"][elo]:wl["cand_indptr"][ehi]]
        h.set_candidates(cp, ci, elo, ehi)                    # per fold, like the train CSR
    else:
        block = int(args.eval_users) if args.eval_users > 0 else max(256, min(4096, int(2048 / max(wl.get("scale", 1.0), 1e-3))))
        elo = ulo if args.eval_offset < 0 else max(ulo, min(uhi, int(args.eval_offset)))
        ehi = min(uhi, elo + block)
        h.set_hashed_candidates(0, 5, SCALED_SEED)
    n_tr = int(tr_ptr[ehi] - tr_ptr[elo])
    n_te = int(te_ptr[ehi] - te_ptr[elo])

    def evaluate():
        if ehi <= elo:
            return None
        return h.eval_topk(200, None, None, elo, ehi, n_train_nz=n_tr, n_test_nz=n_te)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler is not None:
        sampler.start()                # started before the warm-up (same load) so that short timed
        time.sleep(0.3)                # regions still collect samples
    rmse = None
    for _ in range(args.warmup):
        h.flush_l2()
        rmse = als.epoch(lam)
        h.flush_l2()
        evaluate()
    h.reset_kernel_ms()
    launches0 = h.launch_count
    comm0 = h.comm_stats()["bytes_pushed"] if world > 1 else 0.0
    epoch_ms, eval_ms, wall = [], [], []
    out = None
    for _ in range(args.steps):
        h.flush_l2()
        barrier()
        t0 = time.perf_counter()
        h.mark(0)
        rmse = als.epoch(lam)
        h.mark(1)
        epoch_ms.append(h.elapsed_ms(0, 1))
        h.flush_l2()
        h.mark(2)
        out = evaluate()
        h.mark(3)
        eval_ms.append(h.elapsed_ms(2, 3))
        wall.append(time.perf_counter() - t0)
    barrier()
    clocks = sampler.stop() if sampler is not None else None
    launches = h.launch_count - launches0
    comm_bytes = (h.comm_stats()["bytes_pushed"] - comm0) / args.steps if world > 1 else 0.0
    # the epoch ends when its slowest rank ends (the same events on every rank's stream)
    epoch_max = reduce_max(epoch_ms)
    eval_max = reduce_max(eval_ms)
    wall_max = reduce_max(wall)
    value = float(numpy.mean(epoch_max))
    eval_wall_value = float(numpy.mean(eval_max))       # events around the whole call: includes its device->host copies
    eval_users_total = reduce_sum([float(ehi - elo)])[0]
    launches_total = int(reduce_sum([float(launches)])[0])
    comm_total = reduce_sum([comm_bytes])[0]
    dev = {name: h.kernel_ms(t) for name, t in classes.items()}      # (total ms, regions, last)
    per_step = {name: dev[name][0] / args.steps for name in dev}
    eval_stats = h.eval_stats() if ehi > elo else None
    per_rank = None
    if dist is not None:
        per_rank = [None] * world
        dist.all_gather_object(per_rank, dict(rank=rank, users=[int(ulo), int(uhi)], epoch_ms=float(numpy.mean(epoch_ms)),
                                              eval_users=[int(elo), int(ehi)], eval_stats=eval_stats,
                                              **{name: round(per_step[name], 3) for name in per_step}))
    # device time of the evaluation's kernels (count sweep, top-k, flags) on the slowest rank
    eval_value = reduce_max([per_step["count"] + per_step["topk"] + per_step["flags"]])[0]

    # ---- e2e: host buffers in, host buffers out (pinned host memory, as the contract says), a bounded number of steps
    e2e_steps = max(1, min(args.steps, 3 if wl["device_init"] is not None else args.steps))
    u_host, v_host = h.pinned_empty((m, k)), h.pinned_empty((n, k))
    h.get_factors(u_host, v_host)
    eval_out = None
    if has_lists and ehi > elo:
        nu = ehi - elo
        eval_out = dict(thresholds=h.pinned_empty(nu), ones_per_row=h.pinned_empty(nu, numpy.int64),
                        topk_idx=h.pinned_empty((nu, 200), numpy.int32), topk_val=None,     # the class's report call:
                        topk_hit=h.pinned_empty((nu, 200), numpy.uint8),                      # hyprec_eval_topk_u8
                        train_flags=h.pinned_empty(max(n_tr, 1), numpy.uint8),
                        test_flags=h.pinned_empty(max(n_te, 1), numpy.uint8))
    e2e_epoch, e2e_eval = [], []
    for step in range(1 + e2e_steps):
        barrier()
        t0 = time.perf_counter()
        h.set_factors(u_host, v_host)
        als.epoch(lam)
        h.get_factors(u_host, v_host)
        t1 = time.perf_counter()
        if eval_out is not None:
            h.eval_topk(200, None, None, elo, ehi, n_train_nz=n_tr, n_test_nz=n_te, out=eval_out)
        t2 = time.perf_counter()
        if step >= 1:
            e2e_epoch.append(t1 - t0)
            e2e_eval.append(t2 - t1)
    e2e_epoch_ms = 1e3 * float(numpy.mean(reduce_max(e2e_epoch)))
    e2e_eval_ms = 1e3 * float(numpy.mean(reduce_max(e2e_eval)))
    factor_bytes = (m + n) * k * 8
    eval_d2h = (ehi - elo) * 200 * (4 + 1) + (ehi - elo) * 16 + n_tr + n_te if eval_out is not None else 0

    # ---- metrics of the final model (reported, not timed)
    recall200, recall_at = None, None
    if has_lists and out is not None and world == 1:
        test_likes = numpy.diff(te_ptr).astype(numpy.float64)
        recall200 = float(hm.recall_at_x(200, out["topk_hit"], test_likes))
        try:        # recall@{50..300} from one more untimed pass that keeps the top 300 of every candidate list
            top300 = h.eval_topk(300, None, None, want_counts=False)
            recall_at = {str(x): float(hm.recall_at_x(x, top300["topk_hit"], test_likes)) for x in (50, 100, 150, 200, 250, 300)}
        except Exception as exc:
            recall_at = dict(error=repr(exc))

    # ---- the same through the reference-facing class: ONE fold of CollaborativeFiltering.train() =
    # set_data -> train_one_fold (random user factors, theta item factors, n_iter epochs inside hyprec_train, factors
    # back into the host arrays) -> get_evaluation_report (device evaluation + the reference's metric arithmetic):
    # host matrices in, the 11-tuple out, every upload of the fold inside the timed region
    class_e2e = None
    if has_lists and world == 1 and wl.get("evaluator") is not None and not getattr(args, "no_class", False):
        try:
            class_e2e = class_fold_e2e(wl, args)
        except Exception as exc:
            class_e2e = dict(error=repr(exc))

    lowrank = os.environ.get("HYPREC_LOWRANK", "1") != "0"
    fma_peak = h.probe_fma_tflops(precision)
    line = None
    if rank == 0:
        roofline = solve_roofline(wl, args.precision, world, per_step["solve"], lowrank, fma_peak)
        key = "%s/%s/solve" % (wl["name"], args.precision)
        if world == 1:
            roofline["traffic"], roofline["traffic_source"] = recorded_traffic(key)
        work = algorithmic_work(wl, 8 if args.precision == "fp64" else 4)
        peaks = measured_peaks()
        if has_lists and per_step["count"] > 0:
            sweep_tflops = work["eval_flops"] / float(world) / (per_step["count"] * 1e-3) / 1e12
            roofline["other_kernels"] = dict(tc_score_sweep=dict(
                kernel="tc::tc_score_kernel (tcgen05 kind::tf32, 3xTF32 split) + split / norm / re-check helpers",
                bound="tensor", achieved=3.0 * sweep_tflops, peak=peaks["tf32"], unit="TFLOP/s",
                frac=3.0 * sweep_tflops / peaks["tf32"], flops=3.0 * work["eval_flops"],
                traffic=recorded_traffic("%s/%s/sweep" % (wl["name"], args.precision))[0]))
        line = dict(metric="als_epoch_ms", value=value, unit="ms", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=1e3 * float(numpy.mean(wall_max)), higher_is_better=False, scaling="strong",
                    vs_baseline=None, dtype="f64" if args.precision == "fp64" else "f32", data="synthetic",
                    config=config_block(wl, args), clocks=clocks,
                    e2e=dict(value=e2e_epoch_ms, unit="ms", h2d_bytes_per_step=factor_bytes,
                             d2h_bytes_per_step=factor_bytes + 24 + eval_d2h, steps=e2e_steps, eval_ms=e2e_eval_ms if eval_out is not None else None,
                             eval_users_per_s=(eval_users_total / (e2e_eval_ms * 1e-3)) if eval_out is not None and e2e_eval_ms > 0 else None,
                             note="per step and rank: hyprec_set_factors from pinned float64 host arrays, the epoch, hyprec_rmse, "
                                  "hyprec_get_factors back into them%s" % ("; then the evaluation's outputs into pinned host arrays" if eval_out is not None else "")),
                    gpu_launches=launches_total, roofline=roofline,
                    timing="CUDA events on the library's stream around the epoch (hyprec_mark), max over ranks per step, mean over steps",
                    eval_ms=eval_value, eval_call_ms=eval_wall_value, eval_users=int(eval_users_total),
                    eval_users_per_s=(eval_users_total / (eval_value * 1e-3)) if eval_value > 0 else None,
                    kernel_ms_per_step=per_step,
                    kernel_ms_note="event-timed regions of rank 0's main stream; the row-solve launches of a half-step run on their own "
                                   "streams at the same time, so 'prep' (whitening) and the tail of 'solve' include waiting for SMs "
                                   "those launches occupy -- the launch lists under profiles/ have the per-kernel times",
                    rmse=rmse, recall_at_200=recall200, setup_s=wl["setup_s"], upload_s=upload_s)
        if class_e2e is not None:
            line["e2e"]["class_fold"] = class_e2e
        if recall_at is not None:
            line["recall_at"] = recall_at
        if eval_stats is not None:
            line["eval_topk"] = dict(eval_stats, note="fused: the sweep's epilogue emits only candidate cells above a per-user cut "
                                                      "(no users x items score buffer); fallback_users were scored exactly instead")
        if per_rank is not None:
            line["per_rank_kernel_ms"] = per_rank
        if world > 1:
            line["exchange"] = exchange
            line["comm_bytes_per_epoch"] = comm_total
            line["rank0_solve_kernel_ms"] = per_step["solve"]
            comm_ms = per_step["comm"]
            line["comm"] = dict(bytes_per_epoch_all_ranks=comm_total, rank0_barrier_and_push_ms=comm_ms,
                                note="bytes every rank stored into the other replicas per epoch (factor rows, whitened rows, Gram "
                                     "mailbox); the factor rows leave with one copy kernel per range right behind the GEMM that produced them "
                                     "(timed inside 'prep'), HYPREC_PUSH=fused stores them from the GEMM epilogues instead")

    # ---- CPU baseline on this box's host cores (bounded sample), rank 0, single-GPU runs only
    if rank == 0 and with_cpu and world == 1:
        literal = wl["train"] is not None
        if wl["u0"] is not None:
            cu, cv = wl["u0"], wl["v0"]
        else:
            cu, cv = numpy.array(u_host), numpy.array(v_host)
        cpu_ms, cpu_detail = cpu_epoch_sample(wl, cu, cv, args.cpu_rows, literal=literal)
        cpu_vec_ms = cpu_ms
        if literal:
            cpu_vec_ms, _ = cpu_epoch_sample(wl, cu, cv, 8 * args.cpu_rows, literal=False)
        cpu_eval = cpu_eval_sample(wl, 16) if has_lists else None
        how = ("the reference's per-row dense arithmetic (oracle.als.als_step_literal)" if literal else
               "the vectorised sparse-identity restatement (oracle.als.als_half_step; full Gram per half-step)")
        line["cpu_baseline"] = dict(value=cpu_ms, unit="ms", cores=os.cpu_count(), blas_threads=blas_threads(), kind="port",
                                    sample="%d users + %d items at full width through %s, extrapolated to %d + %d rows%s" %
                                           (cpu_detail["users"], cpu_detail["items"], how, m, n, ", + RMSE" if literal else ""),
                                    vectorised_value=cpu_vec_ms, eval_users_per_s=cpu_eval)
    elif rank == 0:
        line["cpu_baseline"] = dict(value=None, unit="ms", cores=os.cpu_count(), kind="port",
                                    sample="timed at N=1 only (bench.py --gpus 1 and --impl reference)")
    if isinstance(als, distributed.NativeShardedALS):
        distributed.detach_exchange(h)
    try:
        h.close()
    except Exception:
        pass
    return line


def our_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    line = run_workload(args, wl, rank, world, local)
    # the citeulike-a record (BASELINE configs[1]) next to a scaled headline: rank 0, one GPU, float64
    if wl["device_init"] is not None and not args.no_citeulike:
        record = None
        if rank == 0:
            try:
                sub = argparse.Namespace(**vars(args))
                sub.precision, sub.gpus, sub.steps, sub.warmup = "fp64", 1, max(3, min(args.steps, 10)), max(3, min(args.warmup, 5))
                sub.cpu_rows = 16
                record = run_workload(sub, build_workload("citeulike-a"), 0, 1, local, with_cpu=(world == 1))
            except Exception as exc:       # the headline must not be lost over its companion record
                record = dict(error=repr(exc))
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        if rank == 0:
            line["citeulike_a"] = record
    if world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="scaled", choices=sorted(WORKLOADS) + ["scaled"])
    ap.add_argument("--scale", type=float, default=float(os.environ.get("HYPREC_BENCH_SCALE", "1.0")),
                    help="fraction of the 1M x 2M / 2e8 config (--workload scaled)")
    ap.add_argument("--precision", default=os.environ.get("HYPREC_PRECISION"), choices=["fp64", "fp32"],
                    help="default: fp32 (scaled workload, as BASELINE configs[3] specifies), fp64 (citeulike workloads)")
    ap.add_argument("--cpu-rows", type=int, default=None, help="rows per side in the CPU baseline sample")
    ap.add_argument("--eval-users", type=int, default=0, help="scaled workload: users per rank in the timed evaluation block")
    ap.add_argument("--eval-offset", type=int, default=-1, help="scaled workload: first user of the evaluation block (default: the rank's first user)")
    ap.add_argument("--no-citeulike", action="store_true", help="scaled workload: skip the citeulike-a companion record")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return
        reference_arm(args, make_workload(args))
        return
    our_arm(args, make_workload(args))


def make_workload(args):
    if args.workload == "scaled":
        if args.precision is None:
            args.precision = "fp32"          # configs[3] is specified in float32
        if args.cpu_rows is None:
            args.cpu_rows = 256
        device = None
        try:
            import torch
            if torch.cuda.is_available():
                device = "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0"))
        except Exception:
            device = "cpu"
        return build_scaled_workload(args.scale, device)
    if args.precision is None:
        args.precision = "fp64"
    if args.cpu_rows is None:
        args.cpu_rows = 64 if args.impl == "reference" else 16
    return build_workload(args.workload)


if __name__ == "__main__":
    main()
