#!/usr/bin/env python
"""
bench.py -- the measurement contract.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload citeulike-a|citeulike-t|small] [--precision fp64|fp32]

One "step" = one pass of the hot path over the workload: one weighted-ALS epoch (user half-step,
item half-step, RMSE -- /root/reference/lib/collaborative_filtering.py:241-247) followed by one
evaluation report's device work (row thresholds, rounded counts, streamed top-200 per user, hit and
recall flags -- lib/abstract_recommender.py:100-131).

Workload at N=1 (BASELINE.json configs[1]): citeulike-a shape (5551 x 16980, 204986 positives),
reference 5-fold split (seed 42) fold 0, n_factors=200, lambda=0.01, item factors initialised from a
synthetic LDA theta (Dirichlet rows), candidate lists exactly as load_top_recommendations reads them.

Printed JSON (one line, rank 0):
  metric/value   als_epoch_ms: device time (CUDA events on the launching stream) of one epoch's
                 kernels with every input resident in HBM
  e2e            the same epoch through the C ABI with HOST buffers: factors host->device, epoch,
                 factors device->host; eval through hyprec_eval_topk with host candidate lists
  roofline       dominant kernel (the fused gather+Cholesky row solve): algorithmic FLOPs / event time
                 against the FMA rate measured on this GPU by hyprec_probe_fma_tflops
  cpu_baseline   the oracle's restatement of the reference arithmetic timed on a bounded row sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (synth shape, n_factors, lambda, k_folds)
    "citeulike-a": ("citeulike-a", 200, 0.01, 5),
    "citeulike-t": ("citeulike-t", 300, 0.01, 5),
    "small": ("small", 50, 0.01, 5),
}


# --------------------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------------------
def build_workload(name):
    """Synthetic ratings, the reference's 5-fold split (fold 0), theta-initialised item factors."""
    from hyprec_b200 import synth
    from hyprec_b200.evaluator import Evaluator
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
    import scipy.sparse as sp
    wl = dict(name=name, m=m, n=n, k=k, lam=lam, folds=folds, ratings=ratings, train=train, test=test,
              train_csr=sp.csr_matrix(train), test_csr=sp.csr_matrix(test), full_csr=ratings_csr,
              cand_indptr=cand_indptr, cand_ids=cand_ids, u0=user_vecs, v0=theta.copy(),
              setup_s=time.time() - t0)
    for key in ("train_csr", "test_csr", "full_csr"):
        wl[key].sort_indices()
    return wl


def build_scaled_workload(scale):
    """BASELINE.json configs[3] shape times `scale`: 1M x 2M users x items, 2e8 positives (mean 200 per
    user, log-normal), Zipf(0.8) item popularity, n_factors=256, float32 uniform initial factors.
    Built directly as CSR; a random 20 % of the positives is held out (5-fold-equivalent split).
    No dense matrices and no per-user candidate lists exist at this size: epoch-only workload."""
    import scipy.sparse as sp
    t0 = time.time()
    m, n, nnz = int(1e6 * scale), int(2e6 * scale), int(2e8 * scale)
    k, lam = 256, 0.01
    rs = numpy.random.RandomState(20170303)
    raw = rs.lognormal(mean=numpy.log(200.0) - 0.32, sigma=0.8, size=m)
    deg = numpy.clip(numpy.round(raw * (nnz / raw.sum())), 1, n // 4).astype(numpy.int64)
    p = (numpy.arange(n) + 20.0) ** -0.8
    cdf = numpy.cumsum(p[rs.permutation(n)])
    cdf /= cdf[-1]
    rows = numpy.repeat(numpy.arange(m, dtype=numpy.int64), deg)
    items = numpy.searchsorted(cdf, rs.random_sample(rows.shape[0])).astype(numpy.int64)
    numpy.minimum(items, n - 1, out=items)
    keys = numpy.unique(rows * n + items)                 # drops the duplicates of sampling with replacement
    rows, items = keys // n, (keys % n).astype(numpy.int32)
    keep = rs.random_sample(keys.shape[0]) < 0.8
    def csr_of(mask):
        indptr = numpy.zeros(m + 1, dtype=numpy.int64)
        numpy.cumsum(numpy.bincount(rows[mask], minlength=m), out=indptr[1:])
        return sp.csr_matrix((numpy.ones(int(mask.sum())), items[mask], indptr), shape=(m, n))
    train, test = csr_of(keep), csr_of(~keep)
    full = sp.csr_matrix((numpy.ones(keys.shape[0]), items, numpy.concatenate(([0], numpy.cumsum(numpy.bincount(rows, minlength=m))))),
                         shape=(m, n))
    rs4 = numpy.random.RandomState(4)
    u0 = rs4.random_sample((m, k)).astype(numpy.float32).astype(numpy.float64)
    v0 = rs4.random_sample((n, k)).astype(numpy.float32).astype(numpy.float64)
    return dict(name="scaled-%g" % scale, m=m, n=n, k=k, lam=lam, folds=5, ratings=None, train=None, test=None,
                train_csr=train, test_csr=test, full_csr=full, cand_indptr=None, cand_ids=None, u0=u0, v0=v0,
                setup_s=time.time() - t0)


def algorithmic_work(wl, elem=8):
    """FLOPs / bytes per epoch and per launch of the solve kernel (SURVEY.md section 8d formulas).
    elem: bytes per factor element (the largest low-rank system depends on what fits shared memory)."""
    m, n, k = wl["m"], wl["n"], wl["k"]
    nnz = wl["train_csr"].nnz
    solve_flops = 2 * (2.0 * nnz * k * k) + (m + n) * (k ** 3 / 3.0 + 2.0 * k * k) + 2 * (2.0 * nnz * k)
    cap = min(k, 191 if elem == 8 else 319)
    lowrank = 0.0
    for mat, rows_fixed in ((wl["train_csr"], n), (wl["train_csr"].T.tocsr(), m)):
        deg = numpy.diff(mat.indptr).astype(numpy.float64)
        light = deg <= cap
        d = deg[light]
        lowrank += float(numpy.sum(d * d * k + d ** 3 / 3.0 + 4.0 * d * k))
        lowrank += float(numpy.sum(2.0 * deg[~light] * k * k + k ** 3 / 3.0 + 2.0 * k * k))
    return dict(nnz=nnz, solve_flops_per_epoch=solve_flops, lowrank_flops_per_epoch=lowrank,
                solve_bytes_per_epoch=lambda s: 2.0 * nnz * (k * s + 4) + 2.0 * (m + n) * k * s + 8.0 * (m + n + 2),
                eval_flops=2.0 * m * n * k)


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
def cpu_epoch_sample(wl, n_rows, literal, seed=0):
    """Time the reference arithmetic on a row sample and extrapolate to one epoch (ms).

    literal=True : oracle.als.als_step_literal -- the reference's per-row dense
                   (k x n)(n x k) product + LAPACK solve (collaborative_filtering.py:82-98), with the
                   pure-Python confidence loop replaced by numpy.where (favours the reference).
    literal=False: the sparse-identity restatement (batched numpy solves)."""
    from oracle import als as oals
    rs = numpy.random.RandomState(seed)
    m, n, k, lam = wl["m"], wl["n"], wl["k"], wl["lam"]
    users = numpy.sort(rs.choice(m, min(n_rows, m), replace=False))
    items = numpy.sort(rs.choice(n, min(n_rows, n), replace=False))
    u, v = wl["u0"].copy(), wl["v0"].copy()
    if literal:
        train = wl["train"]
        t0 = time.perf_counter()
        oals.als_step_literal(u[users], v, train[users], lam, "user")
        t_user = time.perf_counter() - t0
        t0 = time.perf_counter()
        oals.als_step_literal(v[items], u, train[:, items], lam, "item")
        t_item = time.perf_counter() - t0
    else:
        csr = wl["train_csr"]
        csc = csr.T.tocsr()
        t0 = time.perf_counter()
        oals.als_half_step(u, v, csr, lam, rows=users)
        t_user = time.perf_counter() - t0
        t0 = time.perf_counter()
        oals.als_half_step(v, u, csc, lam, rows=items)
        t_item = time.perf_counter() - t0
    t_rmse = 0.0
    if wl["train_csr"].nnz <= 2000000:       # the Gram-identity RMSE is a small term; skipped at scaled sizes
        t0 = time.perf_counter()
        oals.rmse_gram(u, v, wl["train_csr"])
        t_rmse = time.perf_counter() - t0
    epoch_ms = 1e3 * (t_user * m / len(users) + t_item * n / len(items) + t_rmse)
    return epoch_ms, dict(users=len(users), items=len(items), t_user_s=t_user, t_item_s=t_item, t_rmse_s=t_rmse)


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


def reference_arm(args, wl):
    """--impl reference: the reference's CPU arithmetic for the same metric/config, bounded sample."""
    cores = os.cpu_count()
    per_step = []
    for step in range(args.warmup + args.steps):
        ms, detail = cpu_epoch_sample(wl, args.cpu_rows, literal=wl["train"] is not None, seed=step)
        if step >= args.warmup:
            per_step.append(ms)
    value = float(numpy.mean(per_step))
    eval_rate = cpu_eval_sample(wl, 24) if wl["cand_ids"] is not None else None
    sample = ("%d users + %d items (random rows, full width) through the literal per-row arithmetic, "
              "extrapolated x m/%d and x n/%d, + full Gram-identity RMSE; per step" %
              (detail["users"], detail["items"], detail["users"], detail["items"]))
    line = dict(impl="reference", metric="als_epoch_ms", value=value, unit="ms", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=value, higher_is_better=False, scaling="strong", vs_baseline=None,
                dtype="f64", data="synthetic", config=config_block(wl, args),
                eval_users_per_s=eval_rate,
                cpu_baseline=dict(value=value, unit="ms", cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit="ms", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def config_block(wl, args):
    if wl["cand_ids"] is None:
        what = ("%s: %dx%d, %d train positives (random 80/20 split), n_factors=%d, lambda=%g, uniform float32 "
                "initial factors, ALS epochs only (no per-user candidate lists at this size)" %
                (wl["name"], wl["m"], wl["n"], wl["train_csr"].nnz, wl["k"], wl["lam"]))
    else:
        what = ("%s shape %dx%d, %d train positives (5-fold split, fold 0), n_factors=%d, lambda=%g, "
                "theta-initialised item factors, top-200 over %d candidates/user" %
                (wl["name"], wl["m"], wl["n"], wl["train_csr"].nnz, wl["k"], wl["lam"],
                 int(wl["cand_indptr"][1] - wl["cand_indptr"][0])))
    return dict(workload=what,
                users=wl["m"], items=wl["n"], n_factors=wl["k"], train_nnz=int(wl["train_csr"].nnz),
                precision=args.precision, l2="flushed between timed steps (384 MiB memset)",
                parallelism="rows sharded over %d GPU(s), factor all-gather per half-step" % args.gpus)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def our_arm(args, wl):
    from hyprec_b200 import _native
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        return our_arm_distributed(args, wl, rank, world, local)

    precision = _native.F64 if args.precision == "fp64" else _native.F32
    h = _native.Handle(local, precision)
    m, n, k, lam = wl["m"], wl["n"], wl["k"], wl["lam"]
    tr, te, fu = wl["train_csr"], wl["test_csr"], wl["full_csr"]
    h.set_train_csr(m, n, tr.indptr, tr.indices, tr.data)
    h.set_eval_csr((fu.indptr, fu.indices, fu.data), (te.indptr, te.indices, te.data))
    h.set_factors(wl["u0"], wl["v0"])
    classes = dict(gram=_native.T_GRAM, solve=_native.T_SOLVE, prep=_native.T_PREP, rmse=_native.T_RMSE,
                   count=_native.T_COUNT, topk=_native.T_TOPK, flags=_native.T_FLAGS)

    def epoch():
        h.als_half_step(_native.SIDE_USER, lam)
        h.als_half_step(_native.SIDE_ITEM, lam)
        return h.rmse()

    has_eval = wl["cand_ids"] is not None
    if has_eval:
        h.set_candidates(wl["cand_indptr"], wl["cand_ids"])      # per fold, like the train CSR

    def evaluate():
        if not has_eval:
            return None
        return h.eval_topk(200, None, None, n_train_nz=tr.nnz, n_test_nz=te.nnz)

    sampler = ClockSampler(local)      # started before the warm-up (same load) so that short timed
    sampler.start()                    # regions still collect samples
    time.sleep(0.3)
    for _ in range(args.warmup):
        h.flush_l2()
        epoch()
        h.flush_l2()
        evaluate()
    h.reset_kernel_ms()
    launches0 = h.launch_count
    wall_epoch = wall_eval = 0.0
    rmse = None
    for _ in range(args.steps):
        h.flush_l2()
        t0 = time.perf_counter()
        rmse = epoch()
        wall_epoch += time.perf_counter() - t0
        h.flush_l2()
        t0 = time.perf_counter()
        out = evaluate()
        wall_eval += time.perf_counter() - t0
    clocks = sampler.stop()
    launches = h.launch_count - launches0
    dev = {name: h.kernel_ms(t) for name, t in classes.items()}      # (total ms, regions, last)
    per_step = {name: dev[name][0] / args.steps for name in dev}
    epoch_ms = per_step["gram"] + per_step["solve"] + per_step["prep"] + per_step["rmse"]
    eval_ms = per_step["count"] + per_step["topk"] + per_step["flags"]

    # ---- e2e: host buffers in, host buffers out, every step (pinned host memory, as the contract says)
    u_host, v_host = h.pinned_empty((m, k)), h.pinned_empty((n, k))
    h.get_factors(u_host, v_host)
    eval_out = None
    if has_eval:
        eval_out = dict(thresholds=h.pinned_empty(m), ones_per_row=h.pinned_empty(m, numpy.int64),
                        topk_idx=h.pinned_empty((m, 200), numpy.int32), topk_val=h.pinned_empty((m, 200)),
                        topk_hit=h.pinned_empty((m, 200)), train_flags=h.pinned_empty(tr.nnz, numpy.uint8),
                        test_flags=h.pinned_empty(te.nnz, numpy.uint8))
    e2e_epoch = e2e_eval = 0.0
    for step in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        h.set_factors(u_host, v_host)
        epoch()
        h.get_factors(u_host, v_host)
        t1 = time.perf_counter()
        if has_eval:
            h.eval_topk(200, None, None, n_train_nz=tr.nnz, n_test_nz=te.nnz, out=eval_out)
        t2 = time.perf_counter()
        if step >= args.warmup:
            e2e_epoch += t1 - t0
            e2e_eval += t2 - t1
    e2e_epoch_ms = 1e3 * e2e_epoch / args.steps
    e2e_eval_ms = 1e3 * e2e_eval / args.steps
    factor_bytes = (m + n) * k * 8
    eval_h2d = 0        # candidate lists are per-fold state (hyprec_set_candidates), like the train CSR
    eval_d2h = m * 200 * (4 + 8 + 8) + m * 16 + tr.nnz + te.nnz

    # ---- metrics of the final model (reported, not timed)
    from hyprec_b200 import metrics as hm
    recall200 = None
    if has_eval:
        test_likes = numpy.asarray(te.sum(axis=1)).ravel()
        recall200 = float(hm.recall_at_x(200, out["topk_hit"], test_likes))

    # ---- roofline of the dominant kernels
    elem = 8 if args.precision == "fp64" else 4
    work = algorithmic_work(wl, elem)
    solve_ms = per_step["solve"]
    lowrank = os.environ.get("HYPREC_LOWRANK", "1") != "0"
    peak = h.probe_fma_tflops(precision)
    solve_flops = work["lowrank_flops_per_epoch"] if lowrank else work["solve_flops_per_epoch"]
    achieved = solve_flops / (solve_ms * 1e-3) / 1e12
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback"
    if os.path.exists(peaks_file):
        hbm_peak, hbm_src = json.load(open(peaks_file))["hbm_gbs"], "measured"
    solve_gbs = work["solve_bytes_per_epoch"](elem) / (solve_ms * 1e-3) / 1e9
    sweep_tflops = work["eval_flops"] / (max(per_step["count"], 1e-9) * 1e-3) / 1e12
    tf32_peak = None
    if os.path.exists(peaks_file):
        tf32_peak = json.load(open(peaks_file))["bf16_tflops"] / 2.0
    # pre-filtered top-k: one 32-byte sector of s~ per candidate + ~200 exact re-scores per user
    topk_bytes = (float(wl["cand_ids"].shape[0]) if has_eval else 0.0) * 36.0 + m * 200 * (k * elem + 20)
    topk_gbs = topk_bytes / (max(per_step["topk"], 1e-9) * 1e-3) / 1e9
    tensor_rows = args.precision == "fp32" and lowrank and os.environ.get("HYPREC_TC_SOLVE", "1") != "0" and \
        os.environ.get("HYPREC_TC", "1") != "0" and k % 4 == 0
    if tensor_rows:
        # float32: the row systems are built and factorised in tensor memory (tc_chol*.cuh): tcgen05 does
        # the rank-k accumulation and the trailing updates (3 TF32 products per algorithmic one)
        kernel_desc = ("tc_chol_rows_kernel + tc_chol_multi_kernel (row systems of one epoch: normal matrix accumulated "
                       "by tcgen05 3xTF32 into tensor memory, 32-column panel Cholesky with tcgen05 trailing updates; "
                       "rows with more positives than n_factors as k x k systems, the rest as deg x deg low-rank systems, "
                       "small ones several per CTA)")
        bound, peak_used, peak_src = "tensor", tf32_peak, "MEASURED_PEAKS.json bf16_tflops / 2 (dense TF32 rate)"
    else:
        kernel_desc = "als_solve_rows_kernel (row systems of one epoch; %s)" % (
            "deg x deg low-rank systems in concurrent degree buckets + direct k x k for heavy rows"
            if lowrank else "direct k x k systems, 2 launches")
        bound, peak_used, peak_src = "%s_fma" % args.precision, peak, "hyprec_probe_fma_tflops on this GPU (FMA chains, best of 4)"
    # DRAM bytes of the row-solve launches of one epoch, from the ncu capture of this workload in this mode
    traffic, traffic_src = None, None
    if wl["name"] == "citeulike-a" and args.precision == "fp64" and lowrank:
        traffic = 118.7e6
        traffic_src = ("ncu dram__bytes_read.sum + dram__bytes_write.sum over the 12 als_solve_rows_kernel launches of one "
                       "epoch, profiles/r01_solve_f64_lowrank_traffic.md (algorithmic: hbm.algorithmic_bytes; the gathers "
                       "are served by L2)")
    roofline = dict(kernel=kernel_desc, bound=bound, achieved=achieved, peak=peak_used, unit="TFLOP/s",
                    frac=achieved / peak_used if peak_used else None, peak_source=peak_src, traffic=traffic,
                    traffic_source=traffic_src,
                    simt_fma_peak=peak, simt_fma_frac=achieved / peak if peak else None,
                    flops_per_epoch=solve_flops, flops_model="low-rank (sum d^2 k + d^3/3 + 4 d k)" if lowrank else
                    "SURVEY 8d direct formula", direct_formula_flops=work["solve_flops_per_epoch"],
                    hbm=dict(achieved=solve_gbs, peak=hbm_peak, unit="GB/s", frac=solve_gbs / hbm_peak,
                             peak_source=hbm_src + " MEASURED_PEAKS.json",
                             algorithmic_bytes=work["solve_bytes_per_epoch"](elem)),
                    other_kernels=None if not has_eval else dict(
                        tc_score_sweep=dict(
                            kernel="tc::tc_score_kernel (tcgen05 kind::tf32, 3xTF32 split; count class also holds the "
                                   "split / norm / re-check helpers, ~0.06 ms)",
                            bound="tensor", achieved=3.0 * sweep_tflops, peak=tf32_peak, unit="TFLOP/s",
                            frac=(3.0 * sweep_tflops / tf32_peak) if tf32_peak else None,
                            peak_source="MEASURED_PEAKS.json bf16_tflops / 2 (dense TF32 runs at half the bf16 rate)",
                            flops=3.0 * work["eval_flops"], traffic=368.4e6,
                            traffic_source="ncu --set full dram__bytes_read+write, profiles/r01_step_v3_summary.md "
                                           "(algorithmic: 36 MB factors in + 378 MB float32 scores out)"),
                        eval_topk=dict(bound="l2/hbm sector gathers of the float32 score rows + ~200 factor rows per user",
                                       achieved=topk_gbs, peak=hbm_peak, unit="GB/s", frac=topk_gbs / hbm_peak,
                                       bytes=topk_bytes, traffic=539.1e6,
                                       traffic_source="ncu --set full, profiles/r01_step_v3_summary.md")))

    # ---- CPU baseline on this box's host cores (bounded sample)
    literal = wl["train"] is not None      # the literal arithmetic needs the dense train matrix
    cpu_ms, cpu_detail = cpu_epoch_sample(wl, args.cpu_rows, literal=literal)
    cpu_vec_ms, _ = cpu_epoch_sample(wl, 8 * args.cpu_rows, literal=False)
    cpu_eval = cpu_eval_sample(wl, 16) if has_eval else None
    cpu_baseline = dict(value=cpu_ms, unit="ms", cores=os.cpu_count(), kind="port",
                        sample="%d users + %d items at full width through the reference's per-row dense arithmetic "
                               "(oracle.als.als_step_literal), extrapolated to %d + %d rows, + RMSE" %
                               (cpu_detail["users"], cpu_detail["items"], m, n),
                        vectorised_value=cpu_vec_ms, eval_users_per_s=cpu_eval)

    line = dict(metric="als_epoch_ms", value=epoch_ms, unit="ms", n_gpus=1, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * (wall_epoch + wall_eval) / args.steps, higher_is_better=False, scaling="strong",
                vs_baseline=None, dtype="f64" if args.precision == "fp64" else "f32", data="synthetic",
                config=config_block(wl, args), clocks=clocks,
                e2e=dict(value=e2e_epoch_ms, unit="ms", h2d_bytes_per_step=factor_bytes + eval_h2d,
                         d2h_bytes_per_step=factor_bytes + (eval_d2h if has_eval else 0), eval_ms=e2e_eval_ms,
                         eval_users_per_s=(m / (e2e_eval_ms * 1e-3)) if has_eval else None),
                gpu_launches=int(launches), roofline=roofline, cpu_baseline=cpu_baseline,
                eval_ms=eval_ms, eval_users_per_s=(m / (eval_ms * 1e-3)) if has_eval else None, wall_epoch_ms=1e3 * wall_epoch / args.steps,
                wall_eval_ms=1e3 * wall_eval / args.steps, kernel_ms_per_step=per_step,
                kernel_ms_note="event-timed regions of the library's main stream; the row-solve launches of a half-step run "
                               "on their own streams at the same time, so 'prep' (whitening) and the tail of 'solve' "
                               "include waiting for SMs those launches occupy -- the launch lists under profiles/ have "
                               "the per-kernel times",
                rmse=rmse,
                recall_at_200=recall200, setup_s=wl["setup_s"])
    # recall@{50..300} of the final model (BASELINE.json configs[1]), from one more untimed pass that keeps the top
    # 300 of every candidate list; computed last and guarded: nothing above depends on it
    if has_eval:
        try:
            top300 = h.eval_topk(300, None, None, want_counts=False)
            line["recall_at"] = {str(x): float(hm.recall_at_x(x, top300["topk_hit"], test_likes))
                                 for x in (50, 100, 150, 200, 250, 300)}
        except Exception as exc:
            line["recall_at"] = dict(error=repr(exc))
    # Evaluation at the scaled size (no per-user lists exist there): device-generated candidate lists over a block of
    # users.  Off by default (--eval-users 0); reported beside the line's metric, never inside it.
    if args.eval_users > 0 and not has_eval:
        try:
            block = int(min(args.eval_users, m))
            n_tr, n_te = int(tr.indptr[block]), int(te.indptr[block])
            h.set_hashed_candidates(0, 5, 20170303)
            h.eval_topk(200, None, None, 0, block, n_train_nz=n_tr, n_test_nz=n_te)          # warm-up, builds the lists
            h.reset_kernel_ms()
            t0 = time.perf_counter()
            h.eval_topk(200, None, None, 0, block, n_train_nz=n_tr, n_test_nz=n_te)
            seconds = time.perf_counter() - t0
            line["hashed_eval"] = dict(users=block, seconds=seconds, users_per_s=block / seconds,
                                       kernel_ms={name: h.kernel_ms(t)[0] for name, t in classes.items()
                                                  if name in ("count", "topk", "flags")},
                                       candidates="test positives + hashed 1/5 share of the unrated items (fold 0)")
            h.set_hashed_candidates(0, 0)
        except Exception as exc:
            line["hashed_eval"] = dict(error=repr(exc))
    print(json.dumps(line))
    try:
        h.close()
    except Exception:
        pass


def sharded_roofline(wl, precision, world, solve_ms):
    """Roofline view of rank 0's row-solve kernels in a sharded epoch: its cost-balanced 1/world share of the
    epoch's algorithmic FLOPs over the device time of its solve launches (pure arithmetic: testable on CPU)."""
    elem = 8 if precision == "fp64" else 4
    work = algorithmic_work(wl, elem)
    flops = work["lowrank_flops_per_epoch"] / float(world)
    achieved = flops / (max(solve_ms, 1e-9) * 1e-3) / 1e12
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    tf32_peak = json.load(open(peaks_file))["bf16_tflops"] / 2.0 if os.path.exists(peaks_file) else None
    tensor_rows = precision == "fp32" and wl["k"] % 4 == 0 and os.environ.get("HYPREC_TC", "1") != "0"
    if tensor_rows and tf32_peak:
        bound, peak, src = "tensor", tf32_peak, "MEASURED_PEAKS.json bf16_tflops / 2 (dense TF32 rate)"
    else:
        # float64 SIMT row kernel: nominal B200 FP64 FMA rate (the single-GPU line measures it with a probe)
        bound, peak, src = "%s_fma" % precision, (37.0 if precision == "fp64" else 74.0), "nominal B200 FMA rate"
    return dict(kernel="row-solve kernels of rank 0 (its share of the sharded epoch)", bound=bound, achieved=achieved,
                peak=peak, unit="TFLOP/s", frac=achieved / peak, peak_source=src, traffic=None,
                flops_per_epoch_rank=flops, flops_model="low-rank (sum d^2 k + d^3/3 + 4 d k) / world")


def our_arm_distributed(args, wl, rank, world, local):
    from hyprec_b200 import distributed
    sampler = ClockSampler(local) if rank == 0 else None
    clocks = {}

    def begin():
        if sampler is not None:
            sampler.start()
            time.sleep(0.3)

    def end():
        if sampler is not None:
            clocks.update(sampler.stop())

    line = distributed.bench_sharded(args, wl, rank, world, local, config_block(wl, args), dict(begin=begin, end=end))
    if rank == 0:
        line["clocks"] = clocks
        try:
            line["roofline"] = sharded_roofline(wl, args.precision, world, line["rank0_solve_kernel_ms"])
        except Exception as exc:       # the line itself must not be lost over its annotation
            line["roofline"] = dict(error=repr(exc))
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="citeulike-a", choices=sorted(WORKLOADS) + ["scaled"])
    ap.add_argument("--scale", type=float, default=0.1, help="fraction of the 1M x 2M / 2e8 config (--workload scaled)")
    ap.add_argument("--precision", default=os.environ.get("HYPREC_PRECISION"), choices=["fp64", "fp32"],
                    help="default: fp64 (citeulike workloads), fp32 (scaled workload, as BASELINE configs[3] specifies)")
    ap.add_argument("--cpu-rows", type=int, default=16, help="rows per side in the CPU baseline sample")
    ap.add_argument("--eval-users", type=int, default=0,
                    help="scaled workload, 1 GPU: also time the evaluation of this many users with device-generated "
                         "candidate lists (0: off)")
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
        return cached_scaled_workload(args.scale)
    if args.precision is None:
        args.precision = "fp64"
    return build_workload(args.workload)


def cached_scaled_workload(scale):
    """The scaled input is deterministic; local rank 0 builds it once and the other ranks (and later
    runs on the same box) read it back from /dev/shm."""
    import pickle
    path = "/dev/shm/hyprec_scaled_%g.pkl" % scale
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not os.path.exists(path):
        if local == 0:
            wl = build_scaled_workload(scale)
            with open(path + ".tmp", "wb") as f:
                pickle.dump(wl, f, protocol=4)
            os.replace(path + ".tmp", path)
            return wl
        while not os.path.exists(path):
            time.sleep(0.5)
    with open(path, "rb") as f:
        return pickle.load(f)


if __name__ == "__main__":
    main()
