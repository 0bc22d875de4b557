"""The SIMT kernels (hyprec_b200/csrc/*.cuh) executed on the CPU through the pthread CUDA shim of
tests/emu/ and compared with the oracle: indexing, tiling, padding, synchronisation and the tie
rules are checked here without a GPU.  The GPU parity tests proper are tests/test_gpu_parity.py."""
import ctypes
import importlib.util
import os

import numpy
import pytest

from conftest import golden_folds, split_lists
from oracle import als as oals
from oracle import evaluation as oev

HERE = os.path.dirname(os.path.abspath(__file__))
I64, DBL = ctypes.c_int64, ctypes.c_double


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(scope="module")
def emu():
    spec = importlib.util.spec_from_file_location("emu_build", os.path.join(HERE, "emu", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return ctypes.CDLL(mod.build())


def half_step(emu, csr, fixed, latent, lam, prior=None, chunk=8, grid=2, smem=1, lo=0, hi=None, f32=False):
    dt = numpy.float32 if f32 else numpy.float64
    out = numpy.ascontiguousarray(latent, dtype=dt).copy()
    fixed = numpy.ascontiguousarray(fixed, dtype=dt)
    prior = None if prior is None else numpy.ascontiguousarray(prior, dtype=dt)
    ip, ix, vals = csr.indptr.astype(numpy.int64), csr.indices.astype(numpy.int32), csr.data.astype(dt)
    hi = out.shape[0] if hi is None else hi
    fn = emu.emu_half_step_f32 if f32 else emu.emu_half_step_f64
    lam_c = ctypes.c_float(lam) if f32 else DBL(lam)
    status = fn(_ptr(ip), _ptr(ix), _ptr(vals), _ptr(fixed), I64(fixed.shape[0]), _ptr(out), I64(lo), I64(hi),
                out.shape[1], lam_c, _ptr(prior), chunk, grid, smem)
    return out, status


@pytest.mark.parametrize("k", [1, 5, 7, 8, 9, 16, 23])
def test_solve_kernel_all_paddings(emu, k):
    """k covers: single tile, rhs row inside the last diagonal tile (k % 8 != 0), rhs row in its
    own tile (k % 8 == 0), several panels."""
    rs = numpy.random.RandomState(k)
    m, n = 6, 40
    train = (rs.rand(m, n) < 0.4).astype(numpy.float64) * rs.randint(1, 3, size=(m, n))
    train[2, :] = 0
    u0, v0 = rs.rand(m, k), rs.rand(n, k)
    want = oals.als_half_step(u0.copy(), v0, oals.to_csr(train), 0.01)
    got, status = half_step(emu, oals.to_csr(train), v0, u0, 0.01, chunk=8, grid=3)
    assert status == 0
    assert numpy.allclose(got, want, rtol=1e-10, atol=1e-12)
    assert numpy.all(got[2] == 0)          # empty row -> exact zeros (App. A Q10)


def test_solve_kernel_golden_epoch(emu, als_cases):
    g = als_cases["als_theta_toy"]
    fold = golden_folds(g)[0]
    lam, train = float(g["lam"]), fold["train"]
    csr, csc = oals.to_csr(train), oals.to_csr(train.T)
    u, _ = half_step(emu, csr, fold["v0"], fold["u0"], lam, chunk=8, grid=4)
    v, _ = half_step(emu, csc, u, fold["v0"], lam, prior=g["theta"], chunk=8, grid=4)
    u_ref = oals.als_half_step(fold["u0"].copy(), fold["v0"], csr, lam)
    v_ref = oals.als_half_step(fold["v0"].copy(), u_ref, csc, lam, prior=g["theta"])
    assert numpy.allclose(u, u_ref, rtol=1e-9, atol=1e-11)
    assert numpy.allclose(v, v_ref, rtol=1e-9, atol=1e-11)


def test_solve_kernel_shard_scratch_and_f32(emu):
    rs = numpy.random.RandomState(2)
    m, n, k = 9, 50, 12
    train = (rs.rand(m, n) < 0.5).astype(numpy.float64)
    u0, v0 = rs.rand(m, k), rs.rand(n, k)
    want = oals.als_half_step(u0.copy(), v0, oals.to_csr(train), 0.05)
    got, _ = half_step(emu, oals.to_csr(train), v0, u0, 0.05, chunk=16, grid=2, smem=0, lo=3, hi=7)
    assert numpy.allclose(got[3:7], want[3:7], rtol=1e-10, atol=1e-12)
    assert numpy.array_equal(got[:3], u0[:3]) and numpy.array_equal(got[7:], u0[7:])   # outside the shard
    got32, _ = half_step(emu, oals.to_csr(train), v0, u0, 0.05, chunk=8, grid=2, f32=True)
    assert numpy.allclose(got32, want, rtol=2e-3, atol=2e-4)


@pytest.mark.parametrize("ts", [4, 8])
@pytest.mark.parametrize("k,cap,threads,use_prior", [(12, 100, 64, False), (16, 12, 32, True), (24, 1000, 128, True),
                                                     (40, 20, 256, False)])
def test_lowrank_half_step(emu, k, cap, threads, use_prior, ts):
    """Woodbury path on whitened factors (factor of 0.1 G + lambda I, L^-1, Q = Y L^-T, deg x deg
    systems, X = Ymat L^-1) with rows above `cap` positives routed to the direct kernel."""
    rs = numpy.random.RandomState(k)
    m, n = 8, 70
    train = (rs.rand(m, n) < 0.35).astype(numpy.float64) * rs.randint(1, 3, size=(m, n))
    train[0, :] = 0
    csr = oals.to_csr(train)
    u0, v0 = rs.rand(m, k), rs.rand(n, k)
    prior = rs.rand(m, k) if use_prior else None
    want = oals.als_half_step(u0.copy(), v0, csr, 0.05, prior=prior)
    out = u0.copy()
    ip, ix = csr.indptr.astype(numpy.int64), csr.indices.astype(numpy.int32)
    status = emu.emu_half_step_lowrank_f64(_ptr(ip), _ptr(ix), _ptr(csr.data), _ptr(v0), I64(n), _ptr(out), I64(m), k,
                                           DBL(0.05), _ptr(prior), cap, threads, 3, k % 2, ts)
    assert status == 0
    assert numpy.allclose(out, want, rtol=1e-10, atol=1e-12)


def test_singular_system_sets_status(emu):
    train = numpy.zeros((3, 4))
    train[0, 1] = 1
    u0, v0 = numpy.ones((3, 2)), numpy.zeros((4, 2))      # G = 0, lambda = 0 -> singular
    _, status = half_step(emu, oals.to_csr(train), v0, u0, 0.0)
    assert status == 1


def test_gram_kernel(emu):
    rs = numpy.random.RandomState(3)
    for rows, k in ((70, 5), (130, 64), (33, 70)):
        y = rs.rand(rows, k)
        g = numpy.empty((k, k))
        emu.emu_gram_f64(_ptr(y), I64(rows), k, 3, _ptr(g))
        assert numpy.allclose(g, y.T.dot(y), rtol=1e-12)
        assert numpy.array_equal(g, g.T)


def test_rmse_terms_and_dense_prediction(emu, als_cases):
    fold = golden_folds(als_cases["als_kfold_toy"])[2]
    u, v, train = fold["u"], fold["v"], fold["train"]
    csr = oals.to_csr(train)
    out = numpy.zeros(3)
    m, n = train.shape
    emu.emu_rmse_terms_f64(_ptr(csr.indptr.astype(numpy.int64)), _ptr(csr.indices.astype(numpy.int32)),
                           _ptr(csr.data), _ptr(u), _ptr(v), I64(m), I64(n), u.shape[1], _ptr(out))
    rmse = numpy.sqrt((out[0] - 2 * out[1] + out[2]) / (m * n))
    assert abs(rmse - fold["report"][2]) < 1e-13
    pred = numpy.empty((m, n))
    emu.emu_predict_f64(_ptr(u), _ptr(v), I64(m), I64(n), u.shape[1], _ptr(pred))
    assert numpy.allclose(pred, u.dot(v.T), rtol=1e-13, atol=1e-15)


def run_eval(emu, u, v, full, cands, cap, grid=3):
    m, n = full.shape
    csr = oals.to_csr(full)
    fp, fi, fv = csr.indptr.astype(numpy.int64), csr.indices.astype(numpy.int32), csr.data.astype(numpy.float64)
    if cands is None:
        cp = ci = None
    else:
        cp = numpy.zeros(m + 1, numpy.int64)
        cp[1:] = numpy.cumsum([len(c) for c in cands])
        ci = numpy.concatenate(cands).astype(numpy.int32)
    thr, ones = numpy.zeros(m), numpy.zeros(m, numpy.int64)
    tidx = numpy.zeros((m, cap), numpy.int32)
    tval, thit = numpy.zeros((m, cap)), numpy.zeros((m, cap))
    emu.emu_eval_f64(_ptr(u), _ptr(v), I64(m), I64(n), u.shape[1], I64(0), I64(m), _ptr(cp), _ptr(ci), cap,
                     _ptr(fp), _ptr(fi), _ptr(fv), _ptr(thr), _ptr(ones), _ptr(tidx), _ptr(tval), _ptr(thit), grid)
    return thr, ones, tidx, tval, thit


@pytest.mark.parametrize("name", ["als_kfold_toy", "als_naive_toy"])
def test_eval_kernels_golden(emu, als_cases, name):
    """Top-200 lists, rounded counts and hits against the unmodified reference, including the
    all-tied zero user and (naive split) candidate lists made of duplicated ids 0/1."""
    g = als_cases[name]
    for fold in golden_folds(g)[:2]:
        cands = split_lists(fold["cand_ids"], fold["cand_len"])
        thr, ones, tidx, tval, thit = run_eval(emu, fold["u"], fold["v"], g["ratings"], cands, 200)
        want = split_lists(fold["rec_ids"], fold["rec_len"])
        for u_id, w in enumerate(want):
            assert list(tidx[u_id][:len(w)]) == list(w)
            assert numpy.all(tidx[u_id][len(w):] == -1)
        assert numpy.array_equal(ones, fold["rounded_ones"])
        ref = oev.evaluation_report(fold["u"], fold["v"], g["ratings"], fold["train"], fold["test"], cands)
        for u_id in range(len(want)):
            assert numpy.array_equal(thit[u_id][:len(want[u_id])], ref["hits"][u_id])


def test_eval_kernel_ties_and_full_catalogue(emu):
    rs = numpy.random.RandomState(9)
    m, n, k = 5, 260, 6
    u = numpy.round((rs.rand(m, k) - 0.3) * 2) / 2
    v = numpy.round((rs.rand(n, k) - 0.3) * 2) / 2           # quantised: many exactly tied scores
    full = (rs.rand(m, n) < 0.3).astype(numpy.float64)
    for cap, cands in ((7, None), (16, [rs.randint(0, n, size=150) for _ in range(m)]), (300, None)):
        _, _, tidx, _, _ = run_eval(emu, u, v, full, cands, cap)
        for u_id in range(m):
            ids = numpy.arange(n) if cands is None else cands[u_id]
            want, _ = oev.top_recommendations_stream(ids, u[u_id].dot(v[ids].T), cap)
            assert list(tidx[u_id][:len(want)]) == want


@pytest.mark.parametrize("coef", [1e-6, 3e-3, 0.5])
def test_eval_prefilter_is_exact(emu, coef):
    """Top-k through the pre-filter (approximate float32 scores perturbed up to 0.9 E, shortlist,
    exact float64 re-score) equals the streamed reference, for a tight bound, a loose bound (large
    shortlist) and a useless bound (every user falls back to exact scoring of all candidates);
    quantised factors give exactly tied scores across the cut-off."""
    rs = numpy.random.RandomState(int(coef * 1e6) % 1000)
    m, n, k, cap = 6, 1500, 6, 20
    u = numpy.round((rs.rand(m, k) - 0.3) * 4) / 4
    v = numpy.round((rs.rand(n, k) - 0.3) * 4) / 4
    u[2] = 0
    full = (rs.rand(m, n) < 0.2).astype(numpy.float64)
    csr = oals.to_csr(full)
    fp, fi, fv = csr.indptr.astype(numpy.int64), csr.indices.astype(numpy.int32), csr.data.astype(numpy.float64)
    for cands in (None, [rs.randint(0, n, size=1200) for _ in range(m)]):
        if cands is None:
            cp = ci = None
        else:
            cp = numpy.zeros(m + 1, numpy.int64)
            cp[1:] = numpy.cumsum([len(c) for c in cands])
            ci = numpy.concatenate(cands).astype(numpy.int32)
        thr, ones = numpy.zeros(m), numpy.zeros(m, numpy.int64)
        tidx = numpy.zeros((m, cap), numpy.int32)
        tval, thit = numpy.zeros((m, cap)), numpy.zeros((m, cap))
        emu.emu_eval_prefilter_f64(_ptr(u), _ptr(v), I64(m), I64(n), k, I64(0), I64(m), _ptr(cp), _ptr(ci), cap,
                                   _ptr(fp), _ptr(fi), _ptr(fv), _ptr(thr), None, _ptr(tidx), _ptr(tval), _ptr(thit),
                                   3, DBL(coef))
        for u_id in range(m):
            ids = numpy.arange(n) if cands is None else cands[u_id]
            want, vals = oev.top_recommendations_stream(ids, u[u_id].dot(v[ids].T), cap)
            assert list(tidx[u_id][:len(want)]) == want, (coef, u_id)
            assert numpy.allclose(tval[u_id][:len(want)], vals, rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("coef,spill", [(1e-6, 0), (3e-3, 0), (0.5, 0), (1e-6, 1), (1e-6, 2), (0.5, 2), (3e-3, 3)])
def test_eval_from_sweep_survivors_is_exact(emu, coef, spill):
    """The top-k without a dense score matrix: candidate bitmap, per-user cut from sampled columns
    (sample_cut_kernel), only the candidate cells with s~ >= cut - 2E handed over (in scrambled order), hash lookup
    while streaming the candidate list, exact float64 re-score.  Tight bound, loose bound, useless bound (every user
    falls back to exact scoring), full catalogue, explicit lists with repeats, a zero user (all scores tied), more
    candidates than the capacity and fewer; the spilled variant (spill & 1); and the two-launch form the library uses
    (spill & 2: pass 1 on the emitted cells alone with small arrays, pass 2 through the stream for the flagged users)."""
    rs = numpy.random.RandomState(int(coef * 1e6) % 1000 + 5)
    m, n, k, cap = 6, 3000, 6, 20
    u = numpy.round((rs.rand(m, k) - 0.3) * 4) / 4
    v = numpy.round((rs.rand(n, k) - 0.3) * 4) / 4
    u[2] = 0
    full = (rs.rand(m, n) < 0.2).astype(numpy.float64)
    csr = oals.to_csr(full)
    fp, fi, fv = csr.indptr.astype(numpy.int64), csr.indices.astype(numpy.int32), csr.data.astype(numpy.float64)
    lists = [rs.randint(0, n, size=1200) for _ in range(m)]
    lists[4] = rs.permutation(n)[:15]                      # fewer candidates than the capacity
    lists[1] = rs.permutation(n)[:1200]                    # no repeats: the emitted cells alone decide (fast attempt)
    lists[3] = rs.permutation(n)[:2500]
    for cands in (None, lists):
        if cands is None:
            cp = ci = None
        else:
            cp = numpy.zeros(m + 1, numpy.int64)
            cp[1:] = numpy.cumsum([len(c) for c in cands])
            ci = numpy.concatenate(cands).astype(numpy.int32)
        thr = numpy.zeros(m)
        tidx = numpy.zeros((m, cap), numpy.int32)
        tval, thit = numpy.zeros((m, cap)), numpy.zeros((m, cap))
        fallbacks = numpy.zeros(1, numpy.uint32)
        emu.emu_eval_survivors_f64(_ptr(u), _ptr(v), I64(m), I64(n), k, I64(0), I64(m), _ptr(cp), _ptr(ci), cap,
                                   _ptr(fp), _ptr(fi), _ptr(fv), _ptr(thr), None, _ptr(tidx), _ptr(tval), _ptr(thit),
                                   3, DBL(coef), spill, _ptr(fallbacks))
        for u_id in range(m):
            ids = numpy.arange(n) if cands is None else cands[u_id]
            want, vals = oev.top_recommendations_stream(ids, u[u_id].dot(v[ids].T), cap)
            assert list(tidx[u_id][:len(want)]) == want, (coef, u_id)
            assert numpy.allclose(tval[u_id][:len(want)], vals, rtol=1e-12, atol=1e-15)
        if coef <= 1e-5:
            assert fallbacks[0] <= 2, fallbacks        # the sample finds a usable cut for (almost) every user
        if coef >= 0.5:
            assert fallbacks[0] >= 3                   # a useless error bound: everything passes, users fall back


def test_nz_flags_kernel(emu, als_cases):
    fold = golden_folds(als_cases["als_kfold_toy"])[0]
    u, v, test = fold["u"], fold["v"], fold["test"]
    csr = oals.to_csr(test)
    pred = u.dot(v.T)
    thr = numpy.array([oev.row_threshold(r) for r in pred])
    flags = numpy.zeros(csr.nnz, dtype=numpy.uint8)
    emu.emu_nz_flags_f64(_ptr(csr.indptr.astype(numpy.int64)), _ptr(csr.indices.astype(numpy.int32)), _ptr(u), _ptr(v),
                         u.shape[1], _ptr(thr), I64(0), I64(u.shape[0]), _ptr(flags))
    rows = numpy.repeat(numpy.arange(u.shape[0]), numpy.diff(csr.indptr))
    assert numpy.array_equal(flags, (pred[rows, csr.indices] >= thr[rows]).astype(numpy.uint8))


@pytest.mark.parametrize("coef", [0.0, 1e-6, 0.5])
def test_eval_kernel_spilled_candidate_arrays(emu, als_cases, coef):
    """Candidate lists that do not fit shared memory keep their per-candidate arrays in a per-CTA global
    scratch (TopkArgs::spill): same lists as the streamed reference with exact scoring (coef 0), through
    the pre-filter with a tight bound, and with a useless bound (fallback to exact scoring) -- tied scores
    across the cut-off, full catalogue and explicit lists, and the naive split's lists of ids 0/1."""
    rs = numpy.random.RandomState(17)
    m, n, k, cap = 6, 1500, 6, 20
    u = numpy.round((rs.rand(m, k) - 0.3) * 4) / 4
    v = numpy.round((rs.rand(n, k) - 0.3) * 4) / 4
    u[2] = 0
    full = (rs.rand(m, n) < 0.2).astype(numpy.float64)

    def run(u, v, full, cands, cap):
        m, n = full.shape
        csr = oals.to_csr(full)
        fp, fi, fv = csr.indptr.astype(numpy.int64), csr.indices.astype(numpy.int32), csr.data.astype(numpy.float64)
        if cands is None:
            cp = ci = None
        else:
            cp = numpy.zeros(m + 1, numpy.int64)
            cp[1:] = numpy.cumsum([len(c) for c in cands])
            ci = numpy.concatenate(cands).astype(numpy.int32)
        thr = numpy.zeros(m)
        tidx = numpy.zeros((m, cap), numpy.int32)
        tval, thit = numpy.zeros((m, cap)), numpy.zeros((m, cap))
        emu.emu_eval_spill_f64(_ptr(u), _ptr(v), I64(m), I64(n), u.shape[1], I64(0), I64(m), _ptr(cp), _ptr(ci), cap,
                               _ptr(fp), _ptr(fi), _ptr(fv), _ptr(thr), None, _ptr(tidx), _ptr(tval), _ptr(thit),
                               3, DBL(coef))
        return tidx, tval, thit

    for cands in (None, [rs.randint(0, n, size=1200) for _ in range(m)]):
        tidx, tval, _ = run(u, v, full, cands, cap)
        for u_id in range(m):
            ids = numpy.arange(n) if cands is None else cands[u_id]
            want, vals = oev.top_recommendations_stream(ids, u[u_id].dot(v[ids].T), cap)
            assert list(tidx[u_id][:len(want)]) == want, (coef, u_id)
            assert numpy.allclose(tval[u_id][:len(want)], vals, rtol=1e-12, atol=1e-15)
    g = als_cases["als_naive_toy"]
    fold = golden_folds(g)[0]
    cands = split_lists(fold["cand_ids"], fold["cand_len"])
    tidx, _, thit = run(fold["u"], fold["v"], g["ratings"].astype(numpy.float64), cands, 200)
    want = split_lists(fold["rec_ids"], fold["rec_len"])
    ref = oev.evaluation_report(fold["u"], fold["v"], g["ratings"], fold["train"], fold["test"], cands)
    for u_id, w in enumerate(want):
        assert list(tidx[u_id][:len(w)]) == list(w)
        assert numpy.all(tidx[u_id][len(w):] == -1)
        assert numpy.array_equal(thit[u_id][:len(w)], ref["hits"][u_id])


@pytest.mark.parametrize("coef", [0.0, 1e-6])
def test_eval_kernel_capacity_300(emu, coef):
    """Top-300 (sort size 512) of lists longer than the capacity, exact and through the pre-filter: what bench.py's
    recall@{50..300} pass and recall_at_cutoffs ask for."""
    rs = numpy.random.RandomState(23)
    m, n, k, cap = 5, 2000, 6, 300
    u = numpy.round((rs.rand(m, k) - 0.3) * 8) / 8
    v = numpy.round((rs.rand(n, k) - 0.3) * 8) / 8
    full = (rs.rand(m, n) < 0.2).astype(numpy.float64)
    csr = oals.to_csr(full)
    fp, fi, fv = csr.indptr.astype(numpy.int64), csr.indices.astype(numpy.int32), csr.data.astype(numpy.float64)
    cands = [rs.permutation(n)[:1700] for _ in range(m)]
    cp = numpy.zeros(m + 1, numpy.int64)
    cp[1:] = numpy.cumsum([len(c) for c in cands])
    ci = numpy.concatenate(cands).astype(numpy.int32)
    thr = numpy.zeros(m)
    tidx = numpy.zeros((m, cap), numpy.int32)
    tval, thit = numpy.zeros((m, cap)), numpy.zeros((m, cap))
    fn = emu.emu_eval_prefilter_f64 if coef > 0 else None
    if fn is None:
        emu.emu_eval_f64(_ptr(u), _ptr(v), I64(m), I64(n), k, I64(0), I64(m), _ptr(cp), _ptr(ci), cap,
                         _ptr(fp), _ptr(fi), _ptr(fv), _ptr(thr), None, _ptr(tidx), _ptr(tval), _ptr(thit), 3)
    else:
        fn(_ptr(u), _ptr(v), I64(m), I64(n), k, I64(0), I64(m), _ptr(cp), _ptr(ci), cap,
           _ptr(fp), _ptr(fi), _ptr(fv), _ptr(thr), None, _ptr(tidx), _ptr(tval), _ptr(thit), 3, DBL(coef))
    for u_id in range(m):
        want, vals = oev.top_recommendations_stream(cands[u_id], u[u_id].dot(v[cands[u_id]].T), cap)
        assert len(want) == cap
        assert list(tidx[u_id]) == want, (coef, u_id)
        assert numpy.allclose(tval[u_id], vals, rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("n,fold,n_folds,seed", [(700, 0, 5, 0), (1037, 3, 5, 20170303), (90, 1, 2, 7)])
def test_hashed_candidate_lists(emu, n, fold, n_folds, seed):
    """Device-generated candidate lists (count pass, host prefix sum, fill pass) against the numpy restatement
    of the rule: test positives + the hashed share of the unrated items, ascending; a user range that does not
    start at 0, a user without ratings, fewer items than threads."""
    rs = numpy.random.RandomState(n)
    m = 9
    full = (rs.rand(m, n) < 0.15).astype(numpy.float64)
    full[4, :] = 0
    test = numpy.where(rs.rand(m, n) < 0.3, full, 0.0)
    fcsr, tcsr = oals.to_csr(full), oals.to_csr(test)
    fp, fi = fcsr.indptr.astype(numpy.int64), fcsr.indices.astype(numpy.int32)
    tp, ti = tcsr.indptr.astype(numpy.int64), tcsr.indices.astype(numpy.int32)
    lo, hi = 2, m
    U64 = ctypes.c_uint64
    counts = numpy.zeros(hi - lo, dtype=numpy.int32)
    emu.emu_hashed_candidates(_ptr(fp), _ptr(fi), _ptr(tp), _ptr(ti), I64(lo), I64(hi), I64(n), fold, n_folds, U64(seed),
                              _ptr(counts), None, None, 0, 3)
    want = oev.hashed_candidates(full, test, range(lo, hi), fold, n_folds, seed)
    assert list(counts) == [len(w) for w in want]
    indptr = numpy.zeros(hi - lo + 1, dtype=numpy.int64)
    indptr[1:] = numpy.cumsum(counts)
    ids = numpy.full(int(indptr[-1]), -1, dtype=numpy.int32)
    emu.emu_hashed_candidates(_ptr(fp), _ptr(fi), _ptr(tp), _ptr(ti), I64(lo), I64(hi), I64(n), fold, n_folds, U64(seed),
                              None, _ptr(indptr), _ptr(ids), 1, 3)
    for j, w in enumerate(want):
        assert numpy.array_equal(ids[indptr[j]:indptr[j + 1]], w)
    # the lists are what the rule says: every test positive, no train positive, about 1 / n_folds of the rest
    for j, u in enumerate(range(lo, hi)):
        got = set(ids[indptr[j]:indptr[j + 1]].tolist())
        assert set(numpy.nonzero(test[u])[0].tolist()) <= got
        assert not (set(numpy.nonzero(full[u] * (test[u] == 0))[0].tolist()) & got)
    share = sum(len(w) for w in want) / float((hi - lo) * n)
    assert abs(share - 1.0 / n_folds) < 0.08


def test_exchange_protocol_three_ranks(emu):
    """csrc/comm.cuh on the CPU: three ranks as three host threads over host-memory arenas -- rows pushed into every
    replica (push_rows_kernel), mailbox slots (push_slot_kernel), signal / wait barriers with sequence numbers,
    rank-ordered reduction (reduce_slots_kernel).  Every rank ends with the same replica and the same sums; a rank
    that never arrives makes the others time out (status 2) instead of hanging."""
    import threading
    world, k, rows, slot_doubles, rounds = 3, 5, 41, 37, 6
    bounds = numpy.array([0, 9, 9 + 17, rows], dtype=numpy.int64)
    off_matrix, off_slots = 0, 8 * rows * k
    off_flags = off_slots + 8 * slot_doubles * world
    size = off_flags + 8 * 8

    def run(active, timeout_ns):
        arenas = [numpy.zeros(size, dtype=numpy.uint8) for _ in range(world)]
        bases = (ctypes.c_void_p * world)(*[a.ctypes.data for a in arenas])
        sums = [numpy.zeros((rounds, slot_doubles)) for _ in range(world)]
        status = [None] * world
        fn = emu.emu_comm_rounds
        fn.restype = ctypes.c_int

        def work(r):
            status[r] = fn(bases, world, r, ctypes.c_uint64(off_matrix), ctypes.c_uint64(off_slots), ctypes.c_uint64(slot_doubles),
                           ctypes.c_uint64(off_flags), k, _ptr(bounds), rounds, _ptr(sums[r]), ctypes.c_uint64(timeout_ns))

        threads = [threading.Thread(target=work, args=(r,)) for r in active]
        for t in threads:
            t.start()
        for t in threads:
            t.join(120)
        return arenas, sums, status

    arenas, sums, status = run(range(world), 20 * 10 ** 9)
    assert status == [0, 0, 0]
    want = numpy.empty((rows, k))
    for r in range(world):
        for row in range(bounds[r], bounds[r + 1]):
            want[row] = 1000.0 * (rounds - 1) + 10.0 * row + numpy.arange(k) + 0.5 * r
    for r in range(world):
        replica = arenas[r][off_matrix:off_matrix + 8 * rows * k].view(numpy.float64).reshape(rows, k)
        assert numpy.array_equal(replica, want)
        assert numpy.array_equal(sums[r], sums[0])
    i = numpy.arange(slot_doubles) + 1.0
    for it in range(rounds):
        assert numpy.array_equal(sums[0][it], sum((r + 1) * i + it for r in range(world)))
    # rank 2 never shows up: ranks 0 and 1 give up after the timeout and report it
    _, _, status = run([0, 1], 2 * 10 ** 8)
    assert status[0] == 2 and status[1] == 2 and status[2] is None


@pytest.mark.parametrize("ts", [4, 8])
@pytest.mark.parametrize("k,cap,threads", [(12, 100, 64), (16, 12, 32), (40, 20, 256)])
def test_rmse_cross_term_from_the_row_solve(emu, k, cap, threads, ts):
    """SolveArgs::cross_out: every row's sum_j r_j (y_j . x) -- its share of the RMSE cross term (evaluator.py:249) --
    falls out of the solve: b.x = |L^-1 b|^2 for k x k rows, (|r|^2 - |z|^2) / 0.9 for the deg x deg rows; rows
    without positives contribute 0.  Non-binary ratings, both tile sizes, both paths in one half-step."""
    rs = numpy.random.RandomState(100 + k)
    m, n = 9, 70
    train = (rs.rand(m, n) < 0.35).astype(numpy.float64) * rs.randint(1, 3, size=(m, n))
    train[0, :] = 0
    csr = oals.to_csr(train)
    u0, v0 = rs.rand(m, k), rs.rand(n, k)
    want = oals.als_half_step(u0.copy(), v0, csr, 0.05)
    out = u0.copy()
    cross = numpy.zeros(m)
    ip, ix = csr.indptr.astype(numpy.int64), csr.indices.astype(numpy.int32)
    emu.emu_set_cross_out(_ptr(cross))
    try:
        status = emu.emu_half_step_lowrank_f64(_ptr(ip), _ptr(ix), _ptr(csr.data), _ptr(v0), I64(n), _ptr(out), I64(m), k,
                                               DBL(0.05), None, cap, threads, 3, k % 2, ts)
    finally:
        emu.emu_set_cross_out(None)
    assert status == 0 and numpy.allclose(out, want, rtol=1e-10, atol=1e-12)
    expect = numpy.array([(train[r] * want[r].dot(v0.T)).sum() for r in range(m)])
    assert numpy.allclose(cross, expect, rtol=1e-10, atol=1e-12), (cross, expect)
    assert cross[0] == 0.0


@pytest.mark.parametrize("k,cap", [(12, 32), (40, 9), (200, 32), (30, 20), (15, 32)])
def test_small_systems_one_warp_per_row_f64(emu, k, cap):
    """small_f64.cuh: deg x deg systems of <= 32 positives, one warp per row -- S built with (emulated) FP64 MMA
    fragments, in-warp Cholesky, shuffle forward / shared-memory backward substitution, y = Q_r^T t -- against the
    oracle's half-step, with the RMSE cross term as a by-product; rows above `cap` go through the k x k kernel,
    degrees 0, 1, 8, 9, 31, 32 included, non-binary ratings, k not a multiple of 8."""
    rs = numpy.random.RandomState(7 + k)
    m, n = 14, 90
    train = numpy.zeros((m, n))
    degrees = [0, 1, 8, 9, 31, 32, 5, 17, 24, 3, 40, 12, 2, 30]
    for u, d in enumerate(degrees):
        train[u, rs.permutation(n)[:d]] = rs.randint(1, 3, size=d)
    csr = oals.to_csr(train)
    u0, v0 = rs.rand(m, k), rs.rand(n, k)
    want = oals.als_half_step(u0.copy(), v0, csr, 0.05)
    out = u0.copy()
    cross = numpy.zeros(m)
    ip, ix = csr.indptr.astype(numpy.int64), csr.indices.astype(numpy.int32)
    emu.emu_set_cross_out(_ptr(cross))
    try:
        status = emu.emu_half_step_lowrank_f64(_ptr(ip), _ptr(ix), _ptr(csr.data), _ptr(v0), I64(n), _ptr(out), I64(m), k,
                                               DBL(0.05), None, cap, 256, 2, 1, 0)
    finally:
        emu.emu_set_cross_out(None)
    assert status == 0
    assert numpy.allclose(out, want, rtol=1e-10, atol=1e-12)
    expect = numpy.array([(train[r] * want[r].dot(v0.T)).sum() for r in range(m)])
    assert numpy.allclose(cross, expect, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("k", [12, 30, 40, 256])
def test_small_systems_one_warp_per_row_f32_factors(emu, k):
    """The float32 instantiation of small_f64.cuh (the float32 mode's <= 32 bucket): float32 whitened factors and
    ratings in, float64 arithmetic inside, float32 y = Q_r^T (I + 0.9 Q_r Q_r^T)^-1 r out, with the row's share of
    the RMSE cross term (|r|^2 - |z|^2) / 0.9 -- against the same formulas in numpy float64 on the float32 inputs
    (tolerance: one float32 rounding of the result)."""
    rs = numpy.random.RandomState(70 + k)
    m, n = 13, 80
    degrees = [0, 1, 8, 9, 31, 32, 5, 17, 24, 3, 12, 2, 30]
    train = numpy.zeros((m, n))
    for u, d in enumerate(degrees):
        train[u, rs.permutation(n)[:d]] = rs.randint(1, 3, size=d)
    csr = oals.to_csr(train)
    q = (rs.rand(n, k) / numpy.sqrt(k)).astype(numpy.float32)
    ip, ix, vals = csr.indptr.astype(numpy.int64), csr.indices.astype(numpy.int32), csr.data.astype(numpy.float32)
    rows = numpy.arange(m, dtype=numpy.int32)[::-1].copy()          # any order: output row i belongs to list position i
    ymat = numpy.full((m, k), -7.0, dtype=numpy.float32)
    cross = numpy.zeros(m)
    status = emu.emu_small_rows_f32(_ptr(ip), _ptr(ix), _ptr(vals), _ptr(q), k, _ptr(rows), I64(m), _ptr(ymat), _ptr(cross), 2)
    assert status == 0
    for at, r in enumerate(rows):
        cols = ix[ip[r]:ip[r + 1]]
        qr = q[cols].astype(numpy.float64)
        rv = vals[ip[r]:ip[r + 1]].astype(numpy.float64)
        s = numpy.eye(len(cols)) + 0.9 * qr.dot(qr.T)
        t = numpy.linalg.solve(s, rv) if len(cols) else rv
        want = qr.T.dot(t) if len(cols) else numpy.zeros(k)
        assert numpy.allclose(ymat[at], want, rtol=2e-7, atol=1e-9), (r, numpy.abs(ymat[at] - want).max())
        if len(cols):
            z = numpy.linalg.solve(numpy.linalg.cholesky(s), rv)
            assert numpy.isclose(cross[r], (rv.dot(rv) - z.dot(z)) / 0.9, rtol=1e-10, atol=1e-12)
        else:
            assert cross[r] == 0.0


@pytest.mark.parametrize("ts,threads", [(4, 512), (4, 256), (8, 256)])
def test_lowrank_rows_of_a_hundred_positives_with_many_threads(emu, ts, threads):
    """The <= 127 bucket with other launch shapes than the default 8 x 8 tiles on 128 threads (HYPREC_BUCKET_TILE /
    HYPREC_BUCKET_THREADS, tools/ab_epoch.py): systems of ~100 unknowns on 4 x 4 tiles and up to 512 threads."""
    rs = numpy.random.RandomState(ts + threads)
    m, n, k = 3, 250, 20
    train = (rs.rand(m, n) < 0.42).astype(numpy.float64)
    csr = oals.to_csr(train)
    assert numpy.diff(csr.indptr).max() > 95
    u0, v0 = rs.rand(m, k), rs.rand(n, k)
    want = oals.als_half_step(u0.copy(), v0, csr, 0.05)
    out = u0.copy()
    ip, ix = csr.indptr.astype(numpy.int64), csr.indices.astype(numpy.int32)
    status = emu.emu_half_step_lowrank_f64(_ptr(ip), _ptr(ix), _ptr(csr.data), _ptr(v0), I64(n), _ptr(out), I64(m), k,
                                           DBL(0.05), None, 1000, threads, 2, 0, ts)
    assert status == 0
    assert numpy.allclose(out, want, rtol=1e-10, atol=1e-12)
