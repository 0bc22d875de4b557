"""GPU tests of the rows added after the first parity suite (file name sorts last on purpose): the
hyper-parameter sweep through the C ABI, recall at several cut-offs from one device pass, the batched
full-catalogue dump."""
import numpy
import pytest

from conftest import golden_folds, load_golden
from grid_helpers import GRID, make_recommender
from hyprec_b200.grid_search import GridSearch
from test_grid_search import check_outcome

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden():
    return load_golden("grid_naive_toy")


@pytest.mark.parametrize("workers", [1, 3])
def test_device_sweep_matches_reference_rows(golden, monkeypatch, workers):
    """The real thing: every model trained and evaluated through the C ABI, several in flight."""
    from hyprec_b200 import CollaborativeFiltering
    monkeypatch.setattr(GridSearch, "dump_csv", lambda self, rows: None)
    rec = make_recommender(CollaborativeFiltering, golden["ratings"])
    search = GridSearch(rec, GRID, verbose=False, workers=workers)
    assert search._device_backed()
    try:
        best, rows = search.train()
        check_outcome(search, best, rows, golden)
        assert numpy.random.random() == float(golden["rng_after"])
        assert numpy.allclose(rec.user_vecs, golden["final_user_vecs"], rtol=1e-7, atol=1e-10)
        # the caller's recommender holds the last model and keeps working on its own handle
        report = rec.get_evaluation_report()
        assert numpy.allclose([float(x) for x in report], golden["rows"][-1][2:], rtol=1e-8)
    finally:
        rec.close()


def test_recall_at_cutoffs_from_one_pass(als_cases):
    """recall@x for several x (BASELINE configs[1]) against the host mirror's dense path, which the CPU suite
    pins to the reference (evaluator.py:268-292)."""
    from hyprec_b200 import CollaborativeFiltering, Evaluator, ModelInitializer
    g = als_cases["als_kfold_toy"]
    fold = golden_folds(g)[0]
    hyper = {'n_factors': int(g["k"]), '_lambda': float(g["lam"])}
    options = {'k_folds': int(g["k_folds"]), 'n_iterations': int(g["n_iter"])}
    ev = Evaluator(g["ratings"])
    cf = CollaborativeFiltering(ModelInitializer(hyper.copy(), options['n_iterations']), ev, hyper, options,
                                load_matrices=False, dump_matrices=False)
    lists = ev.get_kfold_indices()
    cf.set_data(*ev.get_fold(0, lists))
    cf.hyperparameters['fold'] = 0
    cf.user_vecs, cf.item_vecs = fold["u"].copy(), fold["v"].copy()
    cutoffs = (1, 3, 5, 10, 24, 300)
    try:
        cf.get_evaluation_report()
        lists_before, details_before = ev.recommendation_indices, cf.last_report_details
        got = cf.recall_at_cutoffs(cutoffs)
        assert cf.n_recommendations == 200
        # the 300-long lists were the call's own: the evaluator and the report details hold the last report's again
        assert ev.recommendation_indices is lists_before and cf.last_report_details is details_before
        assert max(len(r) for r in ev.recommendation_indices) <= 200
    finally:
        cf.close()
    host = Evaluator(g["ratings"])
    host.set_kfolds(int(g["k_folds"]))
    host.test_indices = lists
    predictions = fold["u"].dot(fold["v"].T)
    avg = numpy.cumsum(predictions, axis=1)[:, -1] / predictions.shape[1]
    rounded = (predictions >= avg[:, None]).astype(numpy.float64)
    host.load_top_recommendations(300, predictions, fold["test"], 0)
    for x in cutoffs:
        want = host.recall_at_x(x, predictions, fold["test"], rounded)
        assert got[x] == want, (x, got[x], want)


def test_batched_dump_matches_per_user_lists(als_cases, tmp_path):
    """dump_recommendations in one device pass == the per-user recommend_items flow of
    abstract_recommender.py:148-171 (host mirror of TopRecommendations on the dense scores)."""
    from hyprec_b200 import CollaborativeFiltering, Evaluator, ModelInitializer, formats
    from hyprec_b200.top_recommendations import TopRecommendations
    g = als_cases["als_kfold_toy"]
    fold = golden_folds(g)[0]
    hyper = {'n_factors': int(g["k"]), '_lambda': float(g["lam"])}
    options = {'k_folds': int(g["k_folds"]), 'n_iterations': int(g["n_iter"])}
    ev = Evaluator(g["ratings"])
    cf = CollaborativeFiltering(ModelInitializer(hyper.copy(), options['n_iterations']), ev, hyper, options,
                                load_matrices=False, dump_matrices=False)
    lists = ev.get_kfold_indices()
    cf.set_data(*ev.get_fold(0, lists))
    cf.hyperparameters['fold'] = 0
    cf.user_vecs, cf.item_vecs = fold["u"].copy(), fold["v"].copy()
    try:
        cf.get_evaluation_report()                      # leaves the fold's candidate lists on the device
        path = cf.dump_recommendations(15, path=str(tmp_path / "top.dat"))
        got = formats.read_recommendations(path)
        single = [i for i, _ in cf.recommend_items(7, 15)]
        report_again = cf.get_evaluation_report()       # candidate lists come back for the next report
    finally:
        cf.close()
    predictions = fold["u"].dot(fold["v"].T)
    m, n = predictions.shape
    assert len(got) == m
    differing = 0
    for user in range(m):
        top = TopRecommendations(15)
        for i in range(n):
            top.insert(i, predictions[user][i])
        want = [idx for idx, val in zip(top.get_indices(), top.get_values()) if val > 1e-6]
        if got[user] != want:                           # only score ties may permute (documented carve-out)
            differing += 1
            assert sorted(numpy.round(predictions[user][got[user]], 10)) == \
                sorted(numpy.round(predictions[user][want], 10))
        if user == 7:
            assert single == got[user] or differing
    assert differing <= 2
    assert got[3] == []                                 # the user without ratings: all scores 0
    assert numpy.allclose([float(x) for x in report_again], fold["report"], rtol=1e-9, equal_nan=True)


@pytest.mark.parametrize("n,use_tc,lists", [(17000, "1", "all"), (30000, "0", "all"), (17000, "1", "explicit"),
                                             (30000, "0", "explicit")])
def test_candidate_lists_beyond_shared_memory(monkeypatch, n, use_tc, lists):
    """Full catalogue / long explicit lists of a citeulike-sized item set (recommend_items and
    dump_recommendations at n = 16 980, the naive split's lists): the per-candidate arrays of the top-k kernel
    spill to a per-CTA global scratch (16 B per candidate with the tensor-core pre-filter, 8 B without)."""
    from hyprec_b200 import _native as native
    from test_gpu_parity import check_eval, make_handle, synthetic
    monkeypatch.setenv("HYPREC_TC", use_tc)
    m, k = 24, 16
    full, train, test, u0, v0, _ = synthetic(m, n, k, 3000, seed=n // 1000)
    rs = numpy.random.RandomState(n)
    cands = None if lists == "all" else [rs.randint(0, n, size=n - 7 * u) for u in range(m)]     # with repeats
    h, _ = make_handle(native, train)
    try:
        h.set_factors(u0, v0)
        h.als_half_step(native.SIDE_USER, 0.01)
        h.als_half_step(native.SIDE_ITEM, 0.01)
        u, v = h.get_factors()
        check_eval(native, h, full, train, test, u, v, cands, 200)
    finally:
        h.close()


def test_hybrid_report_from_blended_factors(als_cases, monkeypatch):
    """is_hybrid report through the device with the blended low-rank factors == the reference's dense host flow."""
    from hyprec_b200 import CollaborativeFiltering, Evaluator, ModelInitializer, synth
    g = als_cases["als_kfold_toy"]
    fold = golden_folds(g)[0]
    theta = synth.dirichlet_theta(g["ratings"].shape[1], 6, seed=11)
    hyper = {'n_factors': int(g["k"]), '_lambda': float(g["lam"])}
    options = {'k_folds': int(g["k_folds"]), 'n_iterations': 1}

    class Content(object):
        def __init__(self):
            self.document_distribution = theta
            self.predictions = None

        def set_data(self, train, test):
            self.train_data = train

        def get_predictions(self):
            from hyprec_b200 import hybrid
            p, a = hybrid.content_factors(self.train_data, self.document_distribution)
            return p.dot(a.T)

    reports = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("HYPREC_HYBRID_LOWRANK", mode)
        ev = Evaluator(g["ratings"])
        cf = CollaborativeFiltering(ModelInitializer(hyper.copy(), 1), ev, hyper, options, load_matrices=False,
                                    dump_matrices=False, is_hybrid=True)
        cf.set_item_based_recommender(Content())
        lists = ev.get_kfold_indices()
        cf.set_data(*ev.get_fold(0, lists))
        cf.hyperparameters['fold'] = 0
        cf.user_vecs, cf.item_vecs = fold["u"].copy(), fold["v"].copy()
        try:
            reports[mode] = numpy.array([float(x) for x in cf.get_evaluation_report()])
        finally:
            cf.close()
    assert numpy.allclose(reports["1"], reports["0"], rtol=1e-8, equal_nan=True), (reports["1"], reports["0"])


def test_hashed_candidates_equal_explicit_lists():
    """hyprec_set_hashed_candidates: the device-generated lists rank exactly like the same lists passed explicitly."""
    from hyprec_b200 import _native as native
    from oracle import evaluation as oev
    from test_gpu_parity import make_handle, synthetic
    m, n, k = 40, 3000, 24
    full, train, test, u0, v0, _ = synthetic(m, n, k, 5000, seed=5)
    h, _ = make_handle(native, train)
    try:
        h.set_eval_csr((full.indptr, full.indices, full.data), (test.indptr, test.indices, test.data))
        h.set_factors(u0, v0)
        h.als_half_step(native.SIDE_USER, 0.01)
        h.als_half_step(native.SIDE_ITEM, 0.01)
        for lo, hi in ((0, m), (7, 29)):
            lists = oev.hashed_candidates(full, test, range(lo, hi), 2, 5, 99)
            cp = numpy.zeros(hi - lo + 1, numpy.int64)
            cp[1:] = numpy.cumsum([len(c) for c in lists])
            want = h.eval_topk(50, cp, numpy.concatenate(lists).astype(numpy.int32), lo, hi, want_counts=False)
            h.set_hashed_candidates(2, 5, 99)
            got = h.eval_topk(50, None, None, lo, hi, want_counts=False)
            again = h.eval_topk(50, None, None, lo, hi, want_counts=False)          # cached lists
            h.set_hashed_candidates(0, 0)
            for key in ("topk_idx", "topk_val", "topk_hit"):
                assert numpy.array_equal(got[key], want[key]) and numpy.array_equal(again[key], want[key])
        everything = h.eval_topk(50, None, None, 7, 29, want_counts=False)          # switched off: all items again
        assert not numpy.array_equal(everything["topk_idx"], want["topk_idx"])
    finally:
        h.close()


@pytest.mark.parametrize("precision", [0, 1])
def test_one_byte_hits_equal_the_float64_hits(precision):
    """hyprec_eval_topk_u8: the hit values as one byte each (and no score array) are the float64 call's, for both the
    fused top-k and a reused output dict; ratings that do not fit a byte are refused, not truncated."""
    from hyprec_b200 import _native as native
    from test_gpu_parity import make_handle, synthetic
    m, n, k = 70, 900, 24
    full, train, test, u0, v0, cands = synthetic(m, n, k, 6000, seed=8)
    cp = numpy.zeros(m + 1, numpy.int64)
    cp[1:] = numpy.cumsum([len(c) for c in cands])
    ids = numpy.concatenate(cands).astype(numpy.int32)
    h, _ = make_handle(native, train, precision)
    try:
        h.set_eval_csr((full.indptr, full.indices, full.data), (test.indptr, test.indices, test.data))
        h.set_factors(u0, v0)
        h.als_half_step(native.SIDE_USER, 0.01)
        h.als_half_step(native.SIDE_ITEM, 0.01)
        want = h.eval_topk(40, cp, ids, n_train_nz=train.nnz, n_test_nz=test.nnz)
        got = h.eval_topk(40, cp, ids, n_train_nz=train.nnz, n_test_nz=test.nnz, compact=True)
        assert got["topk_val"] is None and got["topk_hit"].dtype == numpy.uint8
        assert want["topk_hit"].max() == 1.0
        for key in ("topk_idx", "thresholds", "ones_per_row", "train_flags", "test_flags"):
            assert numpy.array_equal(got[key], want[key]), key
        assert numpy.array_equal(got["topk_hit"].astype(numpy.float64), want["topk_hit"])
        got["topk_hit"][...] = 7
        again = h.eval_topk(40, cp, ids, out=got)                               # reused arrays keep the compact call
        assert again is got and numpy.array_equal(got["topk_hit"].astype(numpy.float64), want["topk_hit"])
        part = h.eval_topk(40, cp[11:31] - cp[11], ids[cp[11]:cp[30]], 11, 30, want_counts=False, compact=True)
        assert numpy.array_equal(part["topk_hit"], got["topk_hit"][11:30])
        odd = full.data.copy()
        odd[3] = 2.5
        h.set_eval_csr((full.indptr, full.indices, odd), (test.indptr, test.indices, test.data))
        with pytest.raises(native.HyprecError) as err:
            h.eval_topk(40, cp, ids, compact=True)
        assert err.value.code == -4
        exact = h.eval_topk(40, cp, ids)                                          # the float64 form still serves them
        assert exact["topk_hit"].max() in (1.0, 2.5)
    finally:
        h.close()
