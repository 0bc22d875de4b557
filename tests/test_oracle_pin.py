"""Pin the oracle (oracle/*.py) against outputs of the UNMODIFIED reference (tests/golden/*.npz,
written by tests/golden/make_golden.py) and against the reference's own known answers
(/root/reference/tests/metrics_tests.py:60-107)."""
import numpy
import pytest
import scipy.sparse as sp

from conftest import golden_folds, load_golden, split_lists
from oracle import als as oals
from oracle import evaluation as oev

CASES = ["als_kfold_4x30", "als_kfold_toy", "als_theta_toy", "als_naive_toy"]


@pytest.mark.parametrize("name", CASES)
def test_partial_train_matches_reference(als_cases, name):
    g = als_cases[name]
    lam, n_iter = float(g["lam"]), int(g["n_iter"])
    prior = g["theta"] if int(g["update_with_items"]) else None
    for fold in golden_folds(g):
        u, v = fold["u0"].copy(), fold["v0"].copy()
        csr = oals.to_csr(fold["train"])
        oals.partial_train(u, v, csr, lam, n_iter, prior=prior, dense_rmse=fold["train"])
        assert numpy.allclose(u, fold["u"], rtol=1e-8, atol=1e-10)
        assert numpy.allclose(v, fold["v"], rtol=1e-8, atol=1e-10)


def test_literal_and_sparse_half_step_agree(als_cases):
    fold = golden_folds(als_cases["als_kfold_toy"])[0]
    train, u0, v0 = fold["train"], fold["u0"], fold["v0"]
    lit = oals.als_step_literal(u0.copy(), v0, train, 0.01, "user")
    fast = oals.als_half_step(u0.copy(), v0, oals.to_csr(train), 0.01)
    assert numpy.allclose(lit, fast, rtol=1e-10, atol=1e-12)
    lit_i = oals.als_step_literal(v0.copy(), u0, train, 0.01, "item", prior=v0)
    fast_i = oals.als_half_step(v0.copy(), u0, oals.to_csr(train.T), 0.01, prior=v0)
    assert numpy.allclose(lit_i, fast_i, rtol=1e-10, atol=1e-12)


def test_rmse_identity(als_cases):
    fold = golden_folds(als_cases["als_kfold_toy"])[1]
    dense = oals.rmse_dense(fold["u"], fold["v"], fold["train"])
    gram = oals.rmse_gram(fold["u"], fold["v"], oals.to_csr(fold["train"]))
    assert abs(dense - gram) < 1e-13
    assert abs(dense - fold["report"][2]) < 1e-13


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("fast", [False, True])
def test_evaluation_report_matches_reference(als_cases, name, fast):
    g = als_cases[name]
    for fold in golden_folds(g):
        cands = split_lists(fold["cand_ids"], fold["cand_len"])
        out = oev.evaluation_report(fold["u"], fold["v"], g["ratings"], fold["train"], fold["test"], cands,
                                    fast_topk=fast)
        want = split_lists(fold["rec_ids"], fold["rec_len"])
        assert [list(w) for w in want] == out["rec_ids"]
        assert numpy.array_equal(out["ones_per_row"], fold["rounded_ones"])
        got = numpy.array([float(x) for x in out["report"]])
        ref = fold["report"]
        both_nan = numpy.isnan(got) & numpy.isnan(ref)
        assert numpy.all(both_nan | (numpy.abs(got - ref) <= 1e-12 * numpy.maximum(1, numpy.abs(ref))))


def test_top_recommendations_streams():
    g = load_golden("top_recommendations")
    for c in range(int(g["n_cases"])):
        cap, stream, want = int(g["c%d_cap" % c]), g["c%d_stream" % c], list(g["c%d_result" % c])
        ids = numpy.arange(len(stream))
        assert oev.top_recommendations_stream(ids, stream, cap)[0] == want
        assert oev.top_recommendations_fast(ids, stream, cap)[0] == want


def test_metric_known_answers():
    """tests/metrics_tests.py:62-69 (MRR) and :88-107 (nDCG), tolerance 1e-3 (:12-13)."""
    g = load_golden("metrics")
    ratings, expected = g["ratings"], g["expected"]
    recs = split_lists(g["rec_ids"], g["rec_len"])
    hits = [ratings[u][numpy.asarray(r, dtype=int)] * expected[u][numpy.asarray(r, dtype=int)] for u, r in enumerate(recs)]
    n_users = 5
    assert abs(oev.mrr_at(1, hits) - (1 / 1 + 1 / 1) / n_users) < 1e-3
    assert abs(oev.mrr_at(4, hits) - (1 / 1 + 1 / 2 + 1 / 3 + 1 / 1) / n_users) < 1e-3
    assert abs(oev.mrr_at(5, hits) - (1 / 1 + 1 / 2 + 1 / 3 + 1 / 1) / n_users) < 1e-3
    for n in (1, 4, 5):
        assert oev.mrr_at(n, hits) == numpy.float16(g["mrr_%d" % n])
        assert oev.ndcg_at(n, hits) == numpy.float16(g["ndcg_%d" % n])
    likes = ratings.sum(axis=1)
    assert oev.recall_at_x(3, hits, likes) == numpy.float16(g["recall_3"])


def test_row_threshold_is_pythons_sequential_sum():
    rs = numpy.random.RandomState(0)
    row = rs.rand(1000) - 0.4
    assert oev.row_threshold(row) == sum(row) / row.shape[0]
