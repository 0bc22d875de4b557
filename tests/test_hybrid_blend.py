"""Hybrid blend: hyprec_b200.linear_regression against scikit-learn's
LinearRegression (what /root/reference/lib/linear_regression.py:56-69 calls) and, when the reference
tree is present, against the reference class itself."""
import os
import sys

import numpy
import pytest

from hyprec_b200.linear_regression import LinearRegression


def _case(seed, m=40, n=70):
    rs = numpy.random.RandomState(seed)
    train = (rs.rand(m, n) < 0.08).astype(numpy.float64)
    test = (rs.rand(m, n) < 0.02).astype(numpy.float64)
    item_based = rs.rand(m, n) * 0.3 + 0.2 * train
    collaborative = rs.randn(m, n) * 0.1 + 0.6 * train
    return train, test, item_based, collaborative


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_matches_sklearn(seed):
    linear_model = pytest.importorskip("sklearn.linear_model")
    train, test, item_based, collaborative = _case(seed)
    ours = LinearRegression(train, test, item_based, collaborative)
    blended = ours.train()
    ref = linear_model.LinearRegression()
    ref.fit(numpy.vstack((item_based.flatten(), collaborative.flatten())).T, train.flatten())
    assert numpy.allclose([ours.regression_coef1, ours.regression_coef2], ref.coef_, rtol=1e-9, atol=1e-12)
    assert abs(ours.intercept - ref.intercept_) < 1e-10
    assert numpy.allclose(blended, ref.coef_[1] * collaborative + ref.coef_[0] * item_based, rtol=1e-9, atol=1e-12)
    assert blended.shape == train.shape


def test_matches_reference_class_when_available():
    ref_root = "/root/reference"
    if not os.path.isdir(ref_root):
        pytest.skip("reference tree not present")
    pytest.importorskip("sklearn.linear_model")
    sys.path.insert(0, ref_root)
    try:
        from lib.linear_regression import LinearRegression as RefLinearRegression
    except Exception as exc:      # missing optional dependencies of the reference
        pytest.skip("reference LinearRegression not importable: %r" % (exc,))
    finally:
        sys.path.remove(ref_root)
    train, test, item_based, collaborative = _case(5)
    want = RefLinearRegression(train, test, item_based, collaborative)
    want_pred = want.train()
    got = LinearRegression(train, test, item_based, collaborative)
    got_pred = got.train()
    assert numpy.allclose(got_pred, want_pred, rtol=1e-9, atol=1e-12)
    assert abs(got.regression_coef1 - want.regression_coef1) < 1e-9
    assert abs(got.regression_coef2 - want.regression_coef2) < 1e-9


def test_constant_feature_is_handled():
    train, test, item_based, collaborative = _case(3)
    item_based[:] = 0.25            # e.g. a content model that was never trained
    blended = LinearRegression(train, test, item_based, collaborative).train()
    assert numpy.all(numpy.isfinite(blended))


# ---------------------------------------------------------------------------------------------------------
# low-rank form of the hybrid scores (hyprec_b200/hybrid.py)
# ---------------------------------------------------------------------------------------------------------
def _lowrank_case(seed, m=37, n=64, k=6, topics=5):
    rs = numpy.random.RandomState(seed)
    train = (rs.rand(m, n) < 0.1).astype(numpy.float64)
    train[3, :] = 0
    theta = rs.dirichlet(numpy.full(topics, 0.3), size=n)
    theta[7] = 1.0 / topics                 # constant row: zero after centring, never normalised (:111-113)
    return train, theta, rs.rand(m, k), rs.rand(n, k)


def _dense_content(train, theta):
    """content_based.py:104-117, literally."""
    v = theta.copy()
    for item in range(v.shape[0]):
        mean = numpy.mean(v[item])
        v[item] -= mean
        item_norm = numpy.sqrt(v[item].dot(v[item]))
        if item_norm > 1e-6:
            v[item] /= item_norm
    weighted = train.dot(v).dot(v.T)
    weights = v.dot(v.T.dot(numpy.ones((v.shape[0],))))
    with numpy.errstate(divide='ignore', invalid='ignore'):
        predictions = weighted / weights
    predictions[~numpy.isfinite(predictions)] = 0.0
    return predictions


@pytest.mark.parametrize("seed", [0, 4])
def test_lowrank_blend_equals_the_dense_blend(seed):
    from hyprec_b200 import hybrid
    train, theta, u, v = _lowrank_case(seed)
    p, a = hybrid.content_factors(train, theta)
    content = _dense_content(train, theta)
    assert numpy.allclose(p.dot(a.T), content, rtol=1e-10, atol=1e-12)
    dense = LinearRegression(train, train, content, u.dot(v.T))
    want = dense.train()
    c1, c2, intercept = hybrid.blend_coefficients(p, a, u, v, train)
    assert abs(c1 - dense.regression_coef1) < 1e-8 * max(1, abs(dense.regression_coef1))
    assert abs(c2 - dense.regression_coef2) < 1e-8 * max(1, abs(dense.regression_coef2))
    assert abs(intercept - dense.intercept) < 1e-8
    w_u, w_v, _, _ = hybrid.blended_factors(train, theta, u, v)
    assert w_u.shape == (train.shape[0], u.shape[1] + theta.shape[1]) and w_v.shape[0] == train.shape[1]
    assert numpy.allclose(w_u.dot(w_v.T), want, rtol=1e-8, atol=1e-10)


def test_lowrank_blend_zero_weight_items_score_zero():
    from hyprec_b200 import hybrid
    train, theta, u, v = _lowrank_case(2)
    theta[:] = 1.0 / theta.shape[1]          # every row constant: V^ = 0, all weights 0, content scores all 0
    p, a = hybrid.content_factors(train, theta)
    assert numpy.all(a == 0) and numpy.all(_dense_content(train, theta) == 0)
    w_u, w_v, c1, c2 = hybrid.blended_factors(train, theta, u, v)
    assert numpy.all(numpy.isfinite(w_u)) and numpy.all(numpy.isfinite(w_v))


def test_lowrank_blend_matches_reference_classes_when_available():
    from oracle import reference_loader
    if not reference_loader.reference_available():
        pytest.skip("reference tree not present")
    pytest.importorskip("sklearn.linear_model")
    import importlib
    from hyprec_b200 import hybrid
    reference_loader.load_reference()
    ContentBased = importlib.import_module("lib.content_based").ContentBased
    RefLinearRegression = importlib.import_module("lib.linear_regression").LinearRegression
    train, theta, u, v = _lowrank_case(9)
    content = ContentBased.__new__(ContentBased)              # no constructor: it needs an abstracts preprocessor
    content.predictions = None
    content.document_distribution = theta.copy()
    content.train_data = train
    content.n_items, content.n_factors = theta.shape
    with numpy.errstate(divide='ignore', invalid='ignore'):
        want_content = content.get_predictions()
    want = RefLinearRegression(train, train, want_content, u.dot(v.T)).train()
    w_u, w_v, _, _ = hybrid.blended_factors(train, theta, u, v)
    assert numpy.allclose(w_u.dot(w_v.T), want, rtol=1e-8, atol=1e-10)


def test_class_hands_blended_factors_to_the_report_by_default(monkeypatch):
    from hyprec_b200 import CollaborativeFiltering, Evaluator, ModelInitializer
    train, theta, u, v = _lowrank_case(6)
    hyper = {'n_factors': u.shape[1], '_lambda': 0.01}
    cf = CollaborativeFiltering(ModelInitializer(hyper.copy(), 1), Evaluator(train.copy()), hyper,
                                {'k_folds': 5, 'n_iterations': 1}, load_matrices=False, dump_matrices=False, is_hybrid=True)

    class Content(object):
        document_distribution = theta

    cf.set_item_based_recommender(Content())
    cf.set_data(train, numpy.zeros(train.shape))
    cf.user_vecs, cf.item_vecs = u, v
    monkeypatch.setenv("HYPREC_HYBRID_LOWRANK", "0")
    assert cf._report_factors() is None                       # switched off: the reference's dense host path
    monkeypatch.delenv("HYPREC_HYBRID_LOWRANK", raising=False)
    w_u, w_v = cf._report_factors()                           # default: blended low-rank factors for the device report
    want = LinearRegression(train, train, _dense_content(train, theta), u.dot(v.T)).train()
    assert numpy.allclose(w_u.dot(w_v.T), want, rtol=1e-8, atol=1e-10)
    assert w_u.flags.c_contiguous and w_v.flags.c_contiguous and w_u.dtype == numpy.float64
    Content.document_distribution = None
    assert cf._report_factors() is None
