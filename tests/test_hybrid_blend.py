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
