"""
Hybrid blend of the content-based and the collaborative scores.

Mirrors /root/reference/lib/linear_regression.py:10-69: an ordinary least squares fit (with intercept,
like ``sklearn.linear_model.LinearRegression()``) of the flattened train labels on the two flattened score
matrices; ``train()`` returns ``coef[0] * item_based + coef[1] * collaborative`` -- the intercept is fitted
and then dropped, exactly as the reference does (:64-69).  The fit needs only the 2 x 2 covariance of the
two features and their covariances with the labels, so no m*n x 2 design matrix is built and scikit-learn
is not needed.
"""
import numpy


class LinearRegression(object):
    """Linear regression to combine the results of two matrices (same interface as the reference)."""

    def __init__(self, train_labels, test_labels, item_based_ratings, collaborative_ratings):
        self.item_based_ratings = item_based_ratings
        self.collaborative_ratings = collaborative_ratings
        self.item_based_ratings_shape = item_based_ratings.shape
        self.collaborative_ratings_shape = collaborative_ratings.shape
        self.flat_train_labels = self.flatten_matrix(train_labels)
        self.flat_test_labels = self.flatten_matrix(test_labels)
        self.regression_coef1 = 0
        self.regression_coef2 = 0
        self.intercept = 0

    def flatten_matrix(self, matrix):
        """linear_regression.py:35-43."""
        return numpy.asarray(matrix).flatten()

    def unflatten(self, matrix, shape):
        """linear_regression.py:45-54."""
        return matrix.reshape(shape)

    def fit(self):
        """Coefficients of y ~ c0 + c1 * item_based + c2 * collaborative (least squares)."""
        x1 = numpy.asarray(self.item_based_ratings, dtype=numpy.float64).ravel()
        x2 = numpy.asarray(self.collaborative_ratings, dtype=numpy.float64).ravel()
        y = numpy.asarray(self.flat_train_labels, dtype=numpy.float64)
        m1, m2, my = x1.mean(), x2.mean(), y.mean()
        d1, d2, dy = x1 - m1, x2 - m2, y - my
        gram = numpy.array([[d1.dot(d1), d1.dot(d2)], [d1.dot(d2), d2.dot(d2)]])
        rhs = numpy.array([d1.dot(dy), d2.dot(dy)])
        coef = numpy.linalg.lstsq(gram, rhs, rcond=None)[0]      # minimum-norm when a feature is constant
        self.regression_coef1, self.regression_coef2 = float(coef[0]), float(coef[1])
        self.intercept = float(my - coef[0] * m1 - coef[1] * m2)
        return self.regression_coef1, self.regression_coef2

    def train(self):
        """linear_regression.py:56-69: the adjusted predictions matrix (intercept not applied)."""
        c1, c2 = self.fit()
        return c2 * self.collaborative_ratings + c1 * self.item_based_ratings
