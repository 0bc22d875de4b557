"""
Low-rank form of the hybrid recommender's scores (the default report path of a hybrid
recommender whose content model exposes its topic matrix, see ``CollaborativeFiltering._report_factors``).

The reference blends two dense m x n matrices: the content model's
``predictions = (R_train V^)(V^)^T / w`` (/root/reference/lib/content_based.py:104-117: V^ = the
document-topic rows, mean-centred and normalised; w = V^ (V^^T 1); non-finite entries -> 0) and the
collaborative ``U V^T``, with the two coefficients of an ordinary least squares fit of the flattened train
matrix on the two flattened score matrices (/root/reference/lib/linear_regression.py:56-69; intercept fitted,
then dropped).  Both terms are low-rank, so

    blended = c1 * P A^T + c2 * U V^T = [c1 P, c2 U] [A, V]^T,      P = R_train V^,  A = V^ / w

is a rank-(k' + k) factorisation that the device's score / top-k / RMSE kernels consume as ordinary factors, and
the normal equations of the fit need only moments that reduce to small Gram matrices:

    sum x1 x2 = <P^T U, A^T V>_F      sum x1^2 = <P^T P, A^T A>_F      sum x1 = (sum_u p_u).(sum_i a_i)
    sum x1 y  = sum_nz r (p_u . a_i)  (and the same with U, V for x2)

No m x n matrix is formed.  Host arithmetic is numpy on k x k sized objects plus one pass over the positives.
"""
import numpy
import scipy.sparse as sp


def content_factors(train, document_distribution):
    """(P, A) with content predictions = P A^T (content_based.py:104-117).

    :param train: users x items train matrix (dense or scipy sparse).
    :param ndarray document_distribution: items x topics (the content model's ``document_distribution``)."""
    v = numpy.array(document_distribution, dtype=numpy.float64)
    for item in range(v.shape[0]):                       # row by row like the reference (:108-113): same rounding
        v[item] -= numpy.mean(v[item])
        norm = numpy.sqrt(v[item].dot(v[item]))
        if norm > 1e-6:
            v[item] /= norm
    train = sp.csr_matrix(train, dtype=numpy.float64) if not sp.issparse(train) else train.tocsr().astype(numpy.float64)
    p = numpy.asarray(train.dot(v))
    weights = v.dot(v.T.dot(numpy.ones((v.shape[0],))))
    with numpy.errstate(divide='ignore', invalid='ignore'):
        a = v / weights[:, None]
    # the reference zeroes every non-finite prediction (:116): a zero weight makes the whole item column
    # non-finite (x / 0), hence a zero row here
    a[~numpy.isfinite(a).all(axis=1)] = 0.0
    return p, a


def blend_coefficients(p, a, user_vecs, item_vecs, train):
    """(c1, c2, intercept) of ``train ~ c0 + c1 * (P A^T) + c2 * (U V^T)`` over all m x n cells -- what
    ``sklearn.linear_model.LinearRegression().fit`` returns in linear_regression.py:62-68 -- from Gram-sized
    moments."""
    train = sp.csr_matrix(train, dtype=numpy.float64) if not sp.issparse(train) else train.tocsr().astype(numpy.float64)
    m, n = train.shape
    count = float(m) * float(n)
    coo = train.tocoo()
    s1 = p.sum(axis=0).dot(a.sum(axis=0))
    s2 = user_vecs.sum(axis=0).dot(item_vecs.sum(axis=0))
    sy = coo.data.sum()
    s11 = numpy.sum(p.T.dot(p) * a.T.dot(a))
    s22 = numpy.sum(user_vecs.T.dot(user_vecs) * item_vecs.T.dot(item_vecs))
    s12 = numpy.sum(p.T.dot(user_vecs) * a.T.dot(item_vecs))
    s1y = coo.data.dot(numpy.einsum("ij,ij->i", p[coo.row], a[coo.col]))
    s2y = coo.data.dot(numpy.einsum("ij,ij->i", user_vecs[coo.row], item_vecs[coo.col]))
    m1, m2, my = s1 / count, s2 / count, sy / count
    gram = numpy.array([[s11 - count * m1 * m1, s12 - count * m1 * m2],
                        [s12 - count * m1 * m2, s22 - count * m2 * m2]])
    rhs = numpy.array([s1y - count * m1 * my, s2y - count * m2 * my])
    coef = numpy.linalg.lstsq(gram, rhs, rcond=None)[0]          # minimum norm when a feature is constant
    return float(coef[0]), float(coef[1]), float(my - coef[0] * m1 - coef[1] * m2)


def blended_factors(train, document_distribution, user_vecs, item_vecs):
    """(W_u, W_v, c1, c2): ``W_u W_v^T`` equals the reference's hybrid prediction matrix
    ``c2 * U V^T + c1 * content`` (linear_regression.py:64-69)."""
    p, a = content_factors(train, document_distribution)
    c1, c2, _ = blend_coefficients(p, a, user_vecs, item_vecs, train)
    w_u = numpy.ascontiguousarray(numpy.hstack([c2 * numpy.asarray(user_vecs, dtype=numpy.float64), c1 * p]))
    w_v = numpy.ascontiguousarray(numpy.hstack([numpy.asarray(item_vecs, dtype=numpy.float64), a]))
    return w_u, w_v, c1, c2
