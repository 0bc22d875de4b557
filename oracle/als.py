"""
TEST INFRASTRUCTURE -- CPU restatement of the weighted-ALS fit (see oracle/__init__.py).

Two restatements of /root/reference/lib/collaborative_filtering.py:69-99 (``als_step``):

* ``als_step_literal``   -- the reference arithmetic term by term (dense confidence vector,
  ``(Y.T * c).dot(Y)``, ``numpy.linalg.solve``); O(rows * n * k^2), small shapes only.
* ``als_half_step``      -- the algebraically identical sparse form
  ``A_r = 0.1 * Y^T Y + 0.9 * sum_{j in R_r} y_j y_j^T + lambda * I``,
  ``b_r = sum_{j in R_r} r_rj y_j (+ lambda * theta_r)`` solved in batches.  SURVEY.md section 7
  measured 4e-15 agreement with the literal path; tests/test_oracle_pin.py re-checks it
  against the unmodified reference class.

Confidence is fixed at 1.0 on train non-zeros and 0.1 elsewhere
(collaborative_filtering.py:135-142); lambda is not scaled by the row degree (:81, :85).
"""
import numpy
import scipy.sparse as sp

CONF_POS = 1.0   # collaborative_filtering.py:139
CONF_NEG = 0.1   # collaborative_filtering.py:135


def to_csr(train):
    """Dense or sparse train matrix -> canonical CSR (float64 data, sorted int32 indices)."""
    csr = sp.csr_matrix(train, dtype=numpy.float64)
    csr.eliminate_zeros()
    csr.sort_indices()
    return csr


def als_step_literal(latent, fixed, train_dense, lam, side="user", prior=None):
    """collaborative_filtering.py:79-98, one row at a time, dense arithmetic. In place."""
    k = latent.shape[1]
    lam_eye = numpy.eye(k) * lam
    for r in range(latent.shape[0]):
        row = train_dense[r, :] if side == "user" else train_dense[:, r]
        conf = numpy.where(row != 0, CONF_POS, CONF_NEG)          # build_confidence_matrix :120-144
        yty = (fixed.T * conf).dot(fixed)                           # :84 / :92
        rhs = (row * conf).dot(fixed)                               # :86 / :98
        if side == "item" and prior is not None:
            rhs = rhs + prior[r, :] * lam                           # :93-96
        latent[r, :] = numpy.linalg.solve(yty + lam_eye, rhs)
    return latent


def mix64(x):
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic) -- hyprec_b200/csrc/comm.cuh mix64."""
    x = numpy.asarray(x, dtype=numpy.uint64)
    with numpy.errstate(over="ignore"):
        x = x + numpy.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> numpy.uint64(30))) * numpy.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> numpy.uint64(27))) * numpy.uint64(0x94D049BB133111EB)
    return x ^ (x >> numpy.uint64(31))


def uniform_factors(n_rows, k, seed, rows=None):
    """The initial factors hyprec_init_factors_uniform draws on the device (include/hyprec_b200.h), standing for
    numpy.random.random((rows, k)) of collaborative_filtering.py:173-176 at sizes whose factors never visit the
    host: element e = row * k + c is float32(mix64(seed * 0xD1342543DE82EF95 + e) >> 40) / 2^24.
    ``rows``: only these rows (float64 array len(rows) x k)."""
    row_ids = numpy.arange(n_rows, dtype=numpy.uint64) if rows is None else numpy.asarray(rows, dtype=numpy.uint64)
    with numpy.errstate(over="ignore"):
        e = row_ids[:, None] * numpy.uint64(k) + numpy.arange(k, dtype=numpy.uint64)[None, :]
        base = numpy.uint64(seed % (1 << 64)) * numpy.uint64(0xD1342543DE82EF95)
        r = mix64(base + e)
    return ((r >> numpy.uint64(40)).astype(numpy.float32) * numpy.float32(1.0 / 16777216.0)).astype(numpy.float64)


def als_half_step(latent, fixed, csr, lam, prior=None, batch=256, rows=None, gram=None):
    """Sparse restatement of one half-step; ``csr`` has one row per row of ``latent``.

    :param ndarray latent: rows x k float64, overwritten in place and returned (:85, :99).
    :param ndarray fixed: the other side's factors.
    :param csr: scipy CSR, shape (rows, fixed.shape[0]) -- train for users, train.T for items.
    :param prior: optional rows x k theta (update_with_items, :93-96).
    :param rows: optional iterable of row ids to solve (others untouched) -- CPU-baseline sampling.
    """
    k = latent.shape[1]
    if gram is None:                 # (``gram``: a precomputed fixed.T.dot(fixed), for row-subsample checks)
        gram = fixed.T.dot(fixed)
    base = CONF_NEG * gram + lam * numpy.eye(k)
    scale = CONF_POS - CONF_NEG
    indptr, indices, data = csr.indptr, csr.indices, csr.data
    row_ids = numpy.arange(latent.shape[0]) if rows is None else numpy.asarray(rows)
    for s in range(0, len(row_ids), batch):
        chunk = row_ids[s:s + batch]
        a = numpy.empty((len(chunk), k, k))
        b = numpy.zeros((len(chunk), k))
        for t, r in enumerate(chunk):
            lo, hi = indptr[r], indptr[r + 1]
            y = fixed[indices[lo:hi]]
            a[t] = base + scale * y.T.dot(y)
            b[t] = data[lo:hi].dot(y) * CONF_POS
            if prior is not None:
                b[t] += prior[r] * lam
        latent[chunk] = numpy.linalg.solve(a, b[:, :, None])[:, :, 0]
    return latent


def rmse_dense(user_vecs, item_vecs, train_dense):
    """evaluator.py:236-252 literally: row-wise rss over ALL cells, / size, sqrt."""
    predicted = user_vecs.dot(item_vecs.T)
    rss = 0
    for i in range(predicted.shape[0]):
        rss += numpy.sum((predicted[i] - train_dense[i]) ** 2)
    return numpy.sqrt(float(rss) / numpy.size(predicted))


def rmse_gram(user_vecs, item_vecs, csr):
    """Same quantity without the m x n matrix:
    rss = <U^T U, V^T V>_F - 2 sum_nz r (u.v) + sum_nz r^2   (SURVEY.md section 8a row a13)."""
    m, n = csr.shape
    gu = user_vecs.T.dot(user_vecs)
    gv = item_vecs.T.dot(item_vecs)
    coo = csr.tocoo()
    dots = numpy.einsum("ij,ij->i", user_vecs[coo.row], item_vecs[coo.col])
    rss = numpy.sum(gu * gv) - 2.0 * numpy.dot(coo.data, dots) + numpy.dot(coo.data, coo.data)
    return numpy.sqrt(rss / (float(m) * float(n)))


def partial_train(user_vecs, item_vecs, csr, lam, n_iter, prior=None, start_error=numpy.inf,
                  dense_rmse=None):
    """collaborative_filtering.py:224-259: user half-step, item half-step, RMSE, early break
    when ``error >= old_error`` (:256).  Returns (epochs_run, [rmse per epoch])."""
    csc = csr.T.tocsr()
    csc.sort_indices()
    error = start_error
    history = []
    for _ in range(n_iter):
        old_error = error
        als_half_step(user_vecs, item_vecs, csr, lam)
        als_half_step(item_vecs, user_vecs, csc, lam, prior=prior)
        if dense_rmse is not None:
            error = rmse_dense(user_vecs, item_vecs, dense_rmse)
        else:
            error = rmse_gram(user_vecs, item_vecs, csr)
        history.append(error)
        if error >= old_error:
            break
    return len(history), history
