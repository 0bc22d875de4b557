"""
Synthetic inputs in the shape of the reference's datasets (citeulike-a / citeulike-t,
/root/reference/config/config.json:8; no dataset files or database exist in this environment).
Generators and seeds are part of the measurement contract.
"""
import numpy
import scipy.sparse as sp

SHAPES = {
    # name: (users, items, positives, min positives per user, seed)
    "citeulike-a": (5551, 16980, 204986, 10, 20170301),
    "citeulike-t": (7947, 25975, 134860, 3, 20170302),
    "toy": (40, 120, 700, 3, 7),
    "small": (300, 900, 9000, 5, 11),
}


def user_degrees(rs, n_users, total, min_degree, max_degree):
    """Log-normal user activity (a few heavy users, many light ones) summing to ``total``."""
    raw = numpy.floor(rs.lognormal(mean=2.9, sigma=0.9, size=n_users))
    spare = total - min_degree * n_users
    deg = min_degree + numpy.floor(raw * (spare / raw.sum())).astype(numpy.int64)
    deg = numpy.minimum(deg, max_degree)
    short = total - int(deg.sum())
    if total > max_degree * n_users:
        raise ValueError("%d positives do not fit %d users of at most %d positives" % (total, n_users, max_degree))
    order = rs.permutation(n_users)
    i = 0
    while short > 0:            # hand the remainder out one by one
        u = order[i % n_users]
        if deg[u] < max_degree:
            deg[u] += 1
            short -= 1
        i += 1
    return deg


def ratings_csr(name="citeulike-a", n_users=None, n_items=None, positives=None, min_degree=None, seed=None):
    """Binary users x items CSR: Zipf-like item popularity p_i ~ (rank+20)^-0.8 under a fixed
    permutation, items drawn per user without replacement (Gumbel top-k)."""
    base = SHAPES[name]
    n_users = n_users or base[0]
    n_items = n_items or base[1]
    positives = positives or base[2]
    min_degree = base[3] if min_degree is None else min_degree
    rs = numpy.random.RandomState(base[4] if seed is None else seed)
    deg = user_degrees(rs, n_users, positives, min_degree, min(n_items // 4, 403))
    logp = -0.8 * numpy.log(numpy.arange(n_items) + 20.0)
    logp = logp[rs.permutation(n_items)]
    indptr = numpy.zeros(n_users + 1, dtype=numpy.int64)
    numpy.cumsum(deg, out=indptr[1:])
    indices = numpy.empty(indptr[-1], dtype=numpy.int32)
    block = max(1, (1 << 24) // n_items)
    for s in range(0, n_users, block):
        e = min(n_users, s + block)
        keys = logp[None, :] + rs.gumbel(size=(e - s, n_items))
        for u in range(s, e):
            d = deg[u]
            top = numpy.argpartition(keys[u - s], n_items - d)[n_items - d:]
            indices[indptr[u]:indptr[u + 1]] = numpy.sort(top)
    data = numpy.ones(len(indices), dtype=numpy.float64)
    return sp.csr_matrix((data, indices, indptr), shape=(n_users, n_items))


def dirichlet_theta(n_items, n_factors, seed=7, alpha=0.1):
    """Stand-in for the LDA document-topic matrix the hybrid configuration initialises item
    factors with (lib/LDA.py:56-70 -> collaborative_filtering.py:174-178): rows ~ Dirichlet(alpha)."""
    rs = numpy.random.RandomState(seed)
    return rs.dirichlet(numpy.full(n_factors, alpha), size=n_items)
