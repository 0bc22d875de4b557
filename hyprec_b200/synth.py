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


def scaled_ratings(scale, device=None, seed=20170303, train_share=0.8):
    """BASELINE.json configs[3] times ``scale``: users x items = 1e6 x 2e6, 2e8 positives (log-normal user
    activity, mean 200), Zipf(0.8) item popularity under a fixed permutation, built directly as CSR; a random
    ``train_share`` of the positives is the train matrix (5-fold-equivalent split), the rest the test matrix.
    No dense matrix and no per-user candidate list exists at this size.

    The random draws are numpy's (seeded, identical on every machine); the two heavy steps -- inverse-CDF lookup
    of 2e8 samples and the sort / de-duplication of the (user, item) keys -- run through torch on ``device``
    (exact comparisons and integer sorts: the same arrays on CPU and GPU).  Returns a dict of plain arrays:
    m, n, and (indptr int64, indices int32) for ``full``, ``train``, ``test``."""
    import torch
    m, n, nnz = int(1e6 * scale), int(2e6 * scale), int(2e8 * scale)
    dev = torch.device(device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))
    rs = numpy.random.RandomState(seed)
    raw = rs.lognormal(mean=numpy.log(200.0) - 0.32, sigma=0.8, size=m)
    deg = numpy.clip(numpy.round(raw * (nnz / raw.sum())), 1, n // 4).astype(numpy.int64)
    p = (numpy.arange(n) + 20.0) ** -0.8
    cdf = numpy.cumsum(p[rs.permutation(n)])
    cdf /= cdf[-1]
    total = int(deg.sum())
    draws = rs.random_sample(total)
    with torch.no_grad():
        rows = torch.repeat_interleave(torch.arange(m, dtype=torch.int64, device=dev), torch.from_numpy(deg).to(dev))
        items = torch.searchsorted(torch.from_numpy(cdf).to(dev), torch.from_numpy(draws).to(dev))
        del draws
        items.clamp_(max=n - 1)
        keys = torch.unique(rows * n + items)           # sorted; drops the duplicates of sampling with replacement
        del rows, items
        n_keys = int(keys.shape[0])
        keep = torch.from_numpy(rs.random_sample(n_keys) < train_share).to(dev)
        out = dict(m=m, n=n)
        for name, mask in (("full", None), ("train", keep), ("test", ~keep)):
            sel = keys if mask is None else keys[mask]
            r = torch.div(sel, n, rounding_mode="floor")
            indptr = torch.zeros(m + 1, dtype=torch.int64, device=dev)
            indptr[1:] = torch.cumsum(torch.bincount(r, minlength=m), 0)
            out[name] = (indptr.cpu().numpy(), (sel - r * n).to(torch.int32).cpu().numpy())
            del sel, r, indptr
        del keys, keep
    if dev.type == "cuda":
        torch.cuda.empty_cache()
    return out
