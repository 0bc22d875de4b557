"""
The scalar metrics of the evaluation report, computed on the host from the integers / flags the
device produces.  Each per-user value is formed with the reference's own arithmetic and the
list is handed to the same ``numpy.mean(..., dtype=numpy.float16)`` call
(/root/reference/lib/evaluator.py:292, :315, :341), so the reported scalars are bit-identical
whenever the top-k lists and hit flags are.
"""
import numpy


def _matrix(hits):
    """(m, K) float64 array when ``hits`` is one (the device's zero-padded output), else None."""
    if isinstance(hits, numpy.ndarray) and hits.ndim == 2 and hits.dtype != object:
        return numpy.asarray(hits, dtype=numpy.float64)
    return None


def recall_at_x(x, hits, test_likes):
    """evaluator.py:281-292.  hits: (m, K) array or list of per-user arrays of rating*rounded along
    the recommendation list (padding must be 0); test_likes[u] = test[u].sum().
    The array form is evaluated for all users at once with the same per-user operations (a sum of the first x
    flags, one division) and the values reach numpy.mean in user order, so the float16 result is the loop's."""
    dense = _matrix(hits)
    if dense is not None:
        likes = numpy.asarray(test_likes, dtype=numpy.float64)
        keep = likes != 0
        sums = numpy.array([numpy.sum(row) for row in dense[keep, :x]]) if dense.shape[1] and not _exact_flags(dense) \
            else dense[keep, :x].sum(axis=1)
        recalls = sums / (numpy.minimum(float(x), likes[keep]) * 1.0)
        return numpy.mean(recalls, dtype=numpy.float16)
    recalls = [numpy.sum(hits[u][:x]) / (min(x, test_likes[u]) * 1.0)
               for u in range(len(hits)) if test_likes[u] != 0]
    return numpy.mean(recalls, dtype=numpy.float16)


def _exact_flags(dense):
    """Sums of small integers are exact in any order: then the axis sum equals the per-row numpy.sum bit for bit."""
    return bool(numpy.all(dense == numpy.round(dense)) and numpy.all(numpy.abs(dense) < 1e6))


def ndcg_at(n_recommendations, hits, lengths):
    """evaluator.py:304-315; lengths[u] = len(recommendation_indices[u])."""
    logs = numpy.log2(numpy.arange(max(n_recommendations, 1)) + 2)
    dense = _matrix(hits)
    if dense is not None and n_recommendations >= 1:
        # cumsum along the row adds in the loop's order; the zero padding behind a short list adds exact zeros
        n = min(n_recommendations, dense.shape[1])
        tops = numpy.minimum(numpy.asarray(lengths, dtype=numpy.int64), n_recommendations)
        dcg = numpy.cumsum(dense[:, :n] / logs[:n], axis=1)[:, n - 1] if n else numpy.zeros(len(dense))
        idcg_by_top = numpy.concatenate(([0.0], numpy.cumsum(1 / logs[:n_recommendations])))
        idcg = idcg_by_top[tops]
        keep = (tops > 0) & (idcg != 0)
        return numpy.mean(dcg[keep] / idcg[keep], dtype=numpy.float16)
    ndcgs = []
    for u in range(len(hits)):
        top = min(int(lengths[u]), n_recommendations)
        if top == 0:
            continue
        # cumsum accumulates sequentially, like the reference's `+=` loop
        dcg = numpy.cumsum(numpy.asarray(hits[u][:top], dtype=numpy.float64) / logs[:top])[-1]
        idcg = numpy.cumsum(1 / logs[:top])[-1]
        if idcg != 0:
            ndcgs.append(dcg / idcg)
    return numpy.mean(ndcgs, dtype=numpy.float16)


def mrr_at(n_recommendations, hits, lengths):
    """evaluator.py:328-341."""
    dense = _matrix(hits)
    if dense is not None:
        n = min(n_recommendations, dense.shape[1])
        tops = numpy.minimum(numpy.asarray(lengths, dtype=numpy.int64), n_recommendations)
        ones = (dense[:, :n] == 1) & (numpy.arange(n)[None, :] < tops[:, None])
        found = ones.any(axis=1)
        first = ones.argmax(axis=1)
        values = numpy.where(found, 1.0 / (first + 1), 0.0)
        return numpy.mean(values, dtype=numpy.float16)
    mrr_list = []
    for u in range(len(hits)):
        top = min(int(lengths[u]), n_recommendations)
        first = numpy.nonzero(numpy.asarray(hits[u][:top]) == 1)[0]
        mrr_list.append(1.0 / (first[0] + 1) if len(first) else 0)
    return numpy.mean(mrr_list, dtype=numpy.float16)
