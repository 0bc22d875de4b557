"""
The scalar metrics of the evaluation report, computed on the host from the integers / flags the
device produces.  Each per-user value is formed with the reference's own arithmetic and the
list is handed to the same ``numpy.mean(..., dtype=numpy.float16)`` call
(/root/reference/lib/evaluator.py:292, :315, :341), so the reported scalars are bit-identical
whenever the top-k lists and hit flags are.
"""
import numpy


def recall_at_x(x, hits, test_likes):
    """evaluator.py:281-292.  hits: (m, K) array or list of per-user arrays of rating*rounded along
    the recommendation list (padding must be 0); test_likes[u] = test[u].sum()."""
    recalls = [numpy.sum(hits[u][:x]) / (min(x, test_likes[u]) * 1.0)
               for u in range(len(hits)) if test_likes[u] != 0]
    return numpy.mean(recalls, dtype=numpy.float16)


def ndcg_at(n_recommendations, hits, lengths):
    """evaluator.py:304-315; lengths[u] = len(recommendation_indices[u])."""
    logs = numpy.log2(numpy.arange(max(n_recommendations, 1)) + 2)
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
    mrr_list = []
    for u in range(len(hits)):
        top = min(int(lengths[u]), n_recommendations)
        first = numpy.nonzero(numpy.asarray(hits[u][:top]) == 1)[0]
        mrr_list.append(1.0 / (first[0] + 1) if len(first) else 0)
    return numpy.mean(mrr_list, dtype=numpy.float16)
