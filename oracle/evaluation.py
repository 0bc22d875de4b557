"""
TEST INFRASTRUCTURE -- CPU restatement of the prediction / top-k / evaluation-report half of
the hot path (see oracle/__init__.py).

Follows, line by line where semantics are subtle:
  * util/top_recommendations.py:23-54   bounded ascending list, bisect_right, strict ``<`` admission
  * lib/evaluator.py:215-234            load_top_recommendations (candidate stream -> descending ids)
  * lib/abstract_recommender.py:173-187 rounded_predictions (sequential Python ``sum`` / n threshold)
  * lib/evaluator.py:254-341            calculate_recall, recall_at_x, calculate_ndcg, calculate_mrr
  * lib/abstract_recommender.py:100-131 get_evaluation_report (the 11-tuple)
"""
import bisect

import numpy
import scipy.sparse as sp


# ----------------------------------------------------------------------------------------------
# top-k
# ----------------------------------------------------------------------------------------------
def top_recommendations_stream(ids, scores, capacity):
    """Literal restatement of TopRecommendations.insert over a candidate stream
    (top_recommendations.py:31-40) followed by ``list(reversed(get_indices()))``
    (evaluator.py:230).  Pure Python; use for small cases and for tie replay."""
    values, indices = [], []
    for idx, val in zip(ids, scores):
        if capacity > len(values):
            pos = bisect.bisect(values, val)
            values.insert(pos, val)
            indices.insert(pos, int(idx))
        elif len(values) != 0:
            if values[0] < val:
                values.pop(0)
                indices.pop(0)
                pos = bisect.bisect(values, val)
                values.insert(pos, val)
                indices.insert(pos, int(idx))
    return list(reversed(indices)), list(reversed(values))


def top_recommendations_fast(ids, scores, capacity):
    """Same result as ``top_recommendations_stream`` in O(T log T).

    Without a tie straddling the cut-off the stream result equals sorting by
    (score descending, stream position descending) -- equal scores: later-inserted first
    (SURVEY.md App. A Q9).  When the cut-off value is tied and not all tied elements fit, which
    of them survive depends on the stream dynamics; that case replays the stream restricted to
    the elements >= cut-off (elements below the cut-off never change which of the others are
    kept: they can only be evicted by, never evict, an element >= cut-off)."""
    ids = numpy.asarray(ids)
    scores = numpy.asarray(scores, dtype=numpy.float64)
    total = len(ids)
    if total == 0 or capacity <= 0:
        return [], []
    pos = numpy.arange(total)
    if total > capacity:
        cut = numpy.partition(scores, total - capacity)[total - capacity]
        keep = scores >= cut
        if int(keep.sum()) != capacity:
            sub = numpy.nonzero(keep)[0]
            return top_recommendations_stream(ids[sub], scores[sub], capacity)
        pos = pos[keep]
    order = numpy.lexsort((-pos, -scores[pos]))
    sel = pos[order]
    return [int(i) for i in ids[sel]], [float(v) for v in scores[sel]]


# ----------------------------------------------------------------------------------------------
# thresholds / rounded predictions
# ----------------------------------------------------------------------------------------------
def row_threshold(pred_row):
    """abstract_recommender.py:183: ``sum(predictions[user]) / n`` -- Python's sequential
    float64 sum.  numpy.cumsum accumulates in the same order with the same adds."""
    return numpy.cumsum(pred_row)[-1] / pred_row.shape[0]


def rounded_predictions(predictions):
    """abstract_recommender.py:180-187: 0 where score < row average, else 1 (float64)."""
    out = numpy.ones_like(predictions)
    for u in range(predictions.shape[0]):
        out[u, predictions[u] < row_threshold(predictions[u])] = 0
    return out


# ----------------------------------------------------------------------------------------------
# metrics from per-user hit lists (the arithmetic of evaluator.py:268-341 on exact inputs)
# ----------------------------------------------------------------------------------------------
def recall_at_x(x, hits, test_likes):
    """evaluator.py:281-292. ``hits[u]`` = ratings*rounded along the user's recommendation list,
    ``test_likes[u]`` = test[u].sum().  Users without test likes are skipped (:286)."""
    recalls = []
    for u in range(len(hits)):
        likes = test_likes[u]
        if likes != 0:
            recalls.append(numpy.sum(hits[u][:x]) / (min(x, likes) * 1.0))
    return numpy.mean(recalls, dtype=numpy.float16)


def ndcg_at(n_recommendations, hits):
    """evaluator.py:304-315: ideal DCG counts every listed position (not bounded by #likes)."""
    ndcgs = []
    for u in range(len(hits)):
        dcg = 0
        idcg = 0
        for pos_index, hit in enumerate(hits[u]):
            dcg += hit / numpy.log2(pos_index + 2)
            idcg += 1 / numpy.log2(pos_index + 2)
            if pos_index + 1 == n_recommendations:
                break
        if idcg != 0:
            ndcgs.append(dcg / idcg)
    return numpy.mean(ndcgs, dtype=numpy.float16)


def mrr_at(n_recommendations, hits):
    """evaluator.py:328-341: reciprocal rank of the first hit == 1 within the first n."""
    mrr_list = []
    for u in range(len(hits)):
        mrr = 0
        for mrr_index, hit in enumerate(hits[u]):
            if hit == 1:
                mrr = hit / (mrr_index + 1)
                break
            if mrr_index + 1 == n_recommendations:
                break
        mrr_list.append(mrr)
    return numpy.mean(mrr_list, dtype=numpy.float16)


# ----------------------------------------------------------------------------------------------
# the evaluation report
# ----------------------------------------------------------------------------------------------
def evaluation_report(user_vecs, item_vecs, ratings, train, test, candidates, n_recommendations=200,
                      recall_x=200, block=256, fast_topk=True):
    """abstract_recommender.py:100-131 from factors, never holding more than ``block`` rows of
    the prediction matrix.

    :param ratings, train, test: dense ndarray or scipy sparse, users x items (full / train / test).
    :param candidates: list of per-user candidate id sequences IN STREAM ORDER
        (``evaluator.test_indices[user * (1 + fold)]``, evaluator.py:225), or None => all items
        in index order (abstract_recommender.py:200-203, recommend_items).
    :returns: dict with the 11-tuple under 'report' and every per-user intermediate.
    """
    ratings = sp.csr_matrix(ratings, dtype=numpy.float64)
    train = sp.csr_matrix(train, dtype=numpy.float64)
    test = sp.csr_matrix(test, dtype=numpy.float64)
    m, n = ratings.shape
    topk = top_recommendations_fast if fast_topk else top_recommendations_stream

    thresholds = numpy.zeros(m)
    ones_per_row = numpy.zeros(m, dtype=numpy.int64)
    rec_ids, rec_vals, hits = [], [], []
    train_hit = 0.0
    test_hit = 0.0
    sq_err = 0.0                                   # rss against TRAIN over all cells (:122)
    for s in range(0, m, block):
        e = min(m, s + block)
        pred = user_vecs[s:e].dot(item_vecs.T)
        full_blk = ratings[s:e].toarray()
        train_blk = train[s:e].toarray()
        test_blk = test[s:e].toarray()
        for u in range(s, e):
            row = pred[u - s]
            thr = row_threshold(row)
            thresholds[u] = thr
            rounded = (row >= thr)
            ones_per_row[u] = int(rounded.sum())
            cand = numpy.arange(n) if candidates is None else numpy.asarray(candidates[u]).astype(numpy.int64)
            ids, vals = topk(cand, row[cand], n_recommendations)
            rec_ids.append(ids)
            rec_vals.append(vals)
            ids_arr = numpy.asarray(ids, dtype=numpy.int64)
            hits.append(full_blk[u - s][ids_arr] * rounded[ids_arr])       # evaluator.py:287-288
            train_hit += rounded[train_blk[u - s] != 0].sum()               # evaluator.py:264-266
            test_hit += rounded[test_blk[u - s] != 0].sum()
            sq_err += numpy.sum((row - train_blk[u - s]) ** 2)              # evaluator.py:249

    test_likes = numpy.asarray(test.sum(axis=1)).ravel()
    test_sum = test.sum()
    train_sum = train.sum()
    out = dict(thresholds=thresholds, ones_per_row=ones_per_row, rec_ids=rec_ids, rec_vals=rec_vals,
               hits=hits)
    with numpy.errstate(divide="ignore", invalid="ignore"):
        train_recall = train_hit / train_sum
        test_recall = numpy.float64(test_hit) / test_sum
        ratio = float(ones_per_row.sum()) / ratings.sum()
    rmse = numpy.sqrt(float(sq_err) / (float(m) * float(n)))
    out["report"] = (test_sum, train_sum, rmse, train_recall, test_recall,
                     recall_at_x(recall_x, hits, test_likes), ratio,
                     mrr_at(5, hits), ndcg_at(5, hits), mrr_at(10, hits), ndcg_at(10, hits))
    return out


_M64 = (1 << 64) - 1


def hashed_mix(user, items, seed):
    """The 64-bit mix of hyprec_b200/csrc/topk.cuh (hashed_mix), vectorised over ``items`` (uint64 wrap-around)."""
    with numpy.errstate(over="ignore"):
        x = (numpy.uint64((int(user) * 0x9E3779B97F4A7C15) & _M64) ^
             (numpy.asarray(items, dtype=numpy.uint64) * numpy.uint64(0xC2B2AE3D27D4EB4F)) ^ numpy.uint64(int(seed) & _M64))
        x ^= x >> numpy.uint64(33)
        x *= numpy.uint64(0xFF51AFD7ED558CCD)
        x ^= x >> numpy.uint64(33)
        x *= numpy.uint64(0xC4CEB9FE1A85EC53)
        x ^= x >> numpy.uint64(33)
    return x


def hashed_candidates(ratings, test, users, fold, n_folds, seed):
    """Candidate lists of the scaled workload (no counterpart in the reference, whose lists come from per-user
    shuffles, evaluator.py:144-192; SURVEY.md section 8d row C4 asks for a "5-fold-equivalent" CSR-native split):
    user u's list = its test positives plus the unrated items i with ``mix(u, i, seed) % n_folds == fold``,
    ascending item ids.  Returns a list of int64 arrays, one per user of ``users``."""
    ratings = sp.csr_matrix(ratings)
    test = sp.csr_matrix(test)
    n = ratings.shape[1]
    items = numpy.arange(n, dtype=numpy.int64)
    out = []
    for u in users:
        rated = numpy.zeros(n, dtype=bool)
        rated[ratings.indices[ratings.indptr[u]:ratings.indptr[u + 1]]] = True
        tested = numpy.zeros(n, dtype=bool)
        tested[test.indices[test.indptr[u]:test.indptr[u + 1]]] = True
        share = (hashed_mix(u, items, seed) % numpy.uint64(n_folds)) == numpy.uint64(fold)
        out.append(items[tested | (~rated & share)])
    return out
