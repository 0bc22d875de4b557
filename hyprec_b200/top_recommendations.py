"""
Host-side mirror of /root/reference/util/top_recommendations.py (TopRecommendations, :8-95) plus
the order rule the device top-k kernel reproduces.

Semantics that matter (util/top_recommendations.py:31-40): values are kept
ascending; a new value goes to the right of equal values (``bisect`` == bisect_right); when the
list is full a candidate is admitted only if it is STRICTLY greater than the current minimum,
and the minimum that leaves is the earliest-inserted of the minimal values.  After
``reversed()`` (evaluator.py:230) the list is: score descending, equal scores later-inserted
first.
"""
import bisect

import numpy


class TopRecommendations(object):
    def __init__(self, n_recommendations):
        self.recommendations_values = []
        self.recommendations_indices = []
        self.n_recommendations = n_recommendations

    def insert(self, index, value):
        values = self.recommendations_values
        if self.n_recommendations > len(values):
            self._place(index, value)
        elif len(values) != 0 and values[0] < value:
            values.pop(0)
            self.recommendations_indices.pop(0)
            self._place(index, value)

    def _place(self, index, value):
        at = bisect.bisect(self.recommendations_values, value)
        self.recommendations_values.insert(at, value)
        self.recommendations_indices.insert(at, index)

    def get_indices(self):
        return self.recommendations_indices

    def get_values(self):
        return self.recommendations_values

    def get_recommendations_count(self):
        return len(self.recommendations_values)


def resolve_cutoff_ties(scores, capacity):
    """Positions (ascending) of the elements the stream keeps, for one candidate stream.

    Everything strictly above the capacity-th largest value is always kept.  Elements EQUAL to
    that cut-off behave as a FIFO: one is admitted while the list is not full, rejected when it
    is (strict ``<``), and every later strictly-larger arrival that finds the list full evicts
    the oldest of them.  Elements below the cut-off never influence which of the others stay.
    """
    scores = numpy.asarray(scores, dtype=numpy.float64)
    total = len(scores)
    if total <= capacity:
        return numpy.arange(total)
    cut = numpy.partition(scores, total - capacity)[total - capacity]
    above = scores > cut
    equal = scores == cut
    need = capacity - int(above.sum())
    if int(equal.sum()) == need:
        return numpy.nonzero(above | equal)[0]
    held = 0          # elements >= cut currently in the list
    fifo = []         # positions of cut-valued elements in the list, oldest first
    for pos in numpy.nonzero(above | equal)[0]:
        if above[pos]:
            if held == capacity:
                fifo.pop(0)
            else:
                held += 1
        elif held < capacity:
            fifo.append(pos)
            held += 1
    return numpy.sort(numpy.concatenate((numpy.nonzero(above)[0], numpy.asarray(fifo, dtype=numpy.int64))))


def top_k_stream_order(ids, scores, capacity):
    """(ids, scores) of the reference's streamed top-``capacity`` in final (descending) order."""
    ids = numpy.asarray(ids)
    scores = numpy.asarray(scores, dtype=numpy.float64)
    if len(ids) == 0 or capacity <= 0:
        return [], []
    kept = resolve_cutoff_ties(scores, capacity)
    order = numpy.lexsort((-kept, -scores[kept]))
    sel = kept[order]
    return [int(i) for i in ids[sel]], [float(v) for v in scores[sel]]


def naive_split_recommendations(score0, score1, test_indptr, test_indices, test_values, n_items, capacity):
    """The lists ``load_top_recommendations`` produces under the NAIVE split, in closed form.

    There ``evaluator.test_indices`` is the dense test matrix (evaluator.py:97), so the "candidates" of user u are the
    n_items VALUES of its test row cast to int (evaluator.py:225-228): item 1 wherever the row holds a 1, item 0
    elsewhere.  Only two scores enter -- s0 = score(u, item 0), s1 = score(u, item 1) -- and the
    streamed top-``capacity`` (util/top_recommendations.py:23-40) is determined by their order and by how many of
    each there are:
      s1 > s0 : every 1 (at most capacity), then 0s;     s1 < s0 : every 0 (at most capacity), then 1s;
      s1 == s0: the first ``capacity`` stream elements, reversed (later-inserted first among equals).
    Returns (ids int32 [n_users][capacity] padded with -1, lengths) or None when a test value does not cast to 0 or 1
    (then the general path must stream the row)."""
    score0 = numpy.asarray(score0, dtype=numpy.float64)
    score1 = numpy.asarray(score1, dtype=numpy.float64)
    n_users = len(score0)
    casted = numpy.asarray(test_values).astype(numpy.int64)      # int(value), evaluator.py:226
    if numpy.any((casted != 0) & (casted != 1)):
        return None
    total = min(int(capacity), int(n_items))
    ids = numpy.full((n_users, capacity), -1, dtype=numpy.int32)
    test_indptr = numpy.asarray(test_indptr, dtype=numpy.int64)
    rows = numpy.repeat(numpy.arange(n_users, dtype=numpy.int64), numpy.diff(test_indptr))
    n1 = numpy.bincount(rows[casted == 1], minlength=n_users).astype(numpy.int64)     # ones in the user's test row
    n0 = n_items - n1
    slot = numpy.arange(total, dtype=numpy.int64)[None, :]
    ones_first = (slot < numpy.minimum(total, n1)[:, None]).astype(numpy.int32)        # 1 ... 1 0 ... 0
    zeros_first = (slot >= numpy.minimum(total, n0)[:, None]).astype(numpy.int32)      # 0 ... 0 1 ... 1
    ids[:, :total] = numpy.where((score1 > score0)[:, None], ones_first, zeros_first)
    for user in numpy.nonzero(~(score1 > score0) & ~(score1 < score0))[0]:    # (rare: e.g. all-zero factor rows)
        lo, hi = test_indptr[user], test_indptr[user + 1]
        ones_at = numpy.asarray(test_indices[lo:hi])[casted[lo:hi] == 1]
        first = numpy.zeros(total, dtype=numpy.int32)
        first[ones_at[ones_at < total]] = 1
        ids[user, :total] = first[::-1]
    return ids, numpy.full(n_users, total, dtype=numpy.int64)
