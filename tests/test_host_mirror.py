"""Host-side mirrors (Evaluator, TopRecommendations, ModelInitializer, metrics) against the golden
outputs of the unmodified reference, plus the reference's own invariants
(/root/reference/tests/evaluator_tests.py:42-100, top_recommendations_tests.py:18-46,
util_tests.py:110-124)."""
import numpy
import pytest

from conftest import golden_folds, load_golden, split_lists
from hyprec_b200 import metrics
from hyprec_b200.evaluator import Evaluator
from hyprec_b200.model_initializer import ModelInitializer
from hyprec_b200.top_recommendations import TopRecommendations, top_k_stream_order


@pytest.mark.parametrize("name", ["als_kfold_4x30", "als_kfold_toy"])
def test_kfold_split_is_bit_identical(als_cases, name):
    g = als_cases[name]
    ev = Evaluator(g["ratings"])
    ev.set_kfolds(int(g["k_folds"]))
    lists = ev.get_kfold_indices()
    want = split_lists(g["kfold_ids"], g["kfold_len"])
    assert len(lists) == len(want)
    for a, b in zip(lists, want):
        assert numpy.array_equal(numpy.asarray(a).astype(numpy.int64), b)
    for f, fold in enumerate(golden_folds(g)):
        train, test = ev.get_fold(f, lists)
        assert numpy.array_equal(train, fold["train"])
        assert numpy.array_equal(test, fold["test"])
        indptr, ids = ev.candidate_lists(f)
        assert numpy.array_equal(numpy.diff(indptr), fold["cand_len"])
        assert numpy.array_equal(ids, fold["cand_ids"])


def test_naive_split_is_bit_identical(als_cases):
    g = als_cases["als_naive_toy"]
    fold = golden_folds(g)[0]
    ev = Evaluator(g["ratings"])
    ev.set_kfolds(1)
    train, test = ev.naive_split()
    assert numpy.array_equal(train, fold["train"])
    assert numpy.array_equal(test, fold["test"])
    assert numpy.all(train * test == 0)
    indptr, ids = ev.candidate_lists(0)
    assert numpy.array_equal(ids, fold["cand_ids"])


def test_kfold_invariants():
    """evaluator_tests.py:50-74: folds are disjoint, cover every rating once, train/test disjoint."""
    ratings = numpy.array([[int(not bool((a + u) % 3)) for a in range(18)] for u in range(10)])
    ev = Evaluator(ratings)
    ev.set_kfolds(3)
    lists = ev.get_kfold_indices()
    seen = numpy.zeros(ratings.shape)
    for f in range(3):
        train, test = ev.get_fold(f, lists)
        assert numpy.all(train * test == 0)
        assert numpy.array_equal(train + test, ratings)
        seen += test
    assert numpy.array_equal(seen, ratings)


def test_rmse_matches_definition():
    rs = numpy.random.RandomState(5)
    a, b = rs.rand(7, 9), rs.rand(7, 9)
    ev = Evaluator(b)
    assert abs(ev.get_rmse(a, b) - numpy.sqrt(numpy.mean((a - b) ** 2))) < 1e-12
    assert abs(ev.get_rmse(a) - numpy.sqrt(numpy.mean((a - b) ** 2))) < 1e-12


def test_top_recommendations_golden_and_invariants():
    g = load_golden("top_recommendations")
    for c in range(int(g["n_cases"])):
        cap, stream, want = int(g["c%d_cap" % c]), g["c%d_stream" % c], list(g["c%d_result" % c])
        top = TopRecommendations(cap)
        for idx, val in enumerate(stream):
            top.insert(idx, float(val))
        assert list(reversed(top.get_indices())) == want
        vals = top.get_values()
        assert all(vals[i] <= vals[i + 1] for i in range(len(vals) - 1))
        assert top.get_recommendations_count() == min(cap, len(stream)) == len(top.get_indices())
        assert top_k_stream_order(numpy.arange(len(stream)), stream, cap)[0] == want


def test_metrics_golden():
    g = load_golden("metrics")
    ratings, expected = g["ratings"], g["expected"]
    recs = [list(r) for r in split_lists(g["rec_ids"], g["rec_len"])]
    ev = Evaluator(ratings)
    ev.recs_loaded = True
    ev.recommendation_indices = recs
    hits = [ratings[u][numpy.asarray(r, dtype=int)] * expected[u][numpy.asarray(r, dtype=int)] for u, r in enumerate(recs)]
    lengths = [len(r) for r in recs]
    for n in (1, 4, 5):
        assert ev.calculate_mrr(n, None, ratings, expected) == numpy.float16(g["mrr_%d" % n])
        assert ev.calculate_ndcg(n, None, ratings, expected) == numpy.float16(g["ndcg_%d" % n])
        assert metrics.mrr_at(n, hits, lengths) == numpy.float16(g["mrr_%d" % n])
        assert metrics.ndcg_at(n, hits, lengths) == numpy.float16(g["ndcg_%d" % n])
    assert ev.recall_at_x(3, None, ratings, expected) == numpy.float16(g["recall_3"])
    assert metrics.recall_at_x(3, hits, ratings.sum(axis=1)) == numpy.float16(g["recall_3"])
    # evaluator_tests.py:94-100: deleting top hits cannot raise MRR
    before = ev.calculate_mrr(5, None, ratings, expected)
    ev.recommendation_indices = [r[1:] for r in recs]
    assert ev.calculate_mrr(5, None, ratings, expected) <= before


def test_load_top_recommendations_matches_reference(als_cases):
    g = als_cases["als_kfold_toy"]
    ev = Evaluator(g["ratings"])
    ev.set_kfolds(int(g["k_folds"]))
    ev.get_kfold_indices()
    for f, fold in enumerate(golden_folds(g)):
        pred = fold["u"].dot(fold["v"].T)
        recs = ev.load_top_recommendations(200, pred, fold["test"], f)
        want = [list(w) for w in split_lists(fold["rec_ids"], fold["rec_len"])]
        assert [list(r) for r in recs] == want


def test_model_initializer_roundtrip(tmp_path):
    """util_tests.py:110-124: file name and save/load equality."""
    init = ModelInitializer({'n_factors': 3, '_lambda': 0.01}, 1, folder=str(tmp_path))
    path = init._create_path('user_v', (10, 3), {'n_iterations': 1})
    assert path.endswith('n_factors-3,n_iterations-1,n_rows-10user_v.dat')
    matrix = numpy.random.RandomState(0).rand(10, 3)
    init.save_matrix(matrix, 'user_v')
    found, loaded = init.load_matrix({'n_factors': 3, '_lambda': 0.01}, 'user_v', (10, 3))
    assert found and numpy.array_equal(loaded, matrix)
    found, fresh = init.load_matrix({'n_factors': 3, '_lambda': 0.01}, 'missing', (4, 3))
    assert not found and fresh.shape == (4, 3)


def test_fold_matrices_and_csr_views_match_the_literal_construction():
    """generate_kfold_matrix (vectorised) against the per-user loop of evaluator.py:206-213, with non-binary
    ratings and repeated / unrated ids in the lists; csr_view returns exactly what a scan of the dense matrices
    gives and refuses a matrix that was edited in place."""
    from hyprec_b200.collaborative_filtering import _csr_triplet
    rs = numpy.random.RandomState(8)
    m, n = 31, 57
    ratings = (rs.rand(m, n) < 0.3) * rs.randint(1, 4, size=(m, n)).astype(numpy.float64)
    ratings[4, :] = 0
    ev = Evaluator(ratings)
    ev.set_kfolds(3)
    lists = [rs.randint(0, n, size=int(rs.randint(0, 25))) for _ in range(m)]
    train, test = ev.generate_kfold_matrix(lists)
    want_train, want_test = numpy.array(ratings, dtype=numpy.float64), numpy.zeros(ratings.shape)
    for user in range(m):
        want_test[user, lists[user]] = ratings[user, lists[user]]
        want_train[user, lists[user]] = 0
    assert numpy.array_equal(train, want_train) and numpy.array_equal(test, want_test)
    for matrix in (train, test, ratings):
        got, want = ev.csr_view(matrix), _csr_triplet(matrix)
        assert got is not None
        for a, b in zip(got, want):
            assert a.dtype == b.dtype and numpy.array_equal(a, b)
    assert ev.csr_view(train.copy()) is None                 # not the object the evaluator built
    r, c = numpy.argwhere(train != 0)[0]
    train[r, c] = 7.0                                        # value edited in place
    assert ev.csr_view(train) is None
    train[r, c] = want_train[r, c]
    assert ev.csr_view(train) is not None
    zr, zc = numpy.argwhere(train == 0)[0]
    train[zr, zc] = 1.0                                      # entry added in place
    assert ev.csr_view(train) is None


def test_sparse_ratings_give_the_same_splits_without_dense_matrices():
    """Evaluator(scipy sparse ratings): the k-fold and naive splits consume the random stream exactly as with the
    dense array and hand out scipy CSR train / test matrices equal to the dense ones; candidate
    streams are identical, including the naive split's ids-0/1 degeneration (App. A Q7)."""
    import scipy.sparse as sp
    from hyprec_b200 import Evaluator, synth
    full = synth.ratings_csr("toy", n_users=60, n_items=200, positives=1500, min_degree=2, seed=3)
    dense = numpy.asarray(full.todense())
    for folds in (5, 1):
        ed, es = Evaluator(dense.copy()), Evaluator(full.copy())
        ed.set_kfolds(folds)
        es.set_kfolds(folds)
        if folds == 1:
            td, sd = ed.naive_split()
            ts, ss = es.naive_split()
            assert sp.issparse(ts) and sp.issparse(ss)
            assert (ts != sp.csr_matrix(td)).nnz == 0 and (ss != sp.csr_matrix(sd)).nnz == 0
            cd, cs = ed.candidate_lists(0), es.candidate_lists(0)
            assert numpy.array_equal(cd[0], cs[0]) and numpy.array_equal(cd[1], cs[1])
            continue
        ld, ls = ed.get_kfold_indices(), es.get_kfold_indices()
        assert all(numpy.array_equal(a, b) for a, b in zip(ld, ls))
        for f in range(folds):
            td, sd = ed.get_fold(f, ld)
            ts, ss = es.get_fold(f, ls)
            assert (ts != sp.csr_matrix(td)).nnz == 0 and (ss != sp.csr_matrix(sd)).nnz == 0
            view = es.csr_view(ts)
            assert view is not None and numpy.array_equal(view[1], sp.csr_matrix(td).indices)
            cd, cs = ed.candidate_lists(f), es.candidate_lists(f)
            assert numpy.array_equal(cd[0], cs[0]) and numpy.array_equal(cd[1], cs[1])


def test_class_runs_on_sparse_ratings_and_refuses_huge_dense_predictions(monkeypatch):
    """The reference-facing class on scipy sparse ratings (oracle-backed stand-in for the device): the same
    11-tuple as on the dense array; get_predictions() raises MemoryError above the dense limit."""
    import scipy.sparse as sp
    from grid_helpers import OracleBacked, make_recommender
    from hyprec_b200 import collaborative_filtering as cfmod, synth
    full = synth.ratings_csr("toy", n_users=40, n_items=90, positives=700, min_degree=2, seed=5)
    dense = numpy.asarray(full.todense())
    a = make_recommender(OracleBacked, dense.copy())
    b = make_recommender(OracleBacked, sp.csr_matrix(full))
    ra = numpy.asarray(a.train(), dtype=numpy.float64)
    rb = numpy.asarray(b.train(), dtype=numpy.float64)
    assert sp.issparse(b.train_data) and sp.issparse(b.test_data)
    same = (numpy.isnan(ra) & numpy.isnan(rb)) | (numpy.abs(ra - rb) <= 1e-12 * numpy.maximum(1, numpy.abs(ra)))
    assert numpy.all(same), (ra, rb)
    from hyprec_b200 import CollaborativeFiltering
    monkeypatch.setattr(cfmod, "DENSE_PREDICTIONS_LIMIT_BYTES", 1000)
    b.predictions = None
    with pytest.raises(MemoryError):
        CollaborativeFiltering.get_predictions(b)


def test_naive_split_lists_in_closed_form():
    """naive_split_recommendations == streaming the dense test row through TopRecommendations (App. A Q7), for
    every order of the two scores, more / fewer ones than the capacity, ties, and rows shorter than the capacity."""
    import scipy.sparse as sp
    from hyprec_b200.top_recommendations import naive_split_recommendations, top_k_stream_order
    rs = numpy.random.RandomState(0)
    for n_items, capacity in ((60, 20), (15, 20), (40, 7)):
        m = 30
        test = (rs.rand(m, n_items) < rs.rand(m, 1) * 0.8).astype(numpy.float64)
        test[3] = 0
        test[4] = 1
        scores = rs.randn(m, n_items)
        scores[5:12, 1] = scores[5:12, 0]            # tied
        scores[12] = 0.0
        csr = sp.csr_matrix(test)
        got = naive_split_recommendations(scores[:, 0], scores[:, 1], csr.indptr, csr.indices, csr.data, n_items, capacity)
        assert got is not None
        ids, lengths = got
        for u in range(m):
            stream = test[u].astype(numpy.int64)
            want, _ = top_k_stream_order(stream, scores[u][stream], capacity)
            assert list(ids[u][:lengths[u]]) == want, (n_items, capacity, u)
            assert numpy.all(ids[u][lengths[u]:] == -1)
    csr = sp.csr_matrix(numpy.array([[0, 2.0, 0], [1.0, 0, 0]]))
    assert naive_split_recommendations([0, 0], [1, 1], csr.indptr, csr.indices, csr.data, 3, 2) is None


def test_vectorised_metrics_equal_the_per_user_loops():
    """metrics.recall_at_x / ndcg_at / mrr_at on the device's (m, K) zero-padded hit matrix give bit for bit the
    float16 values of the per-user loops (the reference's arithmetic, evaluator.py:281-341)."""
    from hyprec_b200 import metrics
    rs = numpy.random.RandomState(1)
    for m, cap in ((700, 200), (9000, 40), (33, 7)):
        lengths = numpy.where(rs.rand(m) < 0.1, rs.randint(0, cap, size=m), cap)
        hits = (rs.rand(m, cap) < 0.15).astype(numpy.float64)
        hits[numpy.arange(cap)[None, :] >= lengths[:, None]] = 0.0
        likes = rs.randint(0, 60, size=m).astype(numpy.float64)
        rows = [hits[u][:lengths[u]] for u in range(m)]                 # the list form runs the loops
        for x in (5, 50, 200, 300):
            a, b = metrics.recall_at_x(x, hits, likes), metrics.recall_at_x(x, rows, likes)
            assert a.dtype == numpy.float16 and (a == b or (numpy.isnan(a) and numpy.isnan(b)))
        for n in (1, 5, 10, 250):
            a, b = metrics.ndcg_at(n, hits, lengths), metrics.ndcg_at(n, rows, lengths)
            assert a == b or (numpy.isnan(a) and numpy.isnan(b)), (n, a, b)
            a, b = metrics.mrr_at(n, hits, lengths), metrics.mrr_at(n, rows, lengths)
            assert a == b, (n, a, b)


def test_matrix_facts_cache_hits_and_notices_edits():
    """CollaborativeFiltering._facts: CSR triplet / total / row sums per matrix object, for dense and sparse matrices,
    re-used on the second call and dropped when the caller edits the dense matrix in place."""
    import scipy.sparse as sp
    from grid_helpers import OracleBacked, make_recommender
    from hyprec_b200 import synth
    full = synth.ratings_csr("toy", n_users=30, n_items=70, positives=400, min_degree=2, seed=8)
    dense = numpy.asarray(full.todense())
    cf = make_recommender(OracleBacked, dense.copy())
    for matrix in (dense.copy(), sp.csr_matrix(full)):
        first = cf._facts(matrix)
        again = cf._facts(matrix)
        assert len(first) == 3 and len(again) == 3 and again[0] is first[0]
        assert first[1] == 400.0 and numpy.array_equal(first[2], numpy.asarray(full.sum(axis=1)).ravel())
    matrix = dense.copy()
    before = cf._facts(matrix)
    rows, cols = numpy.nonzero(matrix)
    matrix[rows, cols] = 0.0                      # the caller empties the matrix in place
    after = cf._facts(matrix)
    assert after[1] == 0.0 and after[0] is not before[0]


def test_factor_fingerprint_sees_every_element():
    """collaborative_filtering._fingerprint decides whether the device factors are stale: an in-place edit of ANY
    element (here far from where a strided sample would look) must change it; equal contents must not."""
    from hyprec_b200.collaborative_filtering import CollaborativeFiltering
    rs = numpy.random.RandomState(5)
    u, v = rs.rand(700, 33), rs.rand(1900, 33)
    token = CollaborativeFiltering._fingerprint(u, v)
    assert CollaborativeFiltering._fingerprint(u.copy(), v.copy()) == token
    for arr, pos in ((u, (123, 7)), (v, (1777, 31)), (v, (0, 1))):
        old = arr[pos]
        arr[pos] = numpy.nextafter(old, 2.0)
        assert CollaborativeFiltering._fingerprint(u, v) != token
        arr[pos] = old
    assert CollaborativeFiltering._fingerprint(u, v) == token
    swapped = v.copy()
    swapped[[3, 4]] = swapped[[4, 3]]                # a permutation keeps the xor: the CRC must see it
    assert CollaborativeFiltering._fingerprint(u, swapped) != token
    assert CollaborativeFiltering._fingerprint(u.astype(numpy.float32), v) != token


def _kfold_both_ways(ratings, k_folds, monkeypatch):
    got = []
    for native in ("0", "1"):
        monkeypatch.setenv("HYPREC_NATIVE_SPLIT", native)
        ev = Evaluator(ratings.copy())
        ev.set_kfolds(k_folds)
        lists = ev.get_kfold_indices()
        after = numpy.random.random(4)                  # the stream right behind the split (initial factors come next)
        folds = [ev.candidate_lists(f) for f in range(k_folds)]
        got.append((lists, ev._slot_positives[1], after, folds, ev))
    return got


def test_native_kfold_lists_equal_the_numpy_loop(monkeypatch):
    """hyprec_host_kfold_lists (csrc/host_split.cuh: MT19937 + legacy shuffle restated in C++) against the loop of
    numpy.random.shuffle calls it replaces (evaluator.py:144-192): same lists, same number of leading positives, and
    the global stream left at the same position -- on random small shapes (ragged lists, empty users, k_folds = 1,
    more folds than positives), incl. shapes where a fold's positives outnumber n_items / k_folds and the native call
    must refuse without drawing anything."""
    rs = numpy.random.RandomState(11)
    native_used = refused = 0
    for trial in range(150):
        m, n, k = rs.randint(1, 7), rs.randint(1, 60), rs.randint(1, 8)
        ratings = (rs.rand(m, n) < rs.rand()).astype(float)
        py, nat = _kfold_both_ways(ratings, k, monkeypatch)
        assert len(py[0]) == len(nat[0]) == m * k
        for a, b in zip(py[0], nat[0]):
            assert numpy.array_equal(numpy.asarray(a, dtype=numpy.int64), numpy.asarray(b, dtype=numpy.int64))
        assert numpy.array_equal(py[1], nat[1])
        assert numpy.array_equal(py[2], nat[2])
        for (pa, ia), (pb, ib) in zip(py[3], nat[3]):
            assert numpy.array_equal(pa, pb) and numpy.array_equal(ia, ib)
        if getattr(nat[4], '_flat', None) is not None and nat[4]._flat[0] is nat[0]:
            native_used += 1
        else:
            refused += 1
    assert native_used > 50 and refused > 0


def test_native_kfold_lists_at_a_long_row(monkeypatch):
    """Rows long enough to cross several refills of the 624-word generator state and several mask widths."""
    rs = numpy.random.RandomState(5)
    ratings = (rs.rand(7, 5000) < 0.01).astype(float)
    ratings[3] = 0.0                                    # a user without ratings
    py, nat = _kfold_both_ways(ratings, 5, monkeypatch)
    assert nat[4]._flat[0] is nat[0] and nat[0][0].dtype == numpy.int32
    for a, b in zip(py[0], nat[0]):
        assert numpy.array_equal(a, b)
    assert numpy.array_equal(py[2], nat[2])
    train_py, test_py = py[4].get_fold(2, py[0])
    train_nat, test_nat = nat[4].get_fold(2, nat[0])
    assert numpy.array_equal(train_py, train_nat) and numpy.array_equal(test_py, test_nat)


@pytest.mark.parametrize("sparse", [False, True])
def test_native_naive_split_equals_the_numpy_loop(monkeypatch, sparse):
    """hyprec_host_naive_picks against the per-user numpy.random.choice calls it replaces (evaluator.py:92-93): same
    train / test matrices and the same stream position, for several test percentages, users with one rating (no draw)
    and users without any."""
    import scipy.sparse as sp
    rs = numpy.random.RandomState(3)
    for trial in range(40):
        m, n, k = rs.randint(1, 9), rs.randint(2, 70), rs.randint(1, 6)
        dense = (rs.rand(m, n) < rs.rand() * 0.6).astype(float) * rs.randint(1, 4, size=(m, n))
        if trial % 3 == 0:
            dense[rs.randint(m)] = 0.0
        if trial % 4 == 0:
            dense[rs.randint(m)] = 0.0
            dense[rs.randint(m), rs.randint(n)] = 1.0
        got = []
        for native in ("0", "1"):
            monkeypatch.setenv("HYPREC_NATIVE_SPLIT", native)
            ev = Evaluator(sp.csr_matrix(dense) if sparse else dense.copy())
            ev.set_kfolds(k)
            train, test = ev.naive_split()
            after = numpy.random.random(3)
            as_dense = (lambda x: numpy.asarray(x.todense())) if sparse else (lambda x: x)
            got.append((as_dense(train), as_dense(test), after, ev.csr_view(train), ev.csr_view(test)))
        py, nat = got
        assert numpy.array_equal(py[0], nat[0]) and numpy.array_equal(py[1], nat[1])
        assert numpy.array_equal(py[2], nat[2])
        for a, b in zip(py[3] + py[4], nat[3] + nat[4]):
            assert numpy.array_equal(a, b)


def test_native_kfold_lists_do_not_depend_on_the_thread_count(monkeypatch):
    """The stream is walked once sequentially and the shuffles run on host threads from recorded generator states
    (csrc/host_split.cuh): one thread, three threads and more threads than users give the same lists and leave the
    global generator at the same position."""
    rs = numpy.random.RandomState(9)
    ratings = (rs.rand(23, 400) < 0.05).astype(float)
    got = []
    for threads in ("1", "3", "64"):
        monkeypatch.setenv("HYPREC_HOST_THREADS", threads)
        ev = Evaluator(ratings.copy())
        ev.set_kfolds(4)
        lists = ev.get_kfold_indices()
        assert ev._flat[0] is lists
        got.append((numpy.concatenate(lists), numpy.random.random(3)))
    for other in got[1:]:
        assert numpy.array_equal(got[0][0], other[0]) and numpy.array_equal(got[0][1], other[1])
