"""The hyper-parameter sweep (BASELINE.json configs[4]) against the rows the
unmodified reference GridSearch produced (tests/golden/grid_naive_toy.npz, made by
tests/golden/make_golden_grid.py): keys and combinations (/root/reference/tests/grid_search_tests.py:47-64),
the sequential flow, the device-backed flow with one and several workers, rank striding over gloo, and --
on the GPU -- the real thing through the C ABI."""
import socket

import numpy
import pytest

from conftest import load_golden
from grid_helpers import GRID, OracleBacked, make_recommender
from hyprec_b200 import grid_search as gs
from hyprec_b200.grid_search import GridSearch
from oracle import reference_loader


@pytest.fixture(scope="module")
def golden():
    return load_golden("grid_naive_toy")


def check_rows(rows, golden, rtol=1e-8):
    assert rows[0] == list(golden["header"])
    got = numpy.array([[float(x) for x in row] for row in rows[1:]])
    want = golden["rows"]
    assert got.shape == want.shape
    assert numpy.array_equal(got[:, :4], want[:, :4])                    # n_factors, _lambda, test / train sums
    assert numpy.allclose(got[:, 4], want[:, 4], rtol=rtol)              # RMSE
    assert numpy.allclose(got[:, 5:7], want[:, 5:7], rtol=1e-12)         # recalls: integer counts / sums
    assert numpy.array_equal(got[:, 7], want[:, 7])                      # recall@200 (float16)
    assert numpy.allclose(got[:, 8], want[:, 8], rtol=1e-12)             # ratio
    assert numpy.array_equal(got[:, 9:], want[:, 9:])                    # MRR / nDCG (float16)


def check_outcome(search, best, rows, golden):
    check_rows(rows, golden)
    assert [best['_lambda'], best['n_factors']] == list(golden["best"])
    errors = search.get_all_errors()
    assert list(golden["keys"]) == [search.get_key(c) for c in search.get_all_combinations()]
    for key, tr, te in zip(golden["keys"], golden["train_recall"], golden["test_recall"]):
        assert abs(errors[str(key)]['train_recall'] - tr) < 1e-12
        assert abs(errors[str(key)]['test_recall'] - te) < 1e-12


def test_keys_and_combinations(golden):
    search = GridSearch(make_recommender(OracleBacked, golden["ratings"]), {'_lambda': [0.0001, 0.1], 'n_factors': [10, 20]},
                        verbose=False)
    assert search.get_key({'_lambda': 0, 'n_factors': 10}) == search.get_key({'n_factors': 10, '_lambda': 0}) \
        == '_lambda:0,n_factors:10'
    combos = search.get_all_combinations()
    assert combos == [{'_lambda': 0.0001, 'n_factors': 10}, {'_lambda': 0.0001, 'n_factors': 20},
                      {'_lambda': 0.1, 'n_factors': 10}, {'_lambda': 0.1, 'n_factors': 20}]
    assert search.results_file_name == 'grid_search_results.csv'


def test_flag_lookup_matches_dense_recall():
    rs = numpy.random.RandomState(5)
    m, n = 23, 41
    ratings = (rs.rand(m, n) < 0.3).astype(numpy.int64)
    ratings[7, :] = 0
    in_test = (rs.rand(m, n) < 0.4) & (ratings != 0)
    test = numpy.where(in_test, ratings, 0).astype(numpy.float64)
    train = numpy.where(in_test, 0, ratings).astype(numpy.float64)
    rounded = (rs.rand(m, n) < 0.5).astype(numpy.float64)
    from hyprec_b200.collaborative_filtering import _csr_triplet
    patterns, flags = [], []
    for matrix in (train, test):
        indptr, indices, _ = _csr_triplet(matrix)
        rows = numpy.repeat(numpy.arange(m), numpy.diff(indptr))
        patterns.append((indptr, indices))
        flags.append(rounded[rows, indices].astype(numpy.uint8))
    other = numpy.where((rs.rand(m, n) < 0.5), ratings, 0)               # any subset of the positives
    ev = gs.Evaluator(ratings)
    for matrix in (ratings, test, train, other):
        got = gs.flags_at(matrix, patterns[0], flags[0], patterns[1], flags[1])
        assert gs.recall_from_flags(matrix, got) == ev.calculate_recall(matrix, rounded)
    outside = numpy.zeros((m, n))
    outside[7, 3] = 1                                                    # a cell the device never reported
    assert gs.flags_at(outside, patterns[0], flags[0], patterns[1], flags[1]) is None


@pytest.mark.skipif(not reference_loader.reference_available(), reason="reference tree not present")
def test_sequential_flow_drives_the_reference_class(golden, monkeypatch):
    ref = reference_loader.load_reference()
    hyper = {'n_factors': 4, '_lambda': 0.01}
    cf = ref.CollaborativeFiltering(ref.ModelInitializer(hyper.copy(), 5), ref.Evaluator(golden["ratings"]), hyper,
                                    {'k_folds': 1, 'n_iterations': 5}, load_matrices=False, dump_matrices=False)
    dumped = []
    monkeypatch.setattr(GridSearch, "dump_csv", lambda self, rows: dumped.append(rows))
    search = GridSearch(cf, GRID, verbose=False)
    assert not search._device_backed()
    best, rows = search.train()
    check_outcome(search, best, rows, golden)
    assert dumped == [rows]
    assert numpy.random.random() == float(golden["rng_after"])


@pytest.mark.parametrize("workers", [1, 3])
def test_device_flow_host_logic(golden, monkeypatch, workers):
    """Worker threads, generator hand-over and flag look-ups with the oracle standing in for the device."""
    monkeypatch.setattr(GridSearch, "dump_csv", lambda self, rows: None)
    rec = make_recommender(OracleBacked, golden["ratings"])
    search = GridSearch(rec, GRID, verbose=False, workers=workers)
    assert search._device_backed()
    best, rows = search.train()
    check_outcome(search, best, rows, golden)
    # the caller's recommender and the global generator end where the reference's loop leaves them
    assert numpy.random.random() == float(golden["rng_after"])
    assert rec.hyperparameters['n_factors'] == 8 and rec.hyperparameters['_lambda'] == 1.0
    assert numpy.allclose(rec.user_vecs, golden["final_user_vecs"], rtol=1e-8, atol=1e-12)
    assert numpy.allclose(rec.item_vecs, golden["final_item_vecs"], rtol=1e-8, atol=1e-12)


def test_checkpointing_recommenders_take_the_sequential_flow(golden):
    rec = make_recommender(OracleBacked, golden["ratings"], dump_matrices=True)
    assert not GridSearch(rec, GRID, verbose=False)._device_backed()
    rec = make_recommender(OracleBacked, golden["ratings"])
    rec.set_options({'k_folds': 3, 'n_iterations': 5})
    assert not GridSearch(rec, GRID, verbose=False)._device_backed()


def test_configurations_are_dealt_to_ranks_gloo(golden):
    import torch.multiprocessing as mp
    from grid_helpers import sweep_under_gloo
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = mp.Manager().dict()
    world = 2
    procs = [mp.get_context("spawn").Process(target=sweep_under_gloo, args=(r, world, port, golden["ratings"], out))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    for r in range(world):
        best, rows, n_dumps, (uv, iv) = out[r]
        check_rows([list(golden["header"])] + rows, golden)
        assert [best['_lambda'], best['n_factors']] == list(golden["best"])
        assert n_dumps == (1 if r == 0 else 0)                    # the table is written once, by rank 0
        # every rank leaves train() holding the LAST configuration's model (its owner broadcasts it)
        assert numpy.array_equal(uv, out[0][3][0]) and numpy.array_equal(iv, out[0][3][1])
