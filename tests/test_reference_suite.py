"""The reference's OWN unit tests (/root/reference/tests/*.py, loaded unmodified from the read-only tree -- build
container only) run against this package's classes: every name the test modules import from the reference
(CollaborativeFiltering, Evaluator, GridSearch, ModelInitializer, LinearRegression, TopRecommendations) is rebound to
its counterpart here, as ``hyprec_b200.install()`` does for the application modules.  On the CPU the two device
entry points of the recommender are answered by the oracle (tests/grid_helpers.OracleBacked); everything else --
splits, checkpoints, sweep, blend, metrics, the recommender's host flow -- is the product code.

Two shims cover API removals between the reference's era and this image, like the ``overrides`` / ``mysql`` stubs of
oracle/reference_loader.py: ``unittest.TestCase.assertEquals`` (gone in Python 3.12, linear_regression_tests.py:48) and
``numpy.alltrue`` (gone in numpy 2.0, util_tests.py:124).

Not run: metrics_tests.py (its fixture is a ragged ``numpy.array``, an error on numpy >= 1.24 before any class of
either implementation is reached -- tests/test_oracle_pin.py holds its known answers), content_based_tests.py and
recommender_tests.py (content models / orchestration, out of scope)."""
import functools
import importlib
import unittest
import warnings

import pytest

from oracle import reference_loader

MODULES = ["top_recommendations_tests", "evaluator_tests", "collaborative_tests", "grid_search_tests",
           "linear_regression_tests", "util_tests"]


@pytest.mark.skipif(not reference_loader.reference_available(), reason="reference tree not present")
@pytest.mark.parametrize("name", MODULES)
def test_reference_test_module_passes_on_this_package(name, tmp_path, monkeypatch):
    pytest.importorskip("sklearn.metrics")
    import hyprec_b200
    from grid_helpers import OracleBacked
    from hyprec_b200.linear_regression import LinearRegression
    reference_loader.load_reference()
    module = importlib.import_module("tests." + name)
    assert module.__file__.startswith(reference_loader.REFERENCE_ROOT)
    monkeypatch.chdir(tmp_path)
    import numpy
    monkeypatch.setattr(unittest.TestCase, "assertEquals", unittest.TestCase.assertEqual, raising=False)
    monkeypatch.setattr(numpy, "alltrue", numpy.all, raising=False)

    class ScratchInitializer(hyprec_b200.ModelInitializer):           # checkpoints go to a scratch folder
        def __init__(self, config, n_iterations, verbose=False, folder=None):
            hyprec_b200.ModelInitializer.__init__(self, config, n_iterations, verbose, folder or str(tmp_path))

    class ScratchGridSearch(hyprec_b200.GridSearch):
        def dump_csv(self, all_results):
            self.dumped = all_results

    oracle_backed = hyprec_b200.as_reference_recommender(
        type("CollaborativeFiltering", (OracleBacked,), {}))               # the class name is what __repr__ prints
    replacements = dict(CollaborativeFiltering=oracle_backed, Evaluator=hyprec_b200.Evaluator,
                        GridSearch=ScratchGridSearch, ModelInitializer=ScratchInitializer,
                        LinearRegression=LinearRegression, TopRecommendations=hyprec_b200.TopRecommendations)
    swapped = 0
    for attr, replacement in replacements.items():
        if hasattr(module, attr):
            monkeypatch.setattr(module, attr, replacement)
            swapped += 1
    assert swapped > 0
    suite = unittest.defaultTestLoader.loadTestsFromModule(module)
    assert suite.countTestCases() > 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                               # the reference runs its tests with -W ignore
        result = unittest.TextTestRunner(stream=open("/dev/null", "w"), verbosity=0).run(suite)
    problems = [str(trace) for _, trace in result.errors + result.failures]
    assert result.wasSuccessful(), "\n".join(problems)


@pytest.mark.skipif(not reference_loader.reference_available(), reason="reference tree not present")
def test_runnables_collaborative_prints_the_same_summary(monkeypatch):
    """`python3 runnables.py collaborative` on the reference's built-in toy data (runnables.py:38-58, 111-123;
    hyper-parameters from its config/recommender.json): the summary line with the reference's classes, then with this
    package's classes bound to the names runnables.py imported."""
    import contextlib
    import io
    import hyprec_b200
    from grid_helpers import OracleBacked
    reference_loader.load_reference(heavy=True)
    runnables = importlib.import_module("runnables")

    def summary():
        out = io.StringIO()
        with warnings.catch_warnings(), contextlib.redirect_stdout(out):
            warnings.simplefilter("ignore")
            runner = runnables.RunnableRecommenders(use_database=False, verbose=False, load_matrices=False, dump=False)
            runner.run_collaborative()
        return out.getvalue()

    want = summary()
    assert want.startswith("Summary: Test sum")
    monkeypatch.setattr(runnables, "CollaborativeFiltering",
                        hyprec_b200.as_reference_recommender(type("CollaborativeFiltering", (OracleBacked,), {})))
    monkeypatch.setattr(runnables, "Evaluator", hyprec_b200.Evaluator)
    monkeypatch.setattr(runnables, "ModelInitializer", hyprec_b200.ModelInitializer)
    assert summary() == want
