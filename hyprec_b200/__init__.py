"""
hyprec_b200 -- B200-native implementation of HyPRec's weighted-ALS / prediction / top-k evaluation
hot path behind the reference's own ``CollaborativeFiltering`` API.

    from hyprec_b200 import CollaborativeFiltering, Evaluator, GridSearch, ModelInitializer

``install()`` makes the unmodified reference (RecommenderSystem, GridSearch, runnables.py) use the
device path by rebinding ``lib.collaborative_filtering.CollaborativeFiltering``.
"""
import sys

from hyprec_b200.collaborative_filtering import CollaborativeFiltering
from hyprec_b200.evaluator import Evaluator
from hyprec_b200.grid_search import GridSearch
from hyprec_b200.model_initializer import ModelInitializer
from hyprec_b200.top_recommendations import TopRecommendations

__all__ = ["CollaborativeFiltering", "Evaluator", "GridSearch", "ModelInitializer", "TopRecommendations", "install",
           "as_reference_recommender"]


def as_reference_recommender(cls):
    """``cls`` with the reference's ``lib.abstract_recommender.AbstractRecommender`` appended to its bases, so that
    ``isinstance(recommender, AbstractRecommender)`` holds for callers that import the base class from the
    reference (/root/reference/tests/collaborative_tests.py:41 does).  Nothing is inherited from it: every method
    of the reference's base is defined by this package's mirror, which comes first in the MRO."""
    import importlib
    base = importlib.import_module("lib.abstract_recommender").AbstractRecommender
    if issubclass(cls, base):
        return cls
    return type(cls.__name__, (cls, base), {"__module__": cls.__module__, "__doc__": cls.__doc__})


def install(evaluator=True, grid_search=True):
    """Rebind the reference's class names to the implementations of this package in every reference module
    already imported (lib.collaborative_filtering, lib.recommender_system, lib.grid_search, runnables, ...):
    ``CollaborativeFiltering`` always; ``Evaluator`` (bit-identical splits without the per-user Python set
    arithmetic) and ``GridSearch`` (sweep without dense matrices, several models in flight) unless switched
    off.  Returns the names of the patched modules."""
    import importlib
    swaps = [("lib.collaborative_filtering", "CollaborativeFiltering", as_reference_recommender(CollaborativeFiltering))]
    if evaluator:
        swaps.append(("lib.evaluator", "Evaluator", Evaluator))
    if grid_search:
        swaps.append(("lib.grid_search", "GridSearch", GridSearch))
    patched = []
    for home, attr, replacement in swaps:
        original = getattr(importlib.import_module(home), attr)
        if original is replacement:
            continue
        for name, module in list(sys.modules.items()):
            if module is None or not (name == "runnables" or name.startswith("lib.") or name == "__main__"):
                continue
            if getattr(module, attr, None) is original:
                setattr(module, attr, replacement)
                if name not in patched:
                    patched.append(name)
    return patched
