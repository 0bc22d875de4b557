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

__all__ = ["CollaborativeFiltering", "Evaluator", "GridSearch", "ModelInitializer", "TopRecommendations", "install"]


def install(evaluator=True, grid_search=True):
    """Rebind the reference's class names to the implementations of this package in every reference module
    already imported (lib.collaborative_filtering, lib.recommender_system, lib.grid_search, runnables, ...):
    ``CollaborativeFiltering`` always; ``Evaluator`` (bit-identical splits without the per-user Python set
    arithmetic) and ``GridSearch`` (sweep without dense matrices, several models in flight) unless switched
    off.  Returns the names of the patched modules."""
    import importlib
    swaps = [("lib.collaborative_filtering", "CollaborativeFiltering", CollaborativeFiltering)]
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
