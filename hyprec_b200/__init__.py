"""
hyprec_b200 -- B200-native implementation of HyPRec's weighted-ALS / prediction / top-k evaluation
hot path behind the reference's own ``CollaborativeFiltering`` API.

    from hyprec_b200 import CollaborativeFiltering, Evaluator, ModelInitializer

``install()`` makes the unmodified reference (RecommenderSystem, GridSearch, runnables.py) use the
device path by rebinding ``lib.collaborative_filtering.CollaborativeFiltering``.
"""
import sys

from hyprec_b200.collaborative_filtering import CollaborativeFiltering
from hyprec_b200.evaluator import Evaluator
from hyprec_b200.model_initializer import ModelInitializer
from hyprec_b200.top_recommendations import TopRecommendations

__all__ = ["CollaborativeFiltering", "Evaluator", "ModelInitializer", "TopRecommendations", "install"]


def install():
    """Rebind the reference's class names to the device implementation in every reference module
    already imported (lib.collaborative_filtering, lib.recommender_system, runnables, ...)."""
    import importlib
    target = importlib.import_module("lib.collaborative_filtering")
    original = target.CollaborativeFiltering
    patched = []
    for name, module in list(sys.modules.items()):
        if module is None or not (name == "runnables" or name.startswith("lib.") or name == "__main__"):
            continue
        if getattr(module, "CollaborativeFiltering", None) is original:
            setattr(module, "CollaborativeFiltering", CollaborativeFiltering)
            patched.append(name)
    return patched
