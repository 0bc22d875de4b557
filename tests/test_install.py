"""hyprec_b200.install() rebinds the reference's class names (only where /root/reference exists:
the build container; skipped on the GPU box)."""
import pytest

from oracle import reference_loader


@pytest.mark.skipif(not reference_loader.reference_available(), reason="reference tree not present")
def test_install_rebinds_reference_modules():
    import inspect
    import sys
    ref = reference_loader.load_reference(heavy=True)
    import hyprec_b200
    original = sys.modules["lib.collaborative_filtering"].CollaborativeFiltering
    saved = {(name, attr): getattr(sys.modules[name], attr)
             for name in list(sys.modules) if name.startswith("lib.") and sys.modules[name] is not None
             for attr in ("CollaborativeFiltering", "Evaluator", "GridSearch") if hasattr(sys.modules[name], attr)}
    try:
        patched = hyprec_b200.install()
        assert "lib.collaborative_filtering" in patched and "lib.recommender_system" in patched
        installed = sys.modules["lib.recommender_system"].CollaborativeFiltering
        assert issubclass(installed, hyprec_b200.CollaborativeFiltering) and installed.__name__ == "CollaborativeFiltering"
        assert issubclass(installed, sys.modules["lib.abstract_recommender"].AbstractRecommender)
        assert installed.train is hyprec_b200.CollaborativeFiltering.train       # nothing comes from the reference's base
        assert sys.modules["lib.recommender_system"].Evaluator is hyprec_b200.Evaluator
        assert sys.modules["lib.grid_search"].GridSearch is hyprec_b200.GridSearch
        ref_grid = saved[("lib.grid_search", "GridSearch")]
        assert list(inspect.signature(ref_grid.__init__).parameters) == \
            list(inspect.signature(hyprec_b200.GridSearch.__init__).parameters)[:5]
        for name in ("get_all_combinations", "train", "get_key", "dump_csv", "get_all_errors"):
            assert callable(getattr(hyprec_b200.GridSearch, name)), name
        # same constructor surface as the reference class (collaborative_filtering.py:18-20)
        import inspect
        assert list(inspect.signature(original.__init__).parameters) == \
            list(inspect.signature(hyprec_b200.CollaborativeFiltering.__init__).parameters)
        for name in ("train", "train_k_fold", "train_one_fold", "partial_train", "als_step", "get_predictions",
                     "predict", "set_hyperparameters", "set_options", "set_data", "get_evaluation_report",
                     "rounded_predictions", "recommend_items", "dump_recommendations", "get_ratings",
                     "build_confidence_matrix", "set_item_based_recommender", "_get_options_suffix"):
            assert callable(getattr(hyprec_b200.CollaborativeFiltering, name)), name
    finally:
        for (name, attr), value in saved.items():
            setattr(sys.modules[name], attr, value)
