"""
Import the UNMODIFIED reference from /root/reference (build container only).

The reference needs three things this image lacks:
  * ``overrides``        -- lib/collaborative_filtering.py:8 ; a no-op decorator is enough.
  * ``mysql.connector``  -- util/data_parser.py:5 ; never called by the hot path.
  * keras / chainer / lda2vec -- only when lib.recommender_system / runnables are imported.
They are supplied as in-memory stub modules; no reference source is copied or modified.
"""
import importlib
import os
import sys
import types
from unittest import mock

REFERENCE_ROOT = os.environ.get("HYPREC_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "lib", "collaborative_filtering.py"))


def _install_stubs(heavy=False):
    if "overrides" not in sys.modules:
        shim = types.ModuleType("overrides")
        shim.overrides = lambda f: f
        sys.modules["overrides"] = shim
    if "mysql" not in sys.modules:
        mysql = types.ModuleType("mysql")
        mysql.connector = mock.MagicMock(name="mysql.connector")
        sys.modules["mysql"] = mysql
        sys.modules["mysql.connector"] = mysql.connector
    if heavy:
        for name in ("keras", "keras.layers", "keras.models", "keras.callbacks", "keras.regularizers",
                     "keras.layers.core", "keras.layers.convolutional", "keras.layers.embeddings",
                     "chainer", "chainer.links", "chainer.functions", "chainer.optimizers",
                     "lda2vec", "lda2vec.utils", "lda2vec.preprocess", "lda2vec.corpus"):
            if name not in sys.modules:
                sys.modules[name] = mock.MagicMock(name=name)


def load_reference(heavy=False):
    """Return a namespace with the reference classes of the hot path.

    :param bool heavy: also stub keras/chainer/lda2vec so lib.recommender_system imports.
    """
    if not reference_available():
        raise RuntimeError("reference not present at %s (GPU box?)" % REFERENCE_ROOT)
    _install_stubs(heavy)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    ns.CollaborativeFiltering = importlib.import_module("lib.collaborative_filtering").CollaborativeFiltering
    ns.AbstractRecommender = importlib.import_module("lib.abstract_recommender").AbstractRecommender
    ns.Evaluator = importlib.import_module("lib.evaluator").Evaluator
    ns.TopRecommendations = importlib.import_module("util.top_recommendations").TopRecommendations
    ns.ModelInitializer = importlib.import_module("util.model_initializer").ModelInitializer
    ns.GridSearch = importlib.import_module("lib.grid_search").GridSearch
    if heavy:
        ns.RecommenderSystem = importlib.import_module("lib.recommender_system").RecommenderSystem
    return ns
