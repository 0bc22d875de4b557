import os
import sys

import numpy
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def load_golden(name):
    return dict(numpy.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def split_lists(flat, lengths):
    out, at = [], 0
    for n in lengths:
        out.append(flat[at:at + int(n)])
        at += int(n)
    return out


def golden_folds(g):
    """Per-fold dicts of a golden ALS case."""
    folds = []
    for f in range(int(g["n_folds_run"])):
        folds.append({key[len("f%d_" % f):]: val for key, val in g.items() if key.startswith("f%d_" % f)})
    return folds


@pytest.fixture(scope="session")
def als_cases():
    return {name: load_golden(name) for name in
            ("als_kfold_4x30", "als_kfold_toy", "als_theta_toy", "als_naive_toy")}
