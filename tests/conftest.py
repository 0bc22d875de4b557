import os
import sys

import numpy
import pytest

# Several handles on one GPU each own 8 streams; with the default 8 hardware queues two handles' streams share a
# queue and one handle's barrier kernel would sit in front of the other's signal (tests/test_gpu_sharded.py runs
# two ranks on one device).  Read by the CUDA runtime when the context is created.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _cuda_device_present():
    try:
        import ctypes
        cuda = ctypes.CDLL("libcuda.so.1")
        count = ctypes.c_int(0)
        return cuda.cuInit(0) == 0 and cuda.cuDeviceGetCount(ctypes.byref(count)) == 0 and count.value > 0
    except OSError:
        return False


def pytest_collection_modifyitems(config, items):
    """Tests marked gpu are skipped, not failed, on a machine without a CUDA device (the library has no CPU
    fallback: hyprec_create raises there)."""
    if _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
