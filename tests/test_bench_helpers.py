"""bench.py's pure helpers (no GPU): the roofline annotation of the sharded line and the work model."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


@pytest.fixture(scope="module")
def small():
    return bench.build_workload("small")


def test_algorithmic_work_is_consistent(small):
    w8, w4 = bench.algorithmic_work(small, 8), bench.algorithmic_work(small, 4)
    assert w8["nnz"] == small["train_csr"].nnz
    assert w8["lowrank_flops_per_epoch"] > 0 and w4["lowrank_flops_per_epoch"] > 0
    assert w8["lowrank_flops_per_epoch"] <= w8["solve_flops_per_epoch"] * 1.0001     # the low-rank form never costs more
    assert w8["solve_bytes_per_epoch"](8) == 2 * w8["solve_bytes_per_epoch"](4) - 2.0 * w8["nnz"] * 4 - 8.0 * (small["m"] + small["n"] + 2)


@pytest.mark.parametrize("precision,world", [("fp64", 2), ("fp32", 8)])
def test_sharded_roofline_keys(small, precision, world):
    r = bench.sharded_roofline(small, precision, world, 0.5)
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r
    assert r["unit"] == "TFLOP/s" and r["achieved"] > 0 and 0 < r["frac"] < 1
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-15
