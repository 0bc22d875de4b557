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


def test_clock_sampler_reports_the_clock_under_load():
    """Short timed regions leave most nvidia-smi samples on an idle GPU: the reported clock is the median
    of the samples in the upper half of the power range; the all-samples median is kept next to it."""
    class Proc(object):
        def terminate(self):
            pass

        def wait(self, timeout=None):
            pass

        def kill(self):
            pass

    sampler = bench.ClockSampler(0)
    sampler.proc = Proc()
    idle = ["0", "120", "1965", "80.0", "x", "Not Active", "Not Active", "Not Active", "Not Active"]
    busy = ["0", "1950", "1965", "420.0", "x", "Not Active", "Not Active", "Not Active", "Active"]
    sampler.samples = [idle] * 6 + [busy] * 3
    got = sampler.stop()
    assert got["sm_mhz"] == 1950.0 and got["sm_mhz_all_samples"] == 120.0
    assert got["samples"] == 9 and got["samples_under_load"] == 3
    assert got["reasons"] == ["sw_power_cap"] and got["sm_max_mhz"] == 1965.0
