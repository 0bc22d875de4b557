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
    assert w8["nnz"] == small["train_csr"].nnz == int(small["train_arrays"][0][-1])
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


def test_bench_line_python_flow_with_a_stand_in_handle(monkeypatch, capsys):
    """bench.py's own arm end to end on the CPU with a stand-in for the C-ABI handle: every key of the JSON
    contract is present and serialisable, including the parts computed last (recall@{50..300})."""
    import json
    import numpy
    import bench
    from hyprec_b200 import _native

    class StandIn(object):
        launch_count = 0

        def __init__(self, device, precision):
            self.m = self.n = self.k = 0

        def set_train_csr(self, m, n, *rest):
            self.m, self.n = m, n

        def set_factors(self, u, v):
            self.k = u.shape[1]

        def als_half_step(self, *args):
            StandIn.launch_count += 10

        def eval_topk(self, capacity, cp=None, ci=None, user_lo=0, user_hi=None, want_counts=True, want_topk=True,
                      n_train_nz=None, n_test_nz=None, out=None):
            if out is not None:
                return out
            nu = (self.m if user_hi is None else user_hi) - user_lo
            return dict(thresholds=numpy.zeros(nu), ones_per_row=numpy.zeros(nu, numpy.int64),
                        topk_idx=numpy.zeros((nu, capacity), numpy.int32), topk_val=numpy.zeros((nu, capacity)),
                        topk_hit=numpy.ones((nu, capacity)), train_flags=None, test_flags=None)

        def pinned_empty(self, shape, dtype=numpy.float64):
            return numpy.zeros(shape, dtype)

        def get_factors(self, u=None, v=None):
            return u, v

        def kernel_ms(self, which):
            return 1.0, 1, 1.0

        def rmse(self):
            return 0.25

        def elapsed_ms(self, start, stop):
            return 1.5

        def eval_stats(self):
            return dict(fused=True, fallback_users=0)

        def probe_fma_tflops(self, precision):
            return 36.0

        def __getattr__(self, name):              # set_eval_csr, set_candidates, flush_l2, reset_kernel_ms, close ...
            return lambda *args, **kwargs: None

    monkeypatch.setattr(_native, "Handle", StandIn)
    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: None)

    class Args(object):
        gpus, steps, warmup, impl, workload, scale, precision, cpu_rows, eval_users = 1, 2, 1, "ours", "small", 0.1, None, 2, 0
        no_citeulike = False
        eval_offset = -1

    args = Args()
    bench.our_arm(args, bench.make_workload(args))
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["config"]["workload"].startswith("small shape") and "model" not in line["config"]
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert sorted(line["recall_at"], key=int) == ["50", "100", "150", "200", "250", "300"]
    assert line["gpu_launches"] == 2 * 2 * 10
    assert line["value"] == 1.5 and line["n_gpus"] == 1 and line["scaling"] == "strong"


def test_scaled_generator_matches_between_devices_and_partitions():
    """The scaled workload's arrays (torch sort / searchsorted on whatever device) are CSR-consistent, and the
    sub-matrix helpers of the CPU baseline agree with scipy."""
    import numpy
    import scipy.sparse as sp
    wl = bench.build_scaled_workload(0.002, device="cpu")
    m, n = wl["m"], wl["n"]
    assert (m, n, wl["k"]) == (2000, 4000, 256) and wl["device_init"] == bench.FACTOR_SEED
    full, train, test = wl["full_arrays"], wl["train_arrays"], wl["test_arrays"]
    assert full[0][-1] == train[0][-1] + test[0][-1]
    assert 0.75 < train[0][-1] / float(full[0][-1]) < 0.85
    csr = sp.csr_matrix((numpy.ones(len(train[1])), train[1], train[0]), shape=(m, n))
    assert csr.has_sorted_indices or (numpy.diff(csr.indices) > 0)[numpy.diff(csr.indices) <= 0].size == 0
    rows = numpy.array([0, 5, 77, 1999])
    cols = numpy.array([1, 17, 3999])
    assert (bench.sub_csr(train[0], train[1], rows, n) != csr[rows]).nnz == 0
    assert (bench.sub_csc(train[0], train[1], cols, m) != csr.T.tocsr()[cols]).nnz == 0
    work = bench.algorithmic_work(wl, 4)
    assert work["nnz"] == train[0][-1] and work["lowrank_flops_per_epoch"] <= work["solve_flops_per_epoch"] * 1.0001
