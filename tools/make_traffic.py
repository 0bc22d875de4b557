"""profiles/traffic.json from an ncu launch list: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the kernel
families bench.py's roofline block names, for ONE epoch / ONE evaluation (the last of the capture), with the command, the
commit and a fingerprint of hyprec_b200/csrc so that bench.py drops the figure when the kernels change.

    python tools/make_traffic.py gpurun_out/launches.csv --workload citeulike-a --precision fp64 \
        --command "ncu --metrics ... python tools/profile_step.py --epochs 3"
"""
import argparse
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

SOLVE = ("als_solve_rows_kernel", "lowrank_small_kernel", "tc_chol_rows_kernel", "tc_chol_multi_kernel", "split_partial_kernel",
         "split_reduce_kernel")


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    out = collections.OrderedDict()
    for r in rows:
        d = out.setdefault(int(r[0]), dict(name=r[4], grid=r[8]))
        d[r[12]] = float(r[14].replace(",", ""))
    return list(out.values())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--workload", default="citeulike-a")
    ap.add_argument("--precision", default="fp64")
    ap.add_argument("--command", default="")
    ap.add_argument("--file", default=None, help="name of the committed copy of the capture under profiles/")
    args = ap.parse_args()
    ls = launches(args.csv)
    evals = [i for i, d in enumerate(ls) if "colsum_partial_kernel" in d["name"]]
    if len(evals) < 2:
        raise SystemExit("need at least two evaluations in the capture")
    epoch = ls[evals[-2]:evals[-1]]
    # the epoch's launches follow the previous evaluation's last kernel (nz_flags)
    last_flag = max(i for i, d in enumerate(epoch) if "nz_flags_kernel" in d["name"])
    epoch = epoch[last_flag + 1:]
    evaluation = ls[evals[-1]:]
    flags_after = [i for i, d in enumerate(evaluation) if "nz_flags_kernel" in d["name"]]
    tail_markers = [i for i, d in enumerate(evaluation) if i > 0 and ("build_base_tiles" in d["name"] or "fma_probe" in d["name"])]
    if tail_markers:                       # a bench.py capture goes on after its last evaluation (end-to-end steps, probes)
        flags_after = [i for i in flags_after if i < tail_markers[0]]
    if flags_after:
        evaluation = evaluation[:flags_after[-1] + 1]
    elif tail_markers:
        evaluation = evaluation[:tail_markers[0]]

    def dram(group):
        return sum(d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0) for d in group)

    def us(group):
        return sum(d.get("gpu__time_duration.sum", 0.0) for d in group) / 1e3

    solve = [d for d in epoch if any(name in d["name"] for name in SOLVE) and d["grid"] != "(1, 1, 1)"]
    sweep = [d for d in evaluation if "tc_score_kernel" in d["name"]][-1:]
    topk = [d for d in evaluation if "eval_topk_kernel" in d["name"]]
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    common = dict(command=args.command, commit=commit, csrc_sha=bench.csrc_fingerprint(), file=args.file or os.path.basename(args.csv))
    base = "%s/%s" % (args.workload, args.precision)
    path = os.path.join(ROOT, "profiles", "traffic.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    table[base + "/solve"] = dict(common, dram_bytes=dram(solve), launches=len(solve), serialised_us=us(solve),
                                  what="row-solve launches of one epoch")
    table[base + "/sweep"] = dict(common, dram_bytes=dram(sweep), launches=len(sweep), serialised_us=us(sweep),
                                  what="the evaluation's tensor-core sweep (counts + emission)")
    table[base + "/topk"] = dict(common, dram_bytes=dram(topk), launches=len(topk), serialised_us=us(topk),
                                 what="eval_topk_kernel pass 1 + pass 2")
    table[base + "/evaluation"] = dict(common, dram_bytes=dram(evaluation), launches=len(evaluation), serialised_us=us(evaluation),
                                       what="every launch of one evaluation")
    table[base + "/epoch"] = dict(common, dram_bytes=dram(epoch), launches=len(epoch), serialised_us=us(epoch),
                                  what="every launch of one epoch")
    with open(path, "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)
    for key in sorted(table):
        if key.startswith(base):
            print("%-32s %8.1f MB  %4d launches  %8.1f us" % (key, table[key]["dram_bytes"] / 1e6, table[key]["launches"], table[key]["serialised_us"]))


if __name__ == "__main__":
    main()
