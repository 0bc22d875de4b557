#!/usr/bin/env python
"""Markdown table of one epoch + one evaluation from an ncu launch list (gpu__time_duration.sum, dram__bytes_read.sum,
dram__bytes_write.sum per launch): kernel, grid x block, stream, microseconds (serialised, cold caches: shares, not
absolutes), DRAM bytes and the DRAM rate while the launch ran.

    python tools/launch_table.py profiles/r02_final_launches_citeulike_a_fp64.csv > profiles/r02_final_launch_table_citeulike_a_fp64.md
"""
import collections
import csv
import re
import sys

PEAK_GBS = 6544.3        # MEASURED_PEAKS.json


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    launches = collections.OrderedDict()
    for r in rows:
        d = launches.setdefault(int(r[0]), dict(name=re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("hyprec::", ""),
                                                block=r[7], grid=r[8], stream=r[6]))
        value = float(r[14].replace(",", ""))
        unit = r[13]
        if r[12] == "gpu__time_duration.sum":
            d["us"] = value / 1e3 if unit in ("ns", "nsecond") else (value * 1e3 if unit in ("ms", "msecond") else value)
        else:
            scale = dict(byte=1.0, Kbyte=1e3, Mbyte=1e6, Gbyte=1e9).get(unit, 1.0)
            d["dram"] = d.get("dram", 0.0) + value * scale
    seq = list(launches.values())
    evals = [i for i, d in enumerate(seq) if "colsum_partial_kernel" in d["name"]]
    start = max(i for i in range(evals[-2], evals[-1]) if "nz_flags_kernel" in seq[i]["name"]) + 1
    part = seq[start:]
    total = sum(d["us"] for d in part)
    print("# Launches of the last epoch + evaluation in `%s`\n" % path)
    print("Serialised, cold-cache ncu durations (`--clock-control none`): read the SHARES.  DRAM = read + written bytes of the launch; "
          "`GB/s` = those bytes over the launch's duration (HBM peak measured on this pool: %.0f GB/s).\n" % PEAK_GBS)
    print("| # | kernel | grid x block | stream | us | share | DRAM MB | GB/s | of HBM peak |")
    print("|---|---|---|---|---|---|---|---|---|")
    phase = "epoch"
    for n, d in enumerate(part):
        if "colsum_partial_kernel" in d["name"] and phase == "epoch":
            phase = "evaluation"
            print("| | **evaluation** | | | | | | | |")
        gbs = d.get("dram", 0.0) / (d["us"] * 1e-6) / 1e9 if d["us"] > 0 else 0.0
        print("| %d | `%s` | %s x %s | %s | %.1f | %.1f %% | %.2f | %.0f | %.2f |" % (
            n, d["name"][:70], d["grid"].replace(" ", ""), d["block"].replace(" ", ""), d["stream"], d["us"], 100 * d["us"] / total,
            d.get("dram", 0.0) / 1e6, gbs, gbs / PEAK_GBS))
    by = collections.OrderedDict()
    for d in part:
        key = re.sub(r"<.*", "", d["name"])
        e = by.setdefault(key, [0, 0.0, 0.0])
        e[0] += 1
        e[1] += d["us"]
        e[2] += d.get("dram", 0.0)
    print("\n| kernel family | launches | us | share | DRAM MB |\n|---|---|---|---|---|")
    for key, e in sorted(by.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f %% | %.1f |" % (key, e[0], e[1], 100 * e[1] / total, e[2] / 1e6))
    print("\nTotal %.1f us serialised (the epoch's launches overlap on 8 streams when not under ncu: 1.81 ms per epoch and "
          "1.51 ms per evaluation event-timed, `r02_final_bench_default.json`)." % total)


if __name__ == "__main__":
    main()
