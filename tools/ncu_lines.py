"""Top source lines of an ncu report by warp-stall samples (cuda,sass correlated source page).
    python tools/ncu_lines.py report.ncu-rep [source file] [top N]
"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
src_path = sys.argv[2] if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks, cur, fname = [], None, None
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1]
    elif r and r[0] == "Function Name":
        cur = dict(file=fname, func=r[1], lines=collections.OrderedDict())
        blocks.append(cur)
    elif cur is not None and len(r) > 4 and r[0].isdigit():
        try:
            cur["lines"][int(r[0])] = cur["lines"].get(int(r[0]), 0) + int(r[4])
        except ValueError:
            pass
for b in blocks:
    total = sum(b["lines"].values())
    if total < 1000:
        continue
    print("== %s :: %s (%d samples)" % (b["file"], b["func"][:70], total))
    try:
        src = open(b["file"]).read().splitlines()
    except OSError:
        src = []
    for ln, v in sorted(b["lines"].items(), key=lambda x: -x[1])[:top]:
        text = src[ln - 1].strip()[:100] if ln - 1 < len(src) else ""
        print("%5d %6.2f%%  %s" % (ln, 100.0 * v / total, text))
