"""Per-kernel evidence from the built library: `cuobjdump -sass` opcode counts that show which hardware path a
kernel uses (tcgen05 tensor cores: UTC*MMA; tensor memory: LDTM / STTM; TMA: UTMALDG; mbarriers: SYNCS / UTCBAR;
FP64 / FP32 FMA pipes: DFMA / FFMA; peer / system-scope stores and loads of the exchange: .SYS), plus registers
and shared memory from `cuobjdump -res-usage`.  No GPU needed.

    python tools/sass_summary.py [--out profiles/r02_sass_summary.md]
"""
import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hyprec_b200", "libhyprec_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "SYNCS", "DMMA", "DFMA", "FFMA", "HMMA", "ATOM", "RED", "SHFL", "BAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_sass_summary.md"))
    args = ap.parse_args()
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    usage = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    counts, sys_ops, current = collections.OrderedDict(), collections.Counter(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            current = m.group(1)
            counts[current] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and current:
            op = m.group(1)
            counts[current][op] += 1
            counts[current]["_total"] += 1
            if ".SYS" in m.group(2) and op in ("ST", "STG", "LD", "LDG", "MEMBAR", "ATOM", "ATOMG", "RED"):
                sys_ops[current] += 1
    res = {}
    fn = None
    for line in usage.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        if fn and "REG:" in line:
            reg = re.search(r"REG:(\d+)", line)
            shared = re.search(r"SHARED:(\d+)", line)
            res[fn] = (int(reg.group(1)) if reg else 0, int(shared.group(1)) if shared else 0)
            fn = None
    names = demangle(list(counts))
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    lines = ["# SASS summary of hyprec_b200/libhyprec_b200.so (sm_100a)", "",
             "`python tools/sass_summary.py` at commit `%s`: opcode counts from `cuobjdump -sass`, registers / static shared memory"
             % head, "from `cuobjdump -res-usage` (dynamic shared memory is set at launch).  `UTCHMMA` = tcgen05.mma kind::tf32,",
             "`LDTM` / `STTM` = tcgen05.ld / st (tensor memory), `UTMALDG` = TMA tensor load, `UTCBAR` = tcgen05.commit,",
             "`SYNCS` = mbarrier, `.SYS` = system-scope loads / stores / fences (the multi-GPU exchange).", "",
             "| kernel | instr | regs | smem | " + " | ".join(OPS) + " | .SYS |", "|---|---|---|---|" + "---|" * (len(OPS) + 1)]
    for fn, c in sorted(counts.items(), key=lambda kv: -kv[1]["_total"]):
        nice = names.get(fn, fn)
        nice = re.sub(r"\(.*$", "", nice).replace("void ", "").replace("hyprec::", "").replace("(anonymous namespace)::", "")
        reg, shared = res.get(fn, (0, 0))
        row = [str(sum(v for k, v in c.items() if k.startswith(op))) if op not in ("BAR",) else str(c.get("BAR", 0)) for op in OPS]
        lines.append("| `%s` | %d | %d | %d | %s | %d |" % (nice, c["_total"], reg, shared, " | ".join(row), sys_ops.get(fn, 0)))
    tensor = [names.get(fn, fn) for fn, c in counts.items() if any(k.startswith("UTC") and k.endswith("MMA") for k in c)]
    lines += ["", "Kernels issuing tcgen05 MMAs: %d; kernels with system-scope memory operations: %d." % (len(tensor), len(sys_ops))]
    with open(args.out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(args.out)


if __name__ == "__main__":
    sys.exit(main())
