"""Per-phase clock64() breakdown of every launch class of the row solve (block 0 of each class) on the
scaled synthetic workload: direct heavy rows (slot 7) and the low-rank degree buckets (slots 1..6).
    python tools/solve_phases_scaled.py [--scale 0.1] [--precision fp32]
"""
import argparse
import os
import sys

import numpy

os.environ["HYPREC_SOLVE_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from hyprec_b200 import _native  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=0.1)
ap.add_argument("--precision", default="fp32")
ap.add_argument("--slots", default="3,4,5,6,7")
args = ap.parse_args()
wl = bench.build_scaled_workload(args.scale)
tr = wl["train_csr"]
m, n, k = wl["m"], wl["n"], wl["k"]
names = ["gather+rank-1", "base add", "first diag", "panel solves", "-", "next diag (warp 0)", "wait trailing",
         "back subst", "rows", "sum deg", "finalize"]
for slot in [int(s) for s in args.slots.split(",")]:
    os.environ["HYPREC_SOLVE_PROF_SLOT"] = str(slot)
    h = _native.Handle(0, _native.F64 if args.precision == "fp64" else _native.F32)
    h.set_train_csr(m, n, tr.indptr, tr.indices, tr.data)
    h.set_factors(wl["u0"], wl["v0"])
    for side in (_native.SIDE_USER, _native.SIDE_ITEM):
        h.solve_phase_cycles(reset=True)
        h.als_half_step(side, wl["lam"])
        cyc = h.solve_phase_cycles(reset=True).astype(float)
        rows = max(cyc[8], 1.0)
        tot = cyc[:8].sum() + cyc[10]
        print("slot %d side %d: block 0 did %d rows, mean degree %.1f, %.0f cycles per row" %
              (slot, side, cyc[8], cyc[9] / rows, tot / rows))
        for i in (0, 1, 2, 3, 5, 6, 7, 10):
            print("   %-22s %9.0f cycles/row  %5.1f %%" % (names[i], cyc[i] / rows, 100 * cyc[i] / max(tot, 1.0)))
    h.close()
