"""Per-phase clock64() breakdown of the direct row solve (block 0), citeulike-a shape.
    HYPREC_SOLVE_PROF=1 HYPREC_LOWRANK=0 python tools/solve_phases.py [--precision fp64|fp32] [--k 200]
"""
import argparse
import os
import sys

import numpy
import scipy.sparse as sp

os.environ.setdefault("HYPREC_SOLVE_PROF", "1")
os.environ.setdefault("HYPREC_LOWRANK", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyprec_b200 import _native, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="fp64")
ap.add_argument("--k", type=int, default=200)
args = ap.parse_args()
full = synth.ratings_csr("citeulike-a")
m, n = full.shape
rs = numpy.random.RandomState(1)
h = _native.Handle(0, _native.F64 if args.precision == "fp64" else _native.F32)
h.set_train_csr(m, n, full.indptr, full.indices, full.data)
h.set_factors(rs.rand(m, args.k), rs.rand(n, args.k))
names = ["gather+rank-1", "base add", "first diag", "panel solves", "next column (warp 0)", "next diag (warp 0)",
         "wait trailing", "back subst"]
for side in (_native.SIDE_USER, _native.SIDE_ITEM):
    h.solve_phase_cycles(reset=True)
    h.als_half_step(side, 0.01)
    cyc = h.solve_phase_cycles(reset=True)[:8].astype(float)
    rows = (m if side == 0 else n) / 148.0
    print("side %d: %.0f cycles per row (block 0, ~%.0f rows)" % (side, cyc.sum() / rows, rows))
    for name, c in zip(names, cyc):
        print("   %-22s %9.0f cycles/row  %5.1f %%" % (name, c / rows, 100 * c / cyc.sum()))
h.close()
