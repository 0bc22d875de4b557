#!/bin/bash
# Last seconds of the round's GPU budget: the rebuilt library (host entry points moved, threaded split) through smoke()
# and the class run, twice (first process / second process).
O=gpurun_out/last3
mkdir -p $O
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
tail -3 $O/smoke.log
timeout 30 python tools/class_run.py --sparse > $O/class_run_sparse_a.json 2> $O/class_run_sparse_a.err
timeout 30 python tools/class_run.py --sparse > $O/class_run_sparse_b.json 2> $O/class_run_sparse_b.err
cut -c1-420 $O/class_run_sparse_a.json; cut -c1-420 $O/class_run_sparse_b.json
