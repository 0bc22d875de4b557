#!/bin/bash
# Last GPU call of round 2 (15 GPU-minutes left): verify HEAD on one B200, most important first, every step bounded.
O=gpurun_out/last
mkdir -p $O
set -x
timeout 420 python -m pytest tests -m gpu -x -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -4 $O/gpu_tests.log
timeout 120 python tools/class_run.py --sparse > $O/class_run_sparse.json 2> $O/class_run_sparse.err; echo "class sparse rc=$?"
cat $O/class_run_sparse.json
timeout 120 python tools/class_run.py > $O/class_run_dense.json 2> $O/class_run_dense.err; echo "class dense rc=$?"
cat $O/class_run_dense.json
timeout 420 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
cut -c1-600 $O/bench_default.json
timeout 120 python tools/grid_sweep.py --workers 3 --also-single > $O/grid_sweep.json 2> $O/grid_sweep.err; echo "sweep rc=$?"
cat $O/grid_sweep.json
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
tail -3 $O/smoke.log
timeout 200 python bench.py --workload citeulike-t --steps 20 --warmup 5 > $O/bench_citeulike_t.json 2> $O/bench_citeulike_t.err; echo "bench t rc=$?"
ls -la $O
