#!/bin/bash
# Final verification of HEAD on one B200 (round 2): GPU tests, default bench line, citeulike-t, the class run, the sweep.
O=gpurun_out/last2
mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -3 $O/gpu_tests.log
timeout 60 python tools/ab_epoch.py --epochs 20 "again:" "old_shapes:HYPREC_BUCKET_TILE=4;4;8;8;8;8,HYPREC_BUCKET_THREADS=64;128;128;256;256;256" > $O/ab_default.jsonl 2>&1
cut -c1-300 $O/ab_default.jsonl
timeout 300 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
timeout 100 python bench.py --workload citeulike-t --steps 20 --warmup 5 > $O/bench_citeulike_t.json 2> $O/bench_citeulike_t.err; echo "bench t rc=$?"
timeout 60 python tools/class_run.py > $O/class_run_dense.json 2> $O/class_run_dense.err
timeout 60 python tools/class_run.py --sparse > $O/class_run_sparse.json 2> $O/class_run_sparse.err
cut -c1-330 $O/class_run_dense.json; cut -c1-330 $O/class_run_sparse.json
timeout 60 python tools/grid_sweep.py --workers 3 > $O/grid_sweep.json 2> $O/grid_sweep.err
cut -c1-330 $O/grid_sweep.json
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
tail -3 $O/smoke.log
