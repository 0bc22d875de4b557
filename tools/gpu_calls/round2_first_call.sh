#!/bin/bash
# A/B measurements the round-1 captures point at ("what the last capture ... says"), for one
# gpurun call at the start of the next round.  Everything lands in gpurun_out/.
#   gpurun --timeout 900 -- 'bash tools/round2_first_call.sh'
mkdir -p gpurun_out
set -x
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1
# low-rank / direct split of the float64 path at k = 200 (default cap: 191)
for cap in 0 159 127 95 63; do
  HYPREC_LOWRANK_CAP=$cap timeout 60 python tools/profile_step.py --epochs 4 --no-eval > gpurun_out/r2_cap_$cap.log 2>&1
done
# default bench line, float32 line
timeout 300 python bench.py > gpurun_out/r2_bench_fp64.json 2> gpurun_out/r2_bench_fp64.err
timeout 300 python bench.py --precision fp32 > gpurun_out/r2_bench_fp32.json 2> gpurun_out/r2_bench_fp32.err
# the reference-facing call and the sweep
timeout 120 python tools/class_run.py > gpurun_out/r2_class_run.json 2> gpurun_out/r2_class_run.err
timeout 120 python tools/grid_sweep.py --also-single > gpurun_out/r2_grid_sweep.json 2> gpurun_out/r2_grid_sweep.err
# scaled workload with the evaluation of a user block through device-generated candidate lists
timeout 400 python bench.py --workload scaled --scale 0.1 --eval-users 1024 --steps 2 --warmup 1 \
  > gpurun_out/r2_scaled01_eval.json 2> gpurun_out/r2_scaled01_eval.err
