#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c15_gpu_tests.log 2>&1
python tools/profile_step.py --epochs 4 > gpurun_out/r2c15_profile_step.log 2>&1
HYPREC_SMALL_F64=0 python tools/profile_step.py --epochs 4 --no-eval > gpurun_out/r2c15_profile_step_nosmall.log 2>&1
HYPREC_RMSE_FUSED=0 python tools/profile_step.py --epochs 4 --no-eval > gpurun_out/r2c15_profile_step_nofused.log 2>&1
timeout 300 python bench.py --workload citeulike-a --steps 10 --warmup 3 > gpurun_out/r2c15_bench_cita.json 2> gpurun_out/r2c15_bench_cita.err
timeout 300 python bench.py --workload citeulike-t --steps 5 --warmup 3 > gpurun_out/r2c15_bench_citt.json 2> gpurun_out/r2c15_bench_citt.err
timeout 300 python tools/class_run.py --sparse > gpurun_out/r2c15_class_run_sparse.json 2> gpurun_out/r2c15_class_run_sparse.err
