#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c12_gpu_tests.log 2>&1
timeout 300 python bench.py --workload citeulike-a --steps 10 --warmup 3 > gpurun_out/r2c12_bench_cita.json 2> gpurun_out/r2c12_bench_cita.err
timeout 300 python tools/class_run.py > gpurun_out/r2c12_class_run_dense.json 2> gpurun_out/r2c12_class_run_dense.err
timeout 300 python tools/class_run.py --sparse > gpurun_out/r2c12_class_run_sparse.json 2> gpurun_out/r2c12_class_run_sparse.err
HYPREC_PRECISION=fp32 timeout 600 python tools/class_run.py --scale 0.25 --folds 1 --factors 256 --iterations 3 > gpurun_out/r2c12_class_run_scaled025.json 2> gpurun_out/r2c12_class_run_scaled025.err
