#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c13_gpu_tests.log 2>&1
timeout 300 python tools/class_run.py > gpurun_out/r2c13_class_run_dense.json 2> gpurun_out/r2c13_class_run_dense.err
timeout 300 python tools/class_run.py --sparse > gpurun_out/r2c13_class_run_sparse.json 2> gpurun_out/r2c13_class_run_sparse.err
timeout 300 python tools/grid_sweep.py --also-single > gpurun_out/r2c13_grid_sweep.json 2> gpurun_out/r2c13_grid_sweep.err
timeout 300 python bench.py --workload citeulike-a --steps 10 --warmup 3 > gpurun_out/r2c13_bench_cita.json 2> gpurun_out/r2c13_bench_cita.err
