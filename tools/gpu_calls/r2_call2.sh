#!/bin/bash
# round 2, call 2: new sharded / split / bench code on one GPU
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_bench_parity.py -x -q -s > gpurun_out/r2c2_new_tests.log 2>&1
timeout 300 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_sharded.py --deselect tests/test_gpu_bench_parity.py > gpurun_out/r2c2_old_tests.log 2>&1
timeout 600 python bench.py --scale 0.1 --steps 3 --warmup 2 > gpurun_out/r2c2_bench_scaled01.json 2> gpurun_out/r2c2_bench_scaled01.err
timeout 300 python bench.py --workload citeulike-a --steps 5 --warmup 3 > gpurun_out/r2c2_bench_cita.json 2> gpurun_out/r2c2_bench_cita.err
