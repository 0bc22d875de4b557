#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c6_gpu_tests.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c6_bench_default.json 2> gpurun_out/r2c6_bench_default.err
