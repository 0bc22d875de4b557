#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_gpu_sharded.py -q -s -x > gpurun_out/r2c4_sharded_tests.log 2>&1
timeout 900 python -m pytest tests/test_gpu_bench_parity.py -q -s > gpurun_out/r2c4_parity_tests.log 2>&1
timeout 900 python bench.py --scale 1.0 --steps 3 --warmup 2 --no-citeulike > gpurun_out/r2c4_bench_scaled1_n1.json 2> gpurun_out/r2c4_bench_scaled1_n1.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv > gpurun_out/r2c4_mem.txt
