#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_gpu_bench_parity.py tests/test_gpu_parity.py tests/test_zz_device_extras.py -q -x > gpurun_out/r2c8_tests.log 2>&1
python tools/profile_step.py --epochs 4 > gpurun_out/r2c8_profile_step.log 2>&1
HYPREC_FUSED_TOPK=0 python tools/profile_step.py --epochs 4 > gpurun_out/r2c8_profile_step_exact.log 2>&1
timeout 300 python bench.py --workload citeulike-a --steps 10 --warmup 3 > gpurun_out/r2c8_bench_cita.json 2> gpurun_out/r2c8_bench_cita.err
