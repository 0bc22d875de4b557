#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_gpu_bench_parity.py tests/test_gpu_parity.py tests/test_zz_device_extras.py -q -x > gpurun_out/r2c11_tests.log 2>&1
python tools/profile_step.py --epochs 3 > gpurun_out/r2c7_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -c 600 --csv --log-file gpurun_out/r2c7_launches_cita_fp64.csv python tools/profile_step.py --epochs 3 > gpurun_out/r2c7_ncu.log 2>&1
timeout 300 python bench.py --workload citeulike-a --steps 10 --warmup 3 > gpurun_out/r2c11_bench_cita.json 2> gpurun_out/r2c11_bench_cita.err
