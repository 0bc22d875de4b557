#!/bin/bash
mkdir -p gpurun_out
set -x
python tools/profile_step.py --epochs 3 > gpurun_out/r2c7_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -c 600 --csv --log-file gpurun_out/r2c7_launches_cita_fp64.csv python tools/profile_step.py --epochs 3 > gpurun_out/r2c7_ncu.log 2>&1
