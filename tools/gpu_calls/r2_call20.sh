#!/bin/bash
mkdir -p gpurun_out
set -x
CMD="python bench.py --scale 0.25 --steps 1 --warmup 1 --no-citeulike --cpu-rows 4 --eval-offset 120000"
timeout 300 $CMD > gpurun_out/r2c20_plain.json 2> gpurun_out/r2c20_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2c20_launches_scaled025.csv $CMD > gpurun_out/r2c20_ncu.log 2>&1
tail -2 gpurun_out/r2c20_ncu.log
