#!/bin/bash
O=gpurun_out/ncu_final
mkdir -p $O
C="python tools/profile_step.py --epochs 3"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 60 $C > $O/profile_step_plain.log 2>&1 && \
timeout 100 ncu --metrics $M --clock-control none -c 500 --csv --log-file $O/launches_citeulike_a_fp64.csv $C > $O/profile_step_ncu.log 2>&1
tail -4 $O/profile_step_plain.log
