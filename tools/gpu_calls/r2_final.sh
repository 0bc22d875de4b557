#!/bin/bash
# End-of-round evidence on one B200: GPU tests, smoke, the default bench line, citeulike-t, and ncu launch lists (each
# only after the same command exited 0 without ncu).  Everything lands in gpurun_out/final/.
O=gpurun_out/final
mkdir -p $O
set -x
timeout 900 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -3 $O/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
tail -3 $O/smoke.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
timeout 300 python bench.py --workload citeulike-t --steps 20 --warmup 5 > $O/bench_citeulike_t.json 2> $O/bench_citeulike_t.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
C1="python tools/profile_step.py --epochs 3"
timeout 300 $C1 > $O/profile_step_plain.log 2>&1 && \
timeout 900 ncu --metrics $M --clock-control none -c 700 --csv --log-file $O/launches_citeulike_a_fp64.csv $C1 > $O/profile_step_ncu.log 2>&1
C2="python bench.py --steps 1 --warmup 1 --no-citeulike --cpu-rows 4"
timeout 600 $C2 > $O/bench_scaled1_short.json 2> $O/bench_scaled1_short.err && \
timeout 1500 ncu --metrics $M --clock-control none -c 1500 --csv --log-file $O/launches_scaled1_fp32.csv $C2 > $O/bench_scaled1_ncu.log 2>&1
ls -la $O
