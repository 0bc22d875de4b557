#!/bin/bash
O=gpurun_out/ab5
mkdir -p $O
HYPREC_LOOKAHEAD=2 HYPREC_SOLVE_PROF=1 HYPREC_LOWRANK=0 timeout 120 python tools/solve_phases.py > $O/solve_phases_direct_fp64_rcp_chain.log 2>&1
cat $O/solve_phases_direct_fp64_rcp_chain.log
timeout 200 python tools/ab_epoch.py --epochs 20 \
  "rcp:HYPREC_LOOKAHEAD=2" "rcp_la:HYPREC_LOOKAHEAD=3" "default2:" "rcp2:HYPREC_LOOKAHEAD=2" "rcp_la2:HYPREC_LOOKAHEAD=3" \
  "default3:" "rcp3:HYPREC_LOOKAHEAD=2" \
  > $O/ab_citeulike_a.jsonl 2> $O/ab_citeulike_a.err
cut -c1-300 $O/ab_citeulike_a.jsonl
timeout 200 python tools/ab_epoch.py --shape citeulike-t --k 300 --epochs 10 \
  "rcp:HYPREC_LOOKAHEAD=2" "default2:" "rcp2:HYPREC_LOOKAHEAD=2" "rcp_la:HYPREC_LOOKAHEAD=3" \
  > $O/ab_citeulike_t.jsonl 2> $O/ab_citeulike_t.err
cut -c1-300 $O/ab_citeulike_t.jsonl
HYPREC_LOOKAHEAD=2 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_parity.py -x -q > $O/gpu_tests_rcp.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests_rcp.log
tail -3 $O/gpu_tests_rcp.log
