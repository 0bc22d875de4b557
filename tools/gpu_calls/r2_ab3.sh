#!/bin/bash
O=gpurun_out/ab3
mkdir -p $O
B2="HYPREC_BUCKET_TILE=4;4;4;8;8;8,HYPREC_BUCKET_THREADS=64;128;192;256;256;256"
B12="HYPREC_BUCKET_TILE=4;4;4;8;8;8,HYPREC_BUCKET_THREADS=64;96;256;256;256;256"
timeout 200 python tools/ab_epoch.py --epochs 20 \
  "la:HYPREC_LOOKAHEAD=1" "b2:$B2" "la_b2:HYPREC_LOOKAHEAD=1,$B2" "b12:$B12" "la_b12:HYPREC_LOOKAHEAD=1,$B12" \
  "default2:" "la2:HYPREC_LOOKAHEAD=1" "b2_2:$B2" "la_b2_2:HYPREC_LOOKAHEAD=1,$B2" "b12_2:$B12" "la_b12_2:HYPREC_LOOKAHEAD=1,$B12" \
  > $O/ab_citeulike_a.jsonl 2> $O/ab_citeulike_a.err
cut -c1-330 $O/ab_citeulike_a.jsonl
timeout 200 python tools/ab_epoch.py --shape citeulike-t --k 300 --epochs 10 \
  "la:HYPREC_LOOKAHEAD=1" "la_b2:HYPREC_LOOKAHEAD=1,$B2" "default2:" "la2:HYPREC_LOOKAHEAD=1" \
  > $O/ab_citeulike_t.jsonl 2> $O/ab_citeulike_t.err
cut -c1-330 $O/ab_citeulike_t.jsonl
HYPREC_LOOKAHEAD=1 timeout 400 python -m pytest tests -m gpu -x -q > $O/gpu_tests_lookahead.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests_lookahead.log
tail -4 $O/gpu_tests_lookahead.log
HYPREC_LOOKAHEAD=1 HYPREC_SOLVE_PROF=1 HYPREC_LOWRANK=0 timeout 120 python tools/solve_phases.py > $O/solve_phases_direct_fp64_lookahead.log 2>&1
cat $O/solve_phases_direct_fp64_lookahead.log
