#!/bin/bash
O=gpurun_out/ab1
mkdir -p $O
D="64;128;128;256;256;256"
timeout 200 python tools/ab_epoch.py --epochs 20 \
  "b2_8x256:HYPREC_BUCKET_THREADS=64;128;256;256;256;256" \
  "b2_4x256:HYPREC_BUCKET_TILE=4;4;4;8;8;8,HYPREC_BUCKET_THREADS=64;128;256;256;256;256" \
  "b2_4x512:HYPREC_BUCKET_TILE=4;4;4;8;8;8,HYPREC_BUCKET_THREADS=64;128;512;256;256;256" \
  "b1_4x256:HYPREC_BUCKET_THREADS=64;256;128;256;256;256" \
  "b1_4x256_b2_4x512:HYPREC_BUCKET_TILE=4;4;4;8;8;8,HYPREC_BUCKET_THREADS=64;256;512;256;256;256" \
  "cap191:HYPREC_LOWRANK_CAP=191" \
  "cap191_b23_4x512:HYPREC_LOWRANK_CAP=191,HYPREC_BUCKET_TILE=4;4;4;4;8;8,HYPREC_BUCKET_THREADS=64;128;512;512;256;256" \
  "cap191_b3_8x512:HYPREC_LOWRANK_CAP=191,HYPREC_BUCKET_THREADS=64;128;128;512;256;256" \
  "cap63:HYPREC_LOWRANK_CAP=63" \
  > $O/ab_citeulike_a.jsonl 2> $O/ab_citeulike_a.err
cat $O/ab_citeulike_a.jsonl | cut -c1-420
timeout 200 python tools/ab_epoch.py --shape citeulike-t --k 300 --epochs 10 \
  "b2_4x512:HYPREC_BUCKET_TILE=4;4;4;8;8;8,HYPREC_BUCKET_THREADS=64;128;512;256;256;256" \
  "b1_4x256_b2_4x512:HYPREC_BUCKET_TILE=4;4;4;8;8;8,HYPREC_BUCKET_THREADS=64;256;512;256;256;256" \
  > $O/ab_citeulike_t.jsonl 2> $O/ab_citeulike_t.err
cat $O/ab_citeulike_t.jsonl | cut -c1-420
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 240 ncu --metrics $M --clock-control none -c 500 --csv --log-file $O/launches_citeulike_a_fp64.csv python tools/profile_step.py --epochs 3 > $O/profile_step_ncu.log 2>&1
tail -3 $O/profile_step_ncu.log
